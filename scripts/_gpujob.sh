timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "instance_norm" > gpurun_out/r2_t20.txt 2>&1; tail -4 gpurun_out/r2_t20.txt
timeout 900 python -m pytest tests/test_step_gpu.py -q -x > gpurun_out/r2_t18.txt 2>&1; tail -2 gpurun_out/r2_t18.txt
for v in 0 1 0 1; do
DSGAN_IN_FUSED=$v python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; echo "fused=$v $(cut -c75-175 gpurun_out/r2_bench14.json)"
done
