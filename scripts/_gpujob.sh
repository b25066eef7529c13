timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "conv" > gpurun_out/r2_t17.txt 2>&1; tail -4 gpurun_out/r2_t17.txt
timeout 900 python -m pytest tests/test_step_gpu.py -q -x > gpurun_out/r2_t18.txt 2>&1; tail -2 gpurun_out/r2_t18.txt
for v in 0 1 0 1; do
DSGAN_NM_CONV=$v python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench_nm$v.json 2> gpurun_out/r2_bench_nm$v.err; echo "nm=$v: $(cut -c75-175 gpurun_out/r2_bench_nm$v.json)"
done
python scripts/bench_kernels.py --only smallch --out gpurun_out/r2_k_sc1.jsonl 2>&1 | cut -c1-100
