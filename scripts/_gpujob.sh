timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "conv" > gpurun_out/r2_t17.txt 2>&1; tail -2 gpurun_out/r2_t17.txt
python scripts/bench_kernels.py --only smallch --out gpurun_out/r2_k_sc2.jsonl 2>&1 | grep "64->3" | cut -c1-100
for v in 0 1 0 1; do
DSGAN_NM_NOUT=$v python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "nout=$v $(cut -c75-175 gpurun_out/r2_bench12.json)"
done
