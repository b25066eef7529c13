timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "conv" > gpurun_out/r2_t17.txt 2>&1; tail -2 gpurun_out/r2_t17.txt
python scripts/bench_kernels.py --only smallch --out gpurun_out/r2_k_sc4.jsonl 2>&1 | grep "wgrad" | cut -c1-100
for v in 1 2; do
python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err; echo "$(cut -c75-175 gpurun_out/r2_bench18.json)"
done
