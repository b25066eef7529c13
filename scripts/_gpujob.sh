timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "instance_norm" > gpurun_out/r2_t20.txt 2>&1; tail -2 gpurun_out/r2_t20.txt
python scripts/bench_kernels.py --only inorm --out gpurun_out/r2_k_in3.jsonl 2>&1 | cut -c1-75
for v in 1 2; do
python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err; echo "$(cut -c75-175 gpurun_out/r2_bench20.json)"
done
