timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "dwconv" > gpurun_out/r2_t19.txt 2>&1; tail -2 gpurun_out/r2_t19.txt
python scripts/bench_kernels.py --only dwconv --out gpurun_out/r2_k_dw4.jsonl 2>&1 | cut -c1-90
for v in 1 2; do
python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; echo "$(cut -c75-175 gpurun_out/r2_bench17.json)"
done
