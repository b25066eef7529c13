timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "conv2d or conv_transpose" > gpurun_out/r2_t17.txt 2>&1; tail -2 gpurun_out/r2_t17.txt
python scripts/bench_kernels.py --only "tc_conv" --out gpurun_out/r2_k_conv.jsonl 2>&1 | cut -c1-110
for v in 1 2; do
python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err; echo "$(cut -c75-175 gpurun_out/r2_bench21.json)"
done
