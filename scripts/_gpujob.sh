timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "dwconv" > gpurun_out/r2_t13.txt 2>&1; tail -5 gpurun_out/r2_t13.txt
python scripts/bench_kernels.py --only dwconv --out gpurun_out/r2_k_dw2.jsonl 2>&1 | cut -c1-120
