timeout 900 python -m pytest tests/test_losses_gpu.py -x -q > gpurun_out/r2_t12.txt 2>&1; tail -2 gpurun_out/r2_t12.txt
python scripts/bench_kernels.py --only ssim --out gpurun_out/r2_k_ssim.jsonl 2>&1 | cut -c1-130
