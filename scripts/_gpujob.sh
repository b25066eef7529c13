timeout 1200 python -m pytest tests/test_step_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/r2_t9.txt 2>&1; tail -4 gpurun_out/r2_t9.txt
python bench.py --steps 10 --no-cpu-baseline --no-extra > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; cut -c1-250 gpurun_out/r2_bench6.json
