timeout 1200 python -m pytest tests/test_tc_gemm_gpu.py tests/test_kernels_gpu.py tests/test_step_gpu.py -x -q > gpurun_out/r2_t8.txt 2>&1; tail -4 gpurun_out/r2_t8.txt
python bench.py --detail --steps 10 --no-cpu-baseline --no-extra > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; cut -c1-250 gpurun_out/r2_bench5.json
grep "TF/s" gpurun_out/r2_bench5.err | grep "wgrad" | head -40
