timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "multi_maxpool or maxpool" > gpurun_out/r2_t10.txt 2>&1; tail -15 gpurun_out/r2_t10.txt
timeout 1200 python -m pytest tests/test_step_gpu.py -x -q > gpurun_out/r2_t11.txt 2>&1; tail -4 gpurun_out/r2_t11.txt
python bench.py --detail --steps 10 --no-cpu-baseline --no-extra > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; cut -c1-250 gpurun_out/r2_bench7.json
grep "TF/s" gpurun_out/r2_bench7.err | grep -i "pool" | head
