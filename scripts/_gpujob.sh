timeout 900 python -m pytest tests/test_fused_mlp_gpu.py -q -x > gpurun_out/r2_t15.txt 2>&1; tail -3 gpurun_out/r2_t15.txt
timeout 900 python -m pytest tests/test_step_gpu.py -q -x > gpurun_out/r2_t16.txt 2>&1; tail -3 gpurun_out/r2_t16.txt
python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; cut -c75-175 gpurun_out/r2_bench10.json
