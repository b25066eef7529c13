timeout 600 python -m pytest tests/test_fused_mlp_gpu.py -x -q > gpurun_out/r2_t_mlp.txt 2>&1; tail -15 gpurun_out/r2_t_mlp.txt
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_fused_mlp_gpu.py > gpurun_out/r2_t1.txt 2>&1; tail -5 gpurun_out/r2_t1.txt
timeout 600 python tests/tools/error_budget.py --batch 16 --configs bbb,bfb --top 12 > gpurun_out/r2_budget2.log 2>&1; grep -v "Warn\|Consider\|losses =" gpurun_out/r2_budget2.log | tail -60
