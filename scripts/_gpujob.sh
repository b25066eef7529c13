for d in 79 207 143 128; do echo "DBG=$d"; DSGAN_MLP_DBG=$d python scripts/bench_kernels.py --only mlp --out gpurun_out/r2_k_dbg.jsonl 2>&1 | grep "bwd uc4\|bwd uc3" | cut -c1-130; done
