for v in 0 1 0 1; do
DSGAN_DW_MULTI=$v python bench.py --no-cpu-baseline --no-extra --steps 10 > gpurun_out/r2_bench_dm$v.json 2> gpurun_out/r2_bench_dm$v.err; echo "multi=$v $(cut -c75-175 gpurun_out/r2_bench_dm$v.json)"
done
python bench.py --detail --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r2_detail3.json 2> gpurun_out/r2_detail3.err
