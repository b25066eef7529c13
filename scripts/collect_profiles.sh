#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): writes everything into gpurun_out/final_*; copy what is judged to profiles/.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/final_gputests.txt 2>&1; tail -3 $O/final_gputests.txt
python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err; cut -c1-200 $O/final_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference_arm.json 2> $O/final_ref.err; cut -c1-200 $O/final_bench_reference_arm.json
python scripts/bench_kernels.py --out $O/final_kernel_rooflines.jsonl > $O/final_kernels.log 2>&1; tail -2 $O/final_kernels.log
python tests/tools/error_budget.py --batch 16 --top 8 --out $O/final_error_budget_gpu.json > $O/final_budget.log 2>&1; tail -2 $O/final_budget.log
python bench.py --detail --steps 5 --no-cpu-baseline --no-extra > $O/final_bench_detail.json 2> $O/final_bench_detail.err
# same box A/B of this round's kernel families (environment switches keep the round-1 paths reachable)
for sw in "DSGAN_DW_MMA=0" "DSGAN_NM_CONV=0" "DSGAN_DW_MULTI=0" "X=1"; do
  env $sw python bench.py --no-cpu-baseline --no-extra --steps 10 > $O/final_ab.json 2>/dev/null
  echo "$sw $(python -c "import json;d=json.load(open('$O/final_ab.json'));print(round(d['value'],1),'img/s',round(d['ms_per_step'],3),'ms')")" >> $O/final_ab.txt
done
cat $O/final_ab.txt
# launch list (same command plain first, then under ncu; eager launches so that the list is a plain kernel sequence)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-graph > $O/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 900 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-graph > $O/final_ncu.log 2>&1
tail -1 $O/final_ncu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.txt 2>&1; tail -1 $O/final_smoke.txt
