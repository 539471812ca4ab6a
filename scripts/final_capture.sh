#!/bin/bash
# Round-end evidence run (one GPU): tests, smoke, launch list, full ncu captures, parity report, bench.  TAG names the outputs.
set -x
TAG=${TAG:-r2f}
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --batch 151552 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench_$TAG.log 2>&1
for k in rollout_h_kernel loss_h_kernel wgrad_h_kernel target_h_kernel target_bwd_h_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --batch 75776 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
timeout 900 python scripts/parity_report.py --both > gpurun_out/parity_report_$TAG.txt 2>/dev/null
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
