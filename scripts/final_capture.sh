#!/bin/bash
# Round-end evidence run (one GPU): tests, smoke, launch list, full ncu captures, bench.
set -x
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --batch 131072 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench_r1c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout_tc_kernel|loss_tc_kernel|wgrad_tc_kernel|target_tc_kernel|target_bwd_tc_kernel" -s 5 -c 5 -f -o gpurun_out/prof_all_r1c python bench.py --batch 65536 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_all_r1c.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
tail -c 600 gpurun_out/bench_r1c.json
