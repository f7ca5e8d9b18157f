#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_searchp.json 2> gpurun_out/bench_searchp.log
cat gpurun_out/bench_searchp.json
( time timeout 900 python tools/cli_compare.py ) > gpurun_out/cli_searchp.json 2> gpurun_out/cli_searchp.log
cat gpurun_out/cli_searchp.json; tail -3 gpurun_out/cli_searchp.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_searchp.csv python tools/profile_run.py searchp 2 > gpurun_out/ncu_list.log 2>&1
K='regex:seedSpecKernel|seedBlockKernel|swScoreDpxKernel|swTraceDpxKernel|tracebackDpxKernel'
timeout 1200 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 --launch-count 12 -o gpurun_out/r1_searchp_full -f python tools/profile_run.py searchp 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
