#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
for m in ckpt; do
LAMBDA_B200_TRACE=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_$m.csv python tools/bench_dp.py --queries 100000 --windows 1 --trace --reps 2 2>&1 | grep ms_extend
LAMBDA_B200_TRACE=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_rag_$m.csv python tools/bench_dp.py --queries 10000 --windows 1 --ragged --trace --reps 2 2>&1 | grep ms_extend
done
: > gpurun_out/sweep_trace.jsonl
for w in searchp searchp_real searchn; do
SWEEP_WORKLOAD=$w SWEEP_STEPS=3 python tools/sweep.py TRACE=ckpt,STREAMS=1 TRACE=ckpt,STREAMS=3 2>>gpurun_out/sweep.log | tee -a gpurun_out/sweep_trace.jsonl
done
