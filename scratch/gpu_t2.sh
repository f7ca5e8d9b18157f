#!/bin/bash
mkdir -p gpurun_out
for m in ckpt planes; do
LAMBDA_B200_TRACE=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_$m.csv python tools/bench_dp.py --queries 100000 --windows 1 --trace --reps 2 > /dev/null 2>&1
LAMBDA_B200_TRACE=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_rag_$m.csv python tools/bench_dp.py --queries 10000 --windows 1 --ragged --trace --reps 2 > /dev/null 2>&1
done
