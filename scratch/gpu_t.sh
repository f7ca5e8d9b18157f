#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
for m in ckpt planes; do
LAMBDA_B200_TRACE=$m timeout 300 python tools/bench_dp.py --queries 100000 --windows 1 --trace 2>&1 | tail -2
LAMBDA_B200_TRACE=$m timeout 300 python tools/bench_dp.py --queries 400000 --windows 1 --qlen 150 --wlen 176 --trace 2>&1 | tail -1
LAMBDA_B200_TRACE=$m timeout 300 python tools/bench_dp.py --queries 10000 --windows 1 --ragged --trace 2>&1 | tail -1
done
