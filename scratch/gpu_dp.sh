#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/bench_dp.jsonl
for f in 0 1 2; do
  LAMBDA_B200_LIB=$PWD/lambda_b200/_build/lib_form$f.so timeout 300 python tools/bench_dp.py --queries 100000 --windows 16 2>&1 | tee -a gpurun_out/bench_dp.jsonl
done
