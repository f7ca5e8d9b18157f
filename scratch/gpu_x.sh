#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_T.jsonl
for w in searchp searchp_real searchn searchbs; do
SWEEP_WORKLOAD=$w SWEEP_STEPS=3 python tools/sweep.py STREAMS=1 STREAMS=3 2>>gpurun_out/sweep.log | tee -a gpurun_out/sweep_T.jsonl
done
