#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep_trace.jsonl
for w in searchp searchp_real searchn; do
SWEEP_WORKLOAD=$w SWEEP_STEPS=3 python tools/sweep.py TRACE=ckpt,STREAMS=1 TRACE=planes,STREAMS=1 TRACE=ckpt,STREAMS=3 TRACE=planes,STREAMS=3 2>>gpurun_out/sweep.log | tee -a gpurun_out/sweep_trace.jsonl
done
