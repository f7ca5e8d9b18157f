#!/bin/bash
# parity tests (new CLI tests included), then the trace-plane store variants on the short-read workloads
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_planestore.jsonl
for wl in searchn searchbs; do
for mode in default cs; do
  LAMBDA_B200_PLANE_STORE=$mode timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ps_${wl}_$mode.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'workload': '$wl', 'plane_store': '$mode', 'ms_per_step': d['ms_per_step'], 'serial_ms': d['ms_per_step_serial_1_stream'], 'stage_ms': d['stage_ms'], 'e2e_ms': d['e2e']['ms_per_step'], 'clocks': d['clocks']}))" | tee -a gpurun_out/sweep_planestore.jsonl
done
done
