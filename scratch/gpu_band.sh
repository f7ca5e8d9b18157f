#!/bin/bash
# final check of the round: parity tests, smoke, band sweep on the real-length workload (BASELINE configs[3])
set -x
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
: > gpurun_out/sweep_band.jsonl
for band in 0 16 32 64; do
  timeout 400 python bench.py --workload searchp_real --band $band --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/band_$band.log | tee gpurun_out/bench_searchp_real_band$band.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'band': $band, 'ms_per_step': d['ms_per_step'], 'queries_per_s': d['value'], 'e2e_queries_per_s': d['e2e']['value'], 'gcups': d['gcups'], 'gcups_score_kernel': d['gcups_score_kernel'], 'roofline_frac': d['roofline']['frac'], 'stage_ms': d['stage_ms'], 'funnel': d['funnel'], 'clocks': d['clocks']}))" | tee -a gpurun_out/sweep_band.jsonl
done
