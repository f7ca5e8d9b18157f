#!/bin/bash
# validation of the host-wait policy + NVML clock sampling: tests, headline bench, 2-core oversubscription A/B
set -x
mkdir -p gpurun_out
nproc
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_searchp.json 2> gpurun_out/bench_searchp.log
cat gpurun_out/bench_searchp.json
: > gpurun_out/sweep_sync2.jsonl
for mode in spin yield; do
  LAMBDA_B200_SYNC=$mode timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/sync_$mode.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'cores': 'all', 'mode': '$mode', 'ms_per_step': d['ms_per_step'], 'wall_ms_per_step': d['wall_ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'clocks': d['clocks']}))" | tee -a gpurun_out/sweep_sync2.jsonl
done
for mode in spin auto block; do
  LAMBDA_B200_SYNC=$mode timeout 300 taskset -c 0-1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/sync2c_$mode.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'cores': 2, 'mode': '$mode', 'ms_per_step': d['ms_per_step'], 'wall_ms_per_step': d['wall_ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'clocks': d['clocks']}))" | tee -a gpurun_out/sweep_sync2.jsonl
done
