#!/bin/bash
# round 2, call K: sub-batches in flight per call, now that the host part of a step is small
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "test_score_kernel_all_length_classes or test_extension_matches_oracle" 2>&1 | tail -2
run() { # workload, tag, env...
  wl=$1; tag=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$wl $tag', 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'serial', round(d['ms_per_step_serial_1_stream'], 2))
open('gpurun_out/r2k_sweep.jsonl', 'a').write(json.dumps({'workload': '$wl', 'setting': '$tag', 'ms_per_step': d['ms_per_step'], 'e2e_ms_per_step': d['e2e']['ms_per_step'], 'serial_ms': d['ms_per_step_serial_1_stream']}) + '\n')"
}
rm -f gpurun_out/r2k_sweep.jsonl
for s in 1 2 3 4; do run searchp streams=$s LAMBDA_B200_STREAMS=$s; done
run searchp streams=3,occ=32 LAMBDA_B200_STREAMS=3 LAMBDA_B200_DPX_OCC=32
run searchp streams=2,occ=32 LAMBDA_B200_STREAMS=2 LAMBDA_B200_DPX_OCC=32
for s in 2 3 4 6; do run searchn streams=$s LAMBDA_B200_STREAMS=$s; done
for s in 2 3 4 6; do run searchbs streams=$s LAMBDA_B200_STREAMS=$s; done
run searchn streams=4,minsub=8192 LAMBDA_B200_STREAMS=4 LAMBDA_B200_MIN_SUBBATCH=8192
cat gpurun_out/r2k_sweep.jsonl | wc -l
