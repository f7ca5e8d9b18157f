#!/bin/bash
# round 2, call B: device-resident record flow (finalise on device, deferred timers, export ABI)
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2b_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2b_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2b_pytest.log | head -40
for wl in searchp searchn searchbs searchp_real; do
  ( time timeout 700 python bench.py --workload $wl --steps 5 --warmup 3 ) > gpurun_out/r2b_bench_$wl.json 2> gpurun_out/r2b_bench_$wl.log
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2b_bench_$wl.json'))
    print('$wl', round(d['ms_per_step'],2), round(d['ms_per_step_serial_1_stream'],2), {k: round(v,2) for k,v in d['stage_ms'].items()}, d.get('parity_sample'), round(d['roofline']['frac'],3), round(d['roofline_trace']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],2))
except Exception as e:
    print('$wl FAILED', e)
PY
  tail -3 gpurun_out/r2b_bench_$wl.log
done
