#!/bin/bash
# round 2, call C: seeding (prefix table, text-mode elongation), zero-copy result buffer, BEST2 DP variant
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2c_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2c_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2c_pytest.log | head -40
show() {
  python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(path))
    print(tag, round(d['ms_per_step'],2), round(d['ms_per_step_serial_1_stream'],2), {k: round(v,2) for k,v in d['stage_ms'].items()}, (d.get('parity_sample') or {}).get('identical'), round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],2))
except Exception as e:
    print(tag, 'FAILED', e)
PY
}
for wl in searchn searchbs searchp; do
  timeout 700 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2c_bench_$wl.json 2> gpurun_out/r2c_bench_$wl.log
  show $wl gpurun_out/r2c_bench_$wl.json; tail -2 gpurun_out/r2c_bench_$wl.log
done
LAMBDA_B200_LIB=$PWD/lambda_b200/_build/lib_best0.so timeout 600 python bench.py --workload searchp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ab_best0.json 2> gpurun_out/r2c_ab_best0.log
show searchp_best0 gpurun_out/r2c_ab_best0.json
LAMBDA_B200_SEED_TEXT=0 timeout 600 python bench.py --workload searchn --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ab_notext.json 2> gpurun_out/r2c_ab_notext.log
show searchn_notext gpurun_out/r2c_ab_notext.json
LAMBDA_B200_SEED_PREFIX=0 timeout 600 python bench.py --workload searchn --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ab_noprefix.json 2> gpurun_out/r2c_ab_noprefix.log
show searchn_noprefix gpurun_out/r2c_ab_noprefix.json
LAMBDA_B200_SEED_PREFIX=0 timeout 600 python bench.py --workload searchp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ab_noprefix_p.json 2> gpurun_out/r2c_ab_noprefix_p.log
show searchp_noprefix gpurun_out/r2c_ab_noprefix_p.json
