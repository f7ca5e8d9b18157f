#!/bin/bash
# round 2, call F: packed occurrence table, deeper prefix table
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2f_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2f_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2f_pytest.log | head -40
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
  timeout 700 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2f_bench_$wl.json 2> gpurun_out/r2f_bench_$wl.log
  show $wl gpurun_out/r2f_bench_$wl.json
  LAMBDA_B200_OCC_PACK=0 timeout 700 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_nopack_$wl.json 2> gpurun_out/r2f_nopack_$wl.log
  show ${wl}_nopack gpurun_out/r2f_nopack_$wl.json
done
timeout 900 python tools/cli_compare.py --workload searchp --reps 2 > gpurun_out/r2f_cli_searchp.json 2> gpurun_out/r2f_cli_searchp.log
cat gpurun_out/r2f_cli_searchp.json
