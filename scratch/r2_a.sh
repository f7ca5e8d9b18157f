#!/bin/bash
# round 2, call A: new unified DPX kernels (private profiles, residue-plane trace) -- tests, then the three workloads
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv; nproc
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2a_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2a_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2a_pytest.log | head -40
for wl in searchn searchbs searchp; do
  ( time timeout 700 python bench.py --workload $wl --steps 3 --warmup 3 ) > gpurun_out/r2a_bench_$wl.json 2> gpurun_out/r2a_bench_$wl.log
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2a_bench_$wl.json'))
    print('$wl', round(d['ms_per_step'],2), round(d['ms_per_step_serial_1_stream'],2), {k: round(v,2) for k,v in d['stage_ms'].items()}, d.get('parity_sample'), d['roofline']['frac'])
except Exception as e:
    print('$wl FAILED', e)
PY
  tail -3 gpurun_out/r2a_bench_$wl.log
done
# per-kernel times of one serial searchn step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a_launches_searchn.csv \
    python tools/profile_run.py searchn 1 > gpurun_out/r2a_ncu_searchn.log 2>&1
python - <<'PY'
import csv, collections
t = collections.Counter(); n = collections.Counter()
rows = [r for r in csv.reader(open('gpurun_out/r2a_launches_searchn.csv')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[1:]:
    try: t[r[ki].split('(')[0]] += float(r[vi].replace(',', '')); n[r[ki].split('(')[0]] += 1
    except ValueError: pass
for k, v in t.most_common(14): print(f'{v/1e6:9.2f} ms {n[k]:5d}x  {k}')
PY
