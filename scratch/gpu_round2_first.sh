#!/bin/bash
# First GPU call of the next round (≈8 min of box time): everything written after the round-1 GPU budget ran out gets
# its first run, then the evidence the next kernel work needs.
#   gpurun --timeout 900 -- 'bash scratch/gpu_round2_first.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv; nproc
# 1. full GPU suite WITHOUT -x: tests/test_zz_gpu_option_fuzz.py (random non-default options, other BLAST modes, N under
#    random seeding options) has never run on a GPU; list every failure
( time timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1
grep -n "passed\|failed" gpurun_out/pytest_gpu.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/pytest_gpu.log | head -60
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
# 2. short-read workloads: where DP pass 2 dominates (searchn 65 of 146 ms, searchbs 23 of 73 ms)
for wl in searchn searchbs; do
  ( time timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 ) > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log
  python -c "import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', d['ms_per_step'], d['stage_ms'], d['parity_sample'])"
done
# 2b. never measured on the short-read workloads: the checkpoint trace path (0.7 B/cell instead of 3)
for wl in searchn searchbs; do
  LAMBDA_B200_TRACE=ckpt timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ckpt_$wl.log | \
    python -c "import sys, json; d=json.loads(sys.stdin.read()); print('$wl ckpt', d['ms_per_step'], d['stage_ms'])"
done
# 3. per-kernel times of one serial searchn step (fill vs traceback share of the trace stage)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_searchn.csv \
    python tools/profile_run.py searchn 1 > gpurun_out/ncu_searchn.log 2>&1
python - <<'PY'
import csv, collections
t = collections.Counter(); n = collections.Counter()
rows = [r for r in csv.reader(open('gpurun_out/launches_searchn.csv')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[1:]:
    try: t[r[ki].split('(')[0]] += float(r[vi].replace(',', '')); n[r[ki].split('(')[0]] += 1
    except ValueError: pass
for k, v in t.most_common(12): print(f'{v/1e6:9.2f} ms {n[k]:5d}x  {k}')
PY
# 4. headline bench last (index build 110 s)
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_searchp.json 2> gpurun_out/bench_searchp.log
cat gpurun_out/bench_searchp.json
