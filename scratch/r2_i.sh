#!/bin/bash
# round 2, call I: per-class launches on side streams; all workloads; H2D bandwidth of the box
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2i_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2i_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2i_pytest.log | head -40
show() {
  python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(path))
    print(tag, round(d['ms_per_step'],2), round(d['ms_per_step_serial_1_stream'],2), {k: round(v,2) for k,v in d['stage_ms'].items()}, (d.get('parity_sample') or {}).get('identical'), round(d['roofline']['frac'],3), round(d['roofline_trace']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],2))
except Exception as e:
    print(tag, 'FAILED', e)
PY
}
for wl in searchp_real searchp searchn searchbs searchp_small; do
  timeout 700 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2i_bench_$wl.json 2> gpurun_out/r2i_bench_$wl.log
  show $wl gpurun_out/r2i_bench_$wl.json
done
LAMBDA_B200_CLASS_STREAMS=0 timeout 700 python bench.py --workload searchp_real --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_nostreams_searchp_real.json 2> /dev/null
show searchp_real_nostreams gpurun_out/r2i_nostreams_searchp_real.json
python - <<'PY'
import torch, time
a = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    b.copy_(a, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    b.copy_(a, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("H2D pinned 1 GiB x4: %.1f GB/s" % (4 * (1 << 30) / dt / 1e9))
t0 = time.perf_counter()
for _ in range(4):
    a.copy_(b, non_blocking=True)
torch.cuda.synchronize()
print("D2H pinned 1 GiB x4: %.1f GB/s" % (4 * (1 << 30) / (time.perf_counter() - t0) / 1e9))
PY
