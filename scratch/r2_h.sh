#!/bin/bash
# round 2, call H: two-mismatch seeds on the device, spec kernel at 8 blocks/SM, cold command-line marks + upload sweep
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2h_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2h_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2h_pytest.log | head -40
# racecheck over every seeding kernel (the selection of call G only matched three tests)
( time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "test_search_reproduces_reference_output and none and (prot_flat or nucl or bisulfite or blastx)" ) > gpurun_out/r2h_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2h_sanitizer_racecheck.log
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
  timeout 700 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2h_bench_$wl.json 2> gpurun_out/r2h_bench_$wl.log
  show $wl gpurun_out/r2h_bench_$wl.json
done
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from lambda_b200 import synth
W = bench.WORKLOADS["searchp"]
d = bench.ensure_index("searchp", W["n_seqs"])
q, qo = bench.make_queries("searchp", d, W["n_queries"], W["qlen"], seed=1000)
synth.write_fasta("/tmp/q_searchp.fasta", q, qo.astype(np.int64), "Q")
open("/tmp/searchp_dir", "w").write(d)
PY
D=$(cat /tmp/searchp_dir)
for cfg in "6 16" "12 16" "16 16" "12 64" "24 32" "12 16"; do
  set -- $cfg
  rm -f /tmp/o.m8
  { time LAMBDA_B200_TRACE_TIMES=1 LAMBDA_B200_UPLOAD_THREADS=$1 LAMBDA_B200_UPLOAD_CHUNK_MB=$2 bin/lambda3_b200 searchp -q /tmp/q_searchp.fasta -i $D/db.lba -o /tmp/o.m8 -v 2 ; } > gpurun_out/r2h_cli_t$1_c$2.log 2>&1
  echo "threads $1 chunk $2:"; grep "^real\|Runtime total\|GPU 0\|index:\|lba:\|context created\|search: start\|records on the host" gpurun_out/r2h_cli_t$1_c$2.log | head -24
done
