#!/bin/bash
# round 2, call J: protein trace classes with two alignments per warp, pre-formatted output lines, upload via cudaHostRegister
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2j_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2j_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2j_pytest.log | head -40
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
for wl in searchp searchp_real; do
  timeout 700 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2j_bench_$wl.json 2> gpurun_out/r2j_bench_$wl.log
  show $wl gpurun_out/r2j_bench_$wl.json
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
for mode in copy register copy register; do
  rm -f /tmp/o.m8
  { time LAMBDA_B200_TRACE_TIMES=1 LAMBDA_B200_UPLOAD_MODE=$mode bin/lambda3_b200 searchp -q /tmp/q_searchp.fasta -i $D/db.lba -o /tmp/o.m8 -v 2 ; } > gpurun_out/r2j_cli_$mode.log 2>&1
  echo "upload mode $mode:"; grep "^real\|Runtime total\|GPU 0\|context ready\|device memory allocated\|main blobs" gpurun_out/r2j_cli_$mode.log | tr '\n' ' '; echo
done
LAMBDA_B200_UPLOAD_MODE=register LAMBDA_B200_UPLOAD_CHUNK_MB=128 LAMBDA_B200_UPLOAD_THREADS=4 bash -c 'rm -f /tmp/o.m8; { time LAMBDA_B200_TRACE_TIMES=1 bin/lambda3_b200 searchp -q /tmp/q_searchp.fasta -i '$D'/db.lba -o /tmp/o.m8 -v 2 ; } 2>&1 | grep "^real\|GPU 0\|context ready\|main blobs" | tr "\n" " "'; echo
md5sum /tmp/o.m8
