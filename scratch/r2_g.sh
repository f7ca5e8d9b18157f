#!/bin/bash
# round 2, call G: cold command-line time marks + upload sweep, compute-sanitizer, ncu --set full of one step, spec occupancy A/B
set -x
mkdir -p gpurun_out
# --- 1. command line, cold: where does the time go ---
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from lambda_b200 import synth
W = bench.WORKLOADS["searchp"]
d = bench.ensure_index("searchp", W["n_seqs"])
q, qo = bench.make_queries("searchp", d, W["n_queries"], W["qlen"], seed=1000)
synth.write_fasta("/tmp/q_searchp.fasta", q, qo.astype(np.int64), "Q")
print(d)
open("/tmp/searchp_dir", "w").write(d)
PY
D=$(cat /tmp/searchp_dir)
for cfg in "6 16" "12 16" "16 16" "12 64" "24 32"; do
  set -- $cfg
  rm -f /tmp/o.m8
  ( LAMBDA_B200_TRACE_TIMES=1 LAMBDA_B200_UPLOAD_THREADS=$1 LAMBDA_B200_UPLOAD_CHUNK_MB=$2 /usr/bin/time -f "wall %e s" bin/lambda3_b200 searchp -q /tmp/q_searchp.fasta -i $D/db.lba -o /tmp/o.m8 -v 2 ) > gpurun_out/r2g_cli_t$1_c$2.log 2>&1
  echo "threads $1 chunk $2:"; grep "wall\|Runtime total\|GPU 0\|index:\|lba:\|context created\|search: start\|records on the host" gpurun_out/r2g_cli_t$1_c$2.log | head -24
done
# --- 2. compute-sanitizer over the golden searches (all seeding kernels, both DP passes, finalisation) ---
SEL='test_search_reproduces_reference_output and (prot_flat or nucl or bisulfite) and none'
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$SEL" ) > gpurun_out/r2g_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2g_sanitizer_memcheck.log
( time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$SEL and (auto or block)" ) > gpurun_out/r2g_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2g_sanitizer_racecheck.log
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "test_extension_matches_oracle and (prot_flat or nucl) or test_device_records_export" ) > gpurun_out/r2g_sanitizer_memcheck_ext.log 2>&1; echo "memcheck ext rc=$?"; tail -4 gpurun_out/r2g_sanitizer_memcheck_ext.log
# --- 3. ncu --set full of one serial step, our kernels ---
for wl in searchp searchn; do
  timeout 900 ncu --set full --clock-control none -k regex:"swDpx|traceback|seedSpec|seedBlock|classify|postTrace|widen|chain" -f -o /tmp/r2g_full_$wl python tools/profile_run.py $wl 1 > gpurun_out/r2g_ncu_$wl.log 2>&1
  tail -2 gpurun_out/r2g_ncu_$wl.log
  python tools/ncu_summary.py /tmp/r2g_full_$wl.ncu-rep $wl > gpurun_out/r2g_ncu_summary_$wl.txt 2>&1; cat gpurun_out/r2g_ncu_summary_$wl.txt
  cp profiles/r2_ncu_kernels_$wl.json gpurun_out/ 2>/dev/null
  ncu -i /tmp/r2g_full_$wl.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_raw_$wl.csv 2>/dev/null
  ls -la /tmp/r2g_full_$wl.ncu-rep
done
# --- 4. spec kernel occupancy ---
for mb in 6 8; do for wl in searchbs searchn; do
  LAMBDA_B200_LIB=$PWD/lambda_b200/_build/lib_spec$mb.so timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spec$mb $wl', round(d['ms_per_step'],2), round(d['stage_ms']['ms_seed'],2))"
done; done
