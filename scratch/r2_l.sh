#!/bin/bash
# round 2, call L: final single-GPU evidence -- tests, all bench lines, ncu --set full of one step per workload, command line
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=30 ) > gpurun_out/r2l_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2l_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r2l_pytest.log | head -40
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
show() {
  python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(path))
    print(tag, round(d['ms_per_step'],2), round(d['ms_per_step_serial_1_stream'],2), {k: round(v,2) for k,v in d['stage_ms'].items()}, (d.get('parity_sample') or {}).get('identical'), round(d['roofline']['frac'],3), round(d['roofline_trace']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],2), 'q/s', round(d['value']), round(d['e2e']['value']), 'cpu', round(d.get('cpu_baseline',{}).get('value',0)))
except Exception as e:
    print(tag, 'FAILED', e)
PY
}
# ncu first: the summaries land in profiles/ and the bench lines below pick them up (same kernel sources)
for wl in searchp searchn searchbs; do
  timeout 900 ncu --set full --clock-control none -k regex:"swDpx|traceback|seedSpec|seedBlock|classify|postTrace|widen|chain|markDup|rankCut|gatherFinal|prepQueries|filterKernel" -f -o /tmp/r2l_full_$wl python tools/profile_run.py $wl 1 > gpurun_out/r2l_ncu_$wl.log 2>&1
  tail -1 gpurun_out/r2l_ncu_$wl.log
  python tools/ncu_summary.py /tmp/r2l_full_$wl.ncu-rep $wl > gpurun_out/r2l_ncu_summary_$wl.txt 2>&1; head -12 gpurun_out/r2l_ncu_summary_$wl.txt
  cp profiles/r2_ncu_kernels_$wl.json gpurun_out/
  ncu -i /tmp/r2l_full_$wl.ncu-rep --page raw --csv > gpurun_out/r2l_ncu_raw_$wl.csv 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2l_launches_searchp.csv python tools/profile_run.py searchp 2 > /dev/null 2>&1
for wl in searchp searchn searchbs searchp_real searchp_small; do
  timeout 700 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2l_bench_$wl.json 2> gpurun_out/r2l_bench_$wl.log
  show $wl gpurun_out/r2l_bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_bench_reference.json 2> gpurun_out/r2l_bench_reference.log; cat gpurun_out/r2l_bench_reference.json | cut -c1-400
timeout 900 python tools/cli_compare.py --workload searchp --reps 3 > gpurun_out/r2l_cli_searchp.json 2> gpurun_out/r2l_cli_searchp.log
cat gpurun_out/r2l_cli_searchp.json
