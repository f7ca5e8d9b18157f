#!/bin/bash
# round 2, last call: the committed state end to end -- GPU tests, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2z_pytest.log 2>&1; tail -4 gpurun_out/r2z_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2z_bench.json'))
print('bench', round(d['value']), 'q/s', round(d['ms_per_step'], 2), 'ms; e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), 'parity', d['parity_sample'], 'frac', round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'], d['roofline']['traffic_source'], 'launches', d['gpu_launches'], d['clocks'])
PY
