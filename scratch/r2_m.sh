#!/bin/bash
# round 2, call M (4 GPUs): fitted gather slots, per-domain sub-batch defaults, the command line on several GPUs with only
# the used devices visible
set -x
mkdir -p gpurun_out
show() {
  python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(path))
    print(tag, 'N', d['n_gpus'], d['scaling'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'parity', (d['parity_sample'] or {}).get('identical'), (d['parity_sample'] or {}).get('ranks_identical'), 'gather', (d.get('gather') or {}).get('device_ms_last'), (d.get('gather') or {}).get('slot_records'), 'ranks', d['rank_ms_per_step_min_max'])
except Exception as e:
    print(tag, 'FAILED', e)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4"
timeout 600 $TR --master-port 29611 bench.py --gpus 4 --workload searchn --steps 5 --warmup 3 > gpurun_out/r2m_searchn_4gpu_weak.json 2> gpurun_out/r2m_searchn_4gpu_weak.log
show searchn_weak gpurun_out/r2m_searchn_4gpu_weak.json
timeout 600 $TR --master-port 29612 bench.py --gpus 4 --workload searchbs --steps 5 --warmup 3 > gpurun_out/r2m_searchbs_4gpu_weak.json 2> gpurun_out/r2m_searchbs_4gpu_weak.log
show searchbs_weak gpurun_out/r2m_searchbs_4gpu_weak.json
timeout 900 $TR --master-port 29613 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2m_searchp_4gpu_weak.json 2> gpurun_out/r2m_searchp_4gpu_weak.log
show searchp_weak gpurun_out/r2m_searchp_4gpu_weak.json
timeout 600 $TR --master-port 29614 bench.py --gpus 4 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2m_searchp_4gpu_strong.json 2> gpurun_out/r2m_searchp_4gpu_strong.log
show searchp_strong gpurun_out/r2m_searchp_4gpu_strong.json
timeout 900 python tools/cli_compare.py --workload searchp --reps 2 --gpus 4 > gpurun_out/r2m_cli_searchp_4gpu.json 2> gpurun_out/r2m_cli_searchp_4gpu.log
cat gpurun_out/r2m_cli_searchp_4gpu.json
