#!/bin/bash
# round 2, call E (8 GPUs): the BASELINE configs that need a whole box -- searchp strong / weak at 8, searchn at 4 and 8,
# searchbs at 8, the command line at 1 and 8 GPUs against lambda3 -t $(nproc)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
show() {
  python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(path))
    print(tag, 'N', d['n_gpus'], d['scaling'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'parity', d['parity_sample'], 'gather', (d.get('gather') or {}).get('device_ms_last'), 'ranks', d['rank_ms_per_step_min_max'])
except Exception as e:
    print(tag, 'FAILED', e)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# command line first (builds the searchp index: ~100 s of the reference's mkindexp)
LAMBDA_B200_TRACE_TIMES=1 timeout 900 python tools/cli_compare.py --workload searchp --reps 2 --gpus 1 > gpurun_out/r2e_cli_searchp_1gpu.json 2> gpurun_out/r2e_cli_searchp_1gpu.log
cat gpurun_out/r2e_cli_searchp_1gpu.json; grep "lgpu" gpurun_out/r2e_cli_searchp_1gpu.log | tail -40
timeout 900 python tools/cli_compare.py --workload searchp --reps 2 --gpus 8 > gpurun_out/r2e_cli_searchp_8gpu.json 2> gpurun_out/r2e_cli_searchp_8gpu.log
cat gpurun_out/r2e_cli_searchp_8gpu.json
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2e_searchp_8gpu_strong.json 2> gpurun_out/r2e_searchp_8gpu_strong.log
show searchp_strong gpurun_out/r2e_searchp_8gpu_strong.json
timeout 600 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2e_searchp_8gpu_weak.json 2> gpurun_out/r2e_searchp_8gpu_weak.log
show searchp_weak gpurun_out/r2e_searchp_8gpu_weak.json
for n in 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --workload searchn --steps 5 --warmup 3 > gpurun_out/r2e_searchn_${n}gpu_weak.json 2> gpurun_out/r2e_searchn_${n}gpu_weak.log
  show searchn_$n gpurun_out/r2e_searchn_${n}gpu_weak.json
done
timeout 600 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --workload searchn --steps 5 --warmup 3 --scaling strong > gpurun_out/r2e_searchn_8gpu_strong.json 2> gpurun_out/r2e_searchn_8gpu_strong.log
show searchn_8_strong gpurun_out/r2e_searchn_8gpu_strong.json
timeout 600 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --workload searchbs --steps 5 --warmup 3 > gpurun_out/r2e_searchbs_8gpu_weak.json 2> gpurun_out/r2e_searchbs_8gpu_weak.log
show searchbs_8 gpurun_out/r2e_searchbs_8gpu_weak.json
tail -5 gpurun_out/r2e_searchbs_8gpu_weak.log
