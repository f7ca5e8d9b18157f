#!/bin/bash
# round 2, call N (8 GPUs, the rest of the budget): the short-read workloads at 8 GPUs with the final code
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 95 $TR --master-port 29711 bench.py --gpus 8 --workload searchbs --steps 5 --warmup 3 --parity-queries 2000 > gpurun_out/r2n_searchbs_8gpu_weak.json 2> gpurun_out/r2n_searchbs_8gpu_weak.log
timeout 100 $TR --master-port 29712 bench.py --gpus 8 --workload searchn --steps 5 --warmup 3 --parity-queries 2000 > gpurun_out/r2n_searchn_8gpu_weak.json 2> gpurun_out/r2n_searchn_8gpu_weak.log
python - <<'PY'
import json
for wl in ("searchbs", "searchn"):
    try:
        d = json.load(open(f"gpurun_out/r2n_{wl}_8gpu_weak.json"))
        print(wl, 'N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'parity', d['parity_sample'], 'gather', d['gather'], 'ranks', d['rank_ms_per_step_min_max'])
    except Exception as e:
        print(wl, 'FAILED', e)
PY
