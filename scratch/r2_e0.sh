#!/bin/bash
# round 2, call E0 (2 GPUs): the multi-GPU path end to end at small cost -- weak and strong scaling, device gather, per-rank parity
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "int16 or long_self or export" > gpurun_out/r2e0_pytest.log 2>&1; tail -3 gpurun_out/r2e0_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for mode in weak strong; do
  timeout 600 $TR bench.py --gpus 2 --workload searchn --steps 3 --warmup 3 --scaling $mode > gpurun_out/r2e0_searchn_2gpu_$mode.json 2> gpurun_out/r2e0_searchn_2gpu_$mode.log
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2e0_searchn_2gpu_$mode.json'))
    print('$mode', d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), d['scaling'], d['parity_sample'], d['gather'], d['rank_ms_per_step_min_max'], round(d['e2e']['value']))
except Exception as e:
    print('$mode FAILED', e)
PY
  tail -4 gpurun_out/r2e0_searchn_2gpu_$mode.log
done
timeout 600 python bench.py --workload searchn --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e0_searchn_1gpu.json 2> gpurun_out/r2e0_searchn_1gpu.log
python -c "import json; d=json.load(open('gpurun_out/r2e0_searchn_1gpu.json')); print('1gpu', round(d['value']), round(d['ms_per_step'],2))"
