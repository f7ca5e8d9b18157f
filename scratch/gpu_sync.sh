#!/bin/bash
# blocking vs spinning host waits: 1 GPU sweep, then N GPUs through torchrun
mkdir -p gpurun_out
N=${N:-2}
SWEEP_STEPS=4 python tools/sweep.py BLOCKING_SYNC=1 BLOCKING_SYNC=0 2>>gpurun_out/sweep.log | tee gpurun_out/sweep_sync.jsonl
for b in 1 0; do
LAMBDA_B200_BLOCKING_SYNC=$b python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_${N}gpu_sync$b.json 2> gpurun_out/bench_${N}gpu_sync$b.log
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_sync$b.json')); print('N=$N blocking=$b', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
