#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}; nvidia-smi -L
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 ) > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.log
tail -5 gpurun_out/bench_${N}gpu.log
cat gpurun_out/bench_${N}gpu.json
