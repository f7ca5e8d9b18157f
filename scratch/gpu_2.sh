#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 ) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.log
tail -5 gpurun_out/bench_2gpu.log
cat gpurun_out/bench_2gpu.json
( time python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.log
cat gpurun_out/bench_1gpu.json
