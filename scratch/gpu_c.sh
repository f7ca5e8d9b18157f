#!/bin/bash
# tests + three bench workloads + CLI wall clock + ncu captures
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_searchp.json 2> gpurun_out/bench_searchp.log
cat gpurun_out/bench_searchp.json
( time timeout 900 python tools/cli_compare.py ) > gpurun_out/cli_searchp.json 2> gpurun_out/cli_searchp.log
cat gpurun_out/cli_searchp.json; tail -3 gpurun_out/cli_searchp.log
( time timeout 900 python bench.py --workload searchp_real --steps 5 --warmup 3 ) > gpurun_out/bench_searchp_real.json 2> gpurun_out/bench_searchp_real.log
cat gpurun_out/bench_searchp_real.json
( time timeout 1200 python bench.py --workload searchn --steps 3 --warmup 3 ) > gpurun_out/bench_searchn.json 2> gpurun_out/bench_searchn.log
cat gpurun_out/bench_searchn.json; tail -3 gpurun_out/bench_searchn.log
( time timeout 1200 python bench.py --workload searchbs --steps 3 --warmup 3 ) > gpurun_out/bench_searchbs.json 2> gpurun_out/bench_searchbs.log
cat gpurun_out/bench_searchbs.json; tail -3 gpurun_out/bench_searchbs.log
