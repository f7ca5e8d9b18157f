#!/bin/bash
set -x
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_searchp.json 2> gpurun_out/bench_searchp.log
cat gpurun_out/bench_searchp.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log
cat gpurun_out/bench_ref.json
for w in searchp_real searchn searchbs; do
( time timeout 1200 python bench.py --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.log
cat gpurun_out/bench_$w.json
done
