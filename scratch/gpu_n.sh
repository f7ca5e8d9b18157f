#!/bin/bash
mkdir -p gpurun_out
( time timeout 170 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1
grep -n "passed\|failed" gpurun_out/pytest_gpu.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/pytest_gpu.log | head -40
