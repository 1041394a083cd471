#!/bin/bash
# quick GPU check: parity tests + per-phase breakdown (+ optional bench)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --tb=short -x -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -15 gpurun_out/pytest_gpu_${TAG}.log
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
cat gpurun_out/breakdown_${TAG}.txt | tail -30
if [ -n "$2" ]; then timeout 600 python bench.py --no-cpu > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log | cut -c1-300; fi
