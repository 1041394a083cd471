#!/bin/bash
# out-of-memory recovery (emergency trim) on BASELINE config 4 (58 GB state + functional copy), then the parity tests
TAG=${1:-r3g}
mkdir -p gpurun_out
TNQS_SLOWLOG=1 timeout 600 python bench.py --workload cubic3d --chi 16 --random-state --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_cfg4_${TAG}.log 2>&1
grep "tnqs slow" gpurun_out/bench_cfg4_${TAG}.log | cut -c1-60 | sort | uniq -c | sort -rn | head -8
tail -1 gpurun_out/bench_cfg4_${TAG}.log | cut -c1-250
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -8
