#!/bin/bash
TAG=${1:-r2v}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -12
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_${TAG}.txt | grep -A1 "profiled" | tail -12
timeout 500 python tools/breakdown.py 16 64 random > gpurun_out/breakdown_chi64_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_chi64_${TAG}.txt | tail -14
