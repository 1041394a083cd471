#!/bin/bash
# quick regression: GPU parity tests + per-phase breakdown of a saturated 16x16 chi=32 layer
TAG=${1:-q}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -6
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_${TAG}.txt | tail -14
