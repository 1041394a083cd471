#!/bin/bash
# last call of the round: parity tests + smoke on the final build
TAG=${1:-r3m}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
tail -2 gpurun_out/smoke_${TAG}.log | cut -c1-300
