#!/bin/bash
# tc2 mode product with the decoupled lo ring: stand-alone correctness + bandwidth, then parity tests and breakdowns
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 150 ./tools/tc2_test bench > gpurun_out/tc2_test_${TAG}.txt 2>&1
echo "tc2_test exit $?" >> gpurun_out/tc2_test_${TAG}.txt
grep -E "FAIL|ELIG|batch|failed|exit" gpurun_out/tc2_test_${TAG}.txt | cut -c1-120
grep -E "^bench" gpurun_out/tc2_test_${TAG}.txt | sed 's/.*\[/[/' | cut -c1-160
if grep -q "FAIL" gpurun_out/tc2_test_${TAG}.txt; then echo "stand-alone test failed: skipping the rest"; exit 1; fi
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -8
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_${TAG}.txt | grep -A1 "profiled" | tail -12
timeout 500 python tools/breakdown.py 16 64 random > gpurun_out/breakdown_chi64_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_chi64_${TAG}.txt | tail -12
