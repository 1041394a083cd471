#!/bin/bash
# LOCALP Jacobi (one barrier per round for under-filled launches): parity tests, then the small-batch regime with and without it
TAG=${1:-r3p}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -8
for lp in 0 1; do
  echo "== TNQS_JACOBI_LOCALP=$lp (6x6, chi=32, 18 gates per colour)" >> gpurun_out/jacobi_localp_${TAG}.txt
  TNQS_JACOBI_LOCALP=$lp timeout 200 python tools/breakdown.py 6 32 random 2>&1 | grep -v "BP sweep after" | grep "SU colour" >> gpurun_out/jacobi_localp_${TAG}.txt
done
echo "== default choice, 16x16 chi=32 (128 gates per colour)" >> gpurun_out/jacobi_localp_${TAG}.txt
timeout 200 python tools/breakdown.py 2>&1 | grep -v "BP sweep after" | grep "SU colour\|layer" >> gpurun_out/jacobi_localp_${TAG}.txt
cut -c1-190 gpurun_out/jacobi_localp_${TAG}.txt
