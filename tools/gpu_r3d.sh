#!/bin/bash
# message arena for clones + tail-aware fp64 Gram split: parity tests, bench (e2e), compute-sanitizer on the kernels changed this round
TAG=${1:-r3d}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
tail -2 gpurun_out/smoke_${TAG}.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_${TAG}.log 2>&1
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-300
tail -1 gpurun_out/bench_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e', d['e2e']['value'], 'value', d['value'], 'ms', d['ms_per_step'], 'bp', d['bp_sweep_ms'])"
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_chi.py -k "test_tensor_core_mode_products or (hub and 32 and complex64) or roundtrip" -q -x -p no:cacheprovider --timeout 650 > gpurun_out/sanitizer_memcheck_${TAG}.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck_${TAG}.log
tail -4 gpurun_out/sanitizer_memcheck_${TAG}.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -k "test_tensor_core_mode_products and 16" -q -x -p no:cacheprovider --timeout 550 > gpurun_out/sanitizer_racecheck_${TAG}.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_${TAG}.log
tail -4 gpurun_out/sanitizer_racecheck_${TAG}.log
