#!/bin/bash
# round-2 call A: state of parity at the benchmarked shapes with the round-1 kernels, sanitizer logs, TMA / peak probes, bench line
TAG=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
nproc >> gpurun_out/gpu_${TAG}.txt; free -g | head -2 >> gpurun_out/gpu_${TAG}.txt
timeout 120 ./tools/tma_probe > gpurun_out/tma_probe_${TAG}.txt 2>&1
echo "probe exit $?" >> gpurun_out/tma_probe_${TAG}.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s --timeout 400 > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "PARITY|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-400 | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
tail -3 gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.log 2>&1
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-600
SAN="tests/test_gpu_parity.py -k test_tensor_core_mode_products"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest $SAN -q -x -p no:cacheprovider --timeout 800 > gpurun_out/sanitizer_memcheck_${TAG}.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck_${TAG}.log
tail -5 gpurun_out/sanitizer_memcheck_${TAG}.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest "tests/test_gpu_parity.py" -k "test_tensor_core_mode_products and 8" -q -x -p no:cacheprovider --timeout 800 > gpurun_out/sanitizer_racecheck_${TAG}.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_${TAG}.log
tail -5 gpurun_out/sanitizer_racecheck_${TAG}.log
