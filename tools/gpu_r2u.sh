#!/bin/bash
TAG=${1:-r2u}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -12
timeout 600 python bench.py > gpurun_out/bench_${TAG}.log 2>&1
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-300
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_${TAG}.txt | tail -14
TNQS_SLOWLOG=1 timeout 400 python bench.py --chi 64 --random-state --steps 2 --warmup 1 --no-cpu --inplace > gpurun_out/bench_16x16_chi64_1gpu_${TAG}.log 2>&1
grep "tnqs slow" gpurun_out/bench_16x16_chi64_1gpu_${TAG}.log | cut -c1-200 | tail -12
tail -1 gpurun_out/bench_16x16_chi64_1gpu_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['roofline']['families']
print('chi64 gates/s %.1f ms/layer %.1f steps %s bp_sweep_ms %.2f | mode %.1f ms %.0f GB/s | gram %.1f ms %.0f GB/s | small %.1f ms' % (d['value'], d['ms_per_step'], d['config']['step_ms'], d['bp_sweep_ms'], f['mode_product']['ms_one_layer'], f['mode_product']['GBps'], f['gram']['ms_one_layer'], f['gram']['GBps'], f['jacobi_cholesky_small']['ms_one_layer']))"
