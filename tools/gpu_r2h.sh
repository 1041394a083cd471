#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -12
for chi in 8 16 32 64; do
  timeout 200 python bench.py --L 10 --chi $chi --random-state --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_cfg5_chi${chi}_${TAG}.log 2>&1
  tail -1 gpurun_out/bench_cfg5_chi${chi}_${TAG}.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); f=d['roofline']['families']
    print('cfg5 chi', $chi, 'gates/s %.1f ms/layer %.1f bp_sweep_ms %.2f sweeps/layer %.1f | mode %.1f ms %.0f GB/s | gram %.1f ms %.0f GB/s | small %.1f ms | launches %d' % (d['value'], d['ms_per_step'], d['bp_sweep_ms'], d['config']['bp_sweeps_per_layer'], f['mode_product']['ms_one_layer'], f['mode_product']['GBps'], f['gram']['ms_one_layer'], f['gram']['GBps'], f['jacobi_cholesky_small']['ms_one_layer'], d['gpu_launches']))
except Exception as e: print('cfg5 chi', $chi, 'FAILED', e)
"
done
timeout 300 python bench.py --workload heavyhex --chi 64 --prep 12 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_cfg3_${TAG}.log 2>&1
tail -1 gpurun_out/bench_cfg3_${TAG}.log | cut -c1-700
