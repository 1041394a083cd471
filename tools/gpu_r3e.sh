#!/bin/bash
# final single-GPU evidence of round 2: the other BASELINE configs, the reference's default BP schedule, the reference arm
TAG=${1:-r3e}
mkdir -p gpurun_out
timeout 400 python bench.py --schedule forest --extras --no-cpu > gpurun_out/bench_forest_${TAG}.log 2>&1
tail -1 gpurun_out/bench_forest_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('forest schedule: gates/s %.1f ms/layer %.1f bp_sweep_ms %.2f sweeps/layer %.1f extras %s' % (d['value'], d['ms_per_step'], d['bp_sweep_ms'], d['config']['bp_sweeps_per_layer'], d['extras']))"
for chi in 8 16 32 64; do
  timeout 200 python bench.py --L 10 --chi $chi --random-state --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_cfg5_chi${chi}_${TAG}.log 2>&1
  tail -1 gpurun_out/bench_cfg5_chi${chi}_${TAG}.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); f=d['roofline']['families']
    print('cfg5 chi', $chi, 'gates/s %.1f ms/layer %.1f bp_sweep_ms %.2f | mode %.1f ms %.0f GB/s | gram %.1f ms %.0f GB/s | small %.1f ms | launches %d' % (d['value'], d['ms_per_step'], d['bp_sweep_ms'], f['mode_product']['ms_one_layer'], f['mode_product']['GBps'], f['gram']['ms_one_layer'], f['gram']['GBps'], f['jacobi_cholesky_small']['ms_one_layer'], d['gpu_launches']))
except Exception as e: print('cfg5 chi', $chi, 'FAILED', e)
"
done
timeout 300 python bench.py --workload heavyhex --chi 64 --prep 12 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_cfg3_${TAG}.log 2>&1
tail -1 gpurun_out/bench_cfg3_${TAG}.log | cut -c1-200
timeout 400 python bench.py --workload cubic3d --chi 16 --random-state --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_cfg4_${TAG}.log 2>&1
tail -1 gpurun_out/bench_cfg4_${TAG}.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_${TAG}.log 2>&1
tail -1 gpurun_out/bench_reference_${TAG}.log | cut -c1-400
