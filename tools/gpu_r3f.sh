#!/bin/bash
# why the 10x10 chi=32 random-state run of r3e was slow (cold pools?) + BASELINE config 4
TAG=${1:-r3f}
mkdir -p gpurun_out
TNQS_SLOWLOG=1 timeout 200 python bench.py --L 10 --chi 32 --random-state --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_cfg5_chi32_w1_${TAG}.log 2>&1
grep "tnqs slow" gpurun_out/bench_cfg5_chi32_w1_${TAG}.log | cut -c1-160 | sort | uniq -c | sort -rn | head -12
for w in 3; do
timeout 200 python bench.py --L 10 --chi 32 --random-state --steps 3 --warmup $w --no-cpu > gpurun_out/bench_cfg5_chi32_w${w}_${TAG}.log 2>&1
tail -1 gpurun_out/bench_cfg5_chi32_w${w}_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('warmup $w: gates/s %.1f steps %s bp_sweep_ms %.2f e2e %.1f' % (d['value'], d['config']['step_ms'], d['bp_sweep_ms'], d['e2e']['value']))"
done
timeout 200 python bench.py --L 10 --chi 32 --random-state --steps 3 --warmup 1 --no-cpu --inplace > gpurun_out/bench_cfg5_chi32_inplace_${TAG}.log 2>&1
tail -1 gpurun_out/bench_cfg5_chi32_inplace_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('inplace: gates/s %.1f steps %s bp_sweep_ms %.2f' % (d['value'], d['config']['step_ms'], d['bp_sweep_ms']))"
timeout 500 python bench.py --workload cubic3d --chi 16 --random-state --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_cfg4_${TAG}.log 2>&1
tail -1 gpurun_out/bench_cfg4_${TAG}.log | cut -c1-200
