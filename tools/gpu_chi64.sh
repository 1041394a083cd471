#!/bin/bash
TAG=${1:-c64}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "factorize or single_two_site" --tb=short -p no:cacheprovider --timeout 200 -x > gpurun_out/pytest_kw_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_kw_${TAG}.log
timeout 200 python bench.py --L 6 --chi 64 --random-state --steps 1 --warmup 1 --no-cpu --inplace > gpurun_out/bench_6x6_chi64_${TAG}.log 2>&1
tail -1 gpurun_out/bench_6x6_chi64_${TAG}.log | cut -c1-1500
timeout 420 python bench.py --L 16 --chi 64 --random-state --steps 2 --warmup 2 --no-cpu --inplace > gpurun_out/bench_16x16_chi64_${TAG}.log 2>&1
tail -1 gpurun_out/bench_16x16_chi64_${TAG}.log | cut -c1-1800
nvidia-smi --query-gpu=memory.used --format=csv
