#!/bin/bash
# parity tests + breakdown + ncu launch list of one saturated 8x8 layer
TAG=${1:-q}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --tb=short -x -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -8 gpurun_out/pytest_gpu_${TAG}.log
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
cat gpurun_out/breakdown_${TAG}.txt | tail -30
CMD="python bench.py --L 16 --chi 32 --prep 15 --steps 1 --warmup 1 --no-cpu --cuda-profiler"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_run.log 2>&1
tail -2 gpurun_out/ncu_launch_run.log | cut -c1-200
