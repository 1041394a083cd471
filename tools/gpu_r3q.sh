#!/bin/bash
# launch list (ncu --metrics gpu__time_duration.sum) of the bench's timed layer with the final kernels
TAG=${1:-r3q}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sampler --cuda-profiler"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_16x16_chi32_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
grep -c "gpu__time_duration" gpurun_out/launches_16x16_chi32_${TAG}.csv
