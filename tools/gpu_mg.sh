#!/bin/bash
# multi-GPU bench lines: 16x16 TFIM at chi=64 (random state generated on the device) and chi=32 at this GPU count
N=${1:-2}
TAG=${2:-mg}
WHAT=${3:-64}
mkdir -p gpurun_out
run() {
  if [ "$N" == "1" ]; then timeout 600 python bench.py "$@"; else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; fi
}
if [[ "$WHAT" == *64* ]]; then
  run --chi 64 --random-state --steps 2 --warmup 1 --no-cpu --inplace > gpurun_out/bench_16x16_chi64_${N}gpu_${TAG}.log 2>&1
  tail -1 gpurun_out/bench_16x16_chi64_${N}gpu_${TAG}.log | cut -c1-330
fi
if [[ "$WHAT" == *32* ]]; then
  run --chi 32 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_16x16_chi32_${N}gpu_${TAG}.log 2>&1
  tail -1 gpurun_out/bench_16x16_chi32_${N}gpu_${TAG}.log | cut -c1-330
fi
