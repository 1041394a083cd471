#!/bin/bash
# multi-GPU evidence at N GPUs: the 2-GPU parity test (N >= 2), then the bench at chi=32 and / or chi=64
N=${1:-2}
TAG=${2:-r3c}
WHAT=${3:-32}
mkdir -p gpurun_out
if [ "$N" == "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short -p no:cacheprovider --timeout 500 > gpurun_out/pytest_mgpu_${TAG}.log 2>&1
  tail -3 gpurun_out/pytest_mgpu_${TAG}.log | cut -c1-200
fi
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"
}
if [[ "$WHAT" == *32* ]]; then
  run --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_16x16_chi32_${N}gpu_${TAG}.log 2>&1
  tail -1 gpurun_out/bench_16x16_chi32_${N}gpu_${TAG}.log | cut -c1-400
fi
if [[ "$WHAT" == *64* ]]; then
  run --chi 64 --random-state --steps 2 --warmup 1 --no-cpu --inplace > gpurun_out/bench_16x16_chi64_${N}gpu_${TAG}.log 2>&1
  tail -1 gpurun_out/bench_16x16_chi64_${N}gpu_${TAG}.log | cut -c1-400
fi
