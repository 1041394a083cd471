#!/bin/bash
# multi-GPU: 2-GPU parity check vs the oracle, then chi=32 and chi=64 16x16 bench lines at this GPU count
N=${1:-2}
TAG=${2:-mg}
mkdir -p gpurun_out
if [ "$N" == "2" ]; then
  timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -s --tb=short -p no:cacheprovider --timeout 380 > gpurun_out/pytest_mgpu_${TAG}.log 2>&1
  grep -E "layer|PASSED|FAILED|passed|failed" gpurun_out/pytest_mgpu_${TAG}.log | tail -12
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --chi 64 --random-state --steps 2 --warmup 2 --no-cpu --inplace > gpurun_out/bench_16x16_chi64_${N}gpu_${TAG}.log 2>&1
tail -1 gpurun_out/bench_16x16_chi64_${N}gpu_${TAG}.log | cut -c1-400
