#!/bin/bash
# one GPU call: parity tests, smoke, bench line, ncu launch list + full captures of the dominant kernels
TAG=${1:-r1b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 400 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
tail -2 gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.log 2>&1
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-400
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
cat gpurun_out/breakdown_${TAG}.txt | tail -12
CMD="python bench.py --L 8 --chi 32 --prep 15 --steps 1 --warmup 3 --no-cpu --cuda-profiler"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_run.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_mode -s 6 -c 4 -o gpurun_out/prof_tc_mode_${TAG} -f $CMD > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gram -s 2 -c 4 -o gpurun_out/prof_tc_gram_${TAG} -f $CMD > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:jacobi -c 3 -o gpurun_out/prof_jacobi_${TAG} -f $CMD > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gram_dmma -c 2 -o gpurun_out/prof_sugram_${TAG} -f $CMD > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
