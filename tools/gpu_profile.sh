#!/bin/bash
# round evidence: bench line, ncu launch list of one saturated layer, full captures of the dominant kernels
TAG=${1:-r1b}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.log 2>&1
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-200
CMD="python bench.py --L 8 --chi 32 --prep 15 --steps 1 --warmup 3 --no-cpu --cuda-profiler"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_mode -s 6 -c 4 -o gpurun_out/prof_tc_mode_${TAG} -f $CMD > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gram -s 2 -c 4 -o gpurun_out/prof_tc_gram_${TAG} -f $CMD > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:jacobi -c 3 -o gpurun_out/prof_jacobi_${TAG} -f $CMD > gpurun_out/ncu_c.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:gram_kernel<float, double" -c 2 -o gpurun_out/prof_sugram_${TAG} -f $CMD > gpurun_out/ncu_d.log 2>&1
ls gpurun_out/*.ncu-rep
