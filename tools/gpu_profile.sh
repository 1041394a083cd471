#!/bin/bash
# ncu launch list + full captures of the dominant kernels, restricted to the timed (χ-saturated) steps
mkdir -p gpurun_out
CMD="python bench.py --L 8 --chi 32 --prep 15 --steps 1 --warmup 3 --no-cpu --cuda-profiler"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mode_product -s 20 -c 4 -o gpurun_out/prof_mode -f $CMD > gpurun_out/ncu_mode.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gram_kernel -s 4 -c 4 -o gpurun_out/prof_gram -f $CMD > gpurun_out/ncu_gram.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:jacobi -c 3 -o gpurun_out/prof_jacobi -f $CMD > gpurun_out/ncu_jacobi.log 2>&1
ls -la gpurun_out/
