#!/bin/bash
# ncu evidence of the 16x16 chi=32 bench itself: launch list of one timed layer + full captures of the dominant kernels
TAG=${1:-r2m}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --cuda-profiler"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_16x16_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launch_${TAG}.log | cut -c1-300
cap() {  # name regex skip count
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/prof_$1_${TAG} -f $CMD > gpurun_out/ncu_$1_${TAG}.log 2>&1
  ncu -i gpurun_out/prof_$1_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_$1_${TAG}_raw.csv 2>/dev/null
  ls -la gpurun_out/prof_$1_${TAG}.ncu-rep
  if [ $(stat -c %s gpurun_out/prof_$1_${TAG}.ncu-rep) -gt 15000000 ]; then rm gpurun_out/prof_$1_${TAG}.ncu-rep; fi
}
cap tc2mode tc2_mode 2 3
cap tcgram tc_gram 0 2
cap dmma gram_dmma 0 1
