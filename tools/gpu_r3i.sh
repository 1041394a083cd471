#!/bin/bash
# ncu --set full captures at chi=64 (10x10 random TNS, one timed layer): tc2 mode product, tc2 Gram, fp64 Gram; and of the
# rewritten gram_dmma at chi=32 (16x16 bench layer)
TAG=${1:-r3i}
mkdir -p gpurun_out
CMD64="python bench.py --L 10 --chi 64 --random-state --steps 1 --warmup 1 --no-cpu --no-sampler --inplace --cuda-profiler"
CMD32="python bench.py --steps 1 --warmup 3 --no-cpu --no-sampler --cuda-profiler"
prof() {  # name regex skip count cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o gpurun_out/prof_${name}_${TAG} -f "$@" > gpurun_out/ncu_${name}_${TAG}.log 2>&1
  ncu -i gpurun_out/prof_${name}_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${name}_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${name}_${TAG}.ncu-rep
  wc -l gpurun_out/prof_${name}_${TAG}_raw.csv
}
prof chi64_tc2_mode tc2_mode_kernel 6 6 $CMD64
prof chi64_tc2_gram tc2_gram_kernel 0 3 $CMD64
prof chi64_gram_dmma gram_dmma_kernel 0 2 $CMD64
prof chi32_gram_dmma gram_dmma_kernel 0 2 $CMD32
