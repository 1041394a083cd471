#!/bin/bash
# stand-alone test of the TMA-fed tcgen05 mode product (correctness vs CPU, then bandwidth, then one ncu capture)
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 300 ./tools/tc2_test bench $2 > gpurun_out/tc2_test_${TAG}.txt 2>&1
echo "tc2_test exit $?" >> gpurun_out/tc2_test_${TAG}.txt
grep -E "FAIL|ELIG|batch|failed|exit" gpurun_out/tc2_test_${TAG}.txt | cut -c1-250
if [ "$3" == "ncu" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc2_mode -c 1 -o gpurun_out/prof_tc2_${TAG} -f ./tools/tc2_test prof > gpurun_out/ncu_tc2_${TAG}.log 2>&1
  ncu -i gpurun_out/prof_tc2_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_tc2_${TAG}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_tc2_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_tc2_${TAG}_source.csv 2>/dev/null
  ls -la gpurun_out/prof_tc2_${TAG}*
fi
