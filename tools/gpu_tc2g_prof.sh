#!/bin/bash
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc2_gram -c 4 -o gpurun_out/prof_tc2g_${TAG} -f tools/tc2g_test prof > gpurun_out/ncu_tc2g_${TAG}.log 2>&1
ncu -i gpurun_out/prof_tc2g_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_tc2g_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_tc2g_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_tc2g_${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/prof_tc2g_${TAG}.ncu-rep
tail -5 gpurun_out/ncu_tc2g_${TAG}.log
O=gpurun_out/tc2g_sweep_${TAG}.txt
for v in "TNQS_TC2G_NB=1" "TNQS_TC2G_NB=1 TNQS_TC2G_DBG=1" "TNQS_TC2G_NB=1 TNQS_TC2G_DBG=2" "TNQS_TC2G_NB=4" "TNQS_TC2G_ACC=32"; do
  echo "== $v" >> $O
  env $v timeout 100 tools/tc2g_test quick 2>&1 | grep -A1 "bench" | grep -v "^--" | paste - - | sed 's/  */ /g' | cut -c1-60,120-400 >> $O
done
cat $O | cut -c1-300
