#!/bin/bash
# stand-alone test + bandwidth of the TMA-fed Gram kernel (tools/tc2g_test.cu, built here with nvcc before the call)
TAG=${1:-r2n}
O=gpurun_out/tc2g_test_${TAG}.txt
mkdir -p gpurun_out
timeout 300 tools/tc2g_test bench > $O 2>&1
echo "exit $?" >> $O
for v in $2; do
  echo "== $v" >> $O
  env ${v//,/ } timeout 100 tools/tc2g_test quick 2>&1 | grep -A1 "bench" | grep -v "^--" | paste - - | sed 's/  */ /g' | cut -c1-60,120-400 >> $O
done
cut -c1-250 $O | tail -90
