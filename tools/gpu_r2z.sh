#!/bin/bash
# experiments: pipeline decomposition / role cycle accounting of the tc2 mode product at chi=64, Jacobi block-size overrides
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 200 ./tools/tc2_test decomp > gpurun_out/tc2_decomp_${TAG}.txt 2>&1
timeout 200 ./tools/tc2_test roles > gpurun_out/tc2_roles_${TAG}.txt 2>&1
grep -A1 "dbg=\|decomp" gpurun_out/tc2_decomp_${TAG}.txt | grep "dbg=\|batch" | cut -c1-120
for bc in 16 8; do
  echo "== TNQS_JACOBI_BC=$bc" >> gpurun_out/su_probe_${TAG}.txt
  TNQS_JACOBI_BC=$bc PROBE_DEBUG=1 timeout 200 python tools/su_probe.py >> gpurun_out/su_probe_${TAG}.txt 2>&1
done
grep "==\|colour\|jacobi" gpurun_out/su_probe_${TAG}.txt | cut -c1-200
