#!/bin/bash
# Jacobi column-block size in the under-filled regime (few matrices per launch, as on a rank of an 8-GPU run): 6x6 lattice, chi=32
TAG=${1:-r3n}
mkdir -p gpurun_out
for bc in 16 8; do
  echo "== TNQS_JACOBI_BC=$bc" >> gpurun_out/jacobi_small_batch_${TAG}.txt
  TNQS_JACOBI_BC=$bc timeout 200 python tools/breakdown.py 6 32 random 2>&1 | grep -v "BP sweep after" | grep "SU colour\|mode TF" >> gpurun_out/jacobi_small_batch_${TAG}.txt
done
cat gpurun_out/jacobi_small_batch_${TAG}.txt | cut -c1-200
