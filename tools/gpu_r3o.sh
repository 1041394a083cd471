#!/bin/bash
# source-level ncu profile of the cluster Jacobi in the latency-bound regime (6x6 lattice, chi=32: 18 theta matrices per colour)
TAG=${1:-r3o}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:jacobi_cluster" -s 8 -c 3 -o gpurun_out/prof_jacobi_${TAG} -f python tools/breakdown.py 6 32 random > gpurun_out/ncu_jacobi_${TAG}.log 2>&1
ncu -i gpurun_out/prof_jacobi_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_jacobi_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_jacobi_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_jacobi_${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/prof_jacobi_${TAG}.ncu-rep
wc -l gpurun_out/prof_jacobi_${TAG}_raw.csv gpurun_out/prof_jacobi_${TAG}_source.csv
tail -3 gpurun_out/ncu_jacobi_${TAG}.log | cut -c1-200
