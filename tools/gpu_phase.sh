#!/bin/bash
TAG=${1:-r2w}
mkdir -p gpurun_out
TNQS_SLOWLOG=1 TNQS_PHASELOG=1 timeout 300 python tools/breakdown.py > gpurun_out/phase_${TAG}.txt 2>&1
grep "tnqs slow" gpurun_out/phase_${TAG}.txt | tail -8 | cut -c1-250
