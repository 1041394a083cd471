#!/bin/bash
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 500 python tools/breakdown.py 16 64 random > gpurun_out/breakdown_chi64_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_chi64_${TAG}.txt | tail -16
