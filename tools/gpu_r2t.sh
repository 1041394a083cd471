#!/bin/bash
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 200 tools/tc2_test bench > gpurun_out/tc2_test_${TAG}.txt 2>&1
grep "case(s) failed" gpurun_out/tc2_test_${TAG}.txt
grep -A1 "^bench" gpurun_out/tc2_test_${TAG}.txt | grep batch | cut -c30-200
TNQS_SLOWLOG=1 timeout 400 python bench.py --chi 64 --random-state --steps 1 --warmup 1 --no-cpu --inplace > gpurun_out/slowlog_chi64_${TAG}.log 2>&1
grep "tnqs slow" gpurun_out/slowlog_chi64_${TAG}.log | cut -c1-300 | tail -40
tail -1 gpurun_out/slowlog_chi64_${TAG}.log | cut -c1-200
