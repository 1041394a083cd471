#!/bin/bash
# round-2 evidence call: parity tests, fp64-MMA occupancy probe, ncu launch list + full captures of the 16x16 chi=32 bench layer
TAG=${1:-r2x}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300 | tail -8
timeout 120 tools/dmma_probe > gpurun_out/dmma_probe_${TAG}.txt 2>&1
cat gpurun_out/dmma_probe_${TAG}.txt
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sampler --cuda-profiler"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_16x16_chi32_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launch_${TAG}.log | cut -c1-200
prof() {  # name regex skip count
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/prof_$1_${TAG} -f $CMD > gpurun_out/ncu_$1_${TAG}.log 2>&1
  ncu -i gpurun_out/prof_$1_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_$1_${TAG}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$1_${TAG}.ncu-rep
  wc -l gpurun_out/prof_$1_${TAG}_raw.csv
}
prof tc2_mode tc2_mode_kernel 4 6
prof tc2_gram tc2_gram_kernel 2 4
prof gram_dmma gram_dmma_kernel 0 2
