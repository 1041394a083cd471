#!/bin/bash
# one GPU call: parity tests, smoke, bench line, ncu launch list + full captures of the dominant kernels
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 400 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
tail -2 gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.log 2>&1
tail -1 gpurun_out/bench_${TAG}.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${TAG}.log 2>&1
tail -1 gpurun_out/bench_ref_${TAG}.log | cut -c1-300
timeout 150 python tools/breakdown.py > gpurun_out/breakdown_${TAG}.txt 2>&1
grep -v "BP sweep after" gpurun_out/breakdown_${TAG}.txt | tail -14
timeout 300 python bench.py --L 6 --chi 64 --prep 9 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_6x6_chi64_${TAG}.log 2>&1
tail -1 gpurun_out/bench_6x6_chi64_${TAG}.log | cut -c1-300
CMD="python bench.py --L 8 --chi 32 --prep 15 --steps 1 --warmup 3 --no-cpu --cuda-profiler"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_run.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_mode -s 6 -c 4 -o gpurun_out/prof_tc_mode_${TAG} -f $CMD > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gram -s 2 -c 4 -o gpurun_out/prof_tc_gram_${TAG} -f $CMD > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:jacobi_cluster -c 5 -o gpurun_out/prof_jacobi_${TAG} -f $CMD > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gram_dmma -c 2 -o gpurun_out/prof_sugram_${TAG} -f $CMD > gpurun_out/ncu_d.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:chol_prepare|colgram|colapply|su_factors" -c 5 -o gpurun_out/prof_small_${TAG} -f $CMD > gpurun_out/ncu_e.log 2>&1
# keep gpurun_out small (64 MiB cap): export the raw pages as CSV on the box and drop the reports
for k in tc_mode tc_gram jacobi sugram small; do
  f=gpurun_out/prof_${k}_${TAG}.ncu-rep
  if [ -f $f ]; then ncu -i $f --page raw --csv > gpurun_out/prof_${k}_${TAG}_raw.csv 2>/dev/null; rm -f $f; fi
done
ls -la gpurun_out/*${TAG}*
