CMD="python bench.py --L 8 --chi 32 --prep 15 --steps 1 --warmup 3 --no-cpu --cuda-profiler"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_mode -s 6 -c 4 -o gpurun_out/prof_tc -f $CMD > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
