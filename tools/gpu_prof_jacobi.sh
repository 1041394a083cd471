CMD="python bench.py --L 8 --chi 32 --prep 15 --steps 1 --warmup 3 --no-cpu --cuda-profiler"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum --clock-control none -k regex:jacobi -c 6 --csv --log-file gpurun_out/jac_launches.csv $CMD > /dev/null 2>&1
grep -v "^==" gpurun_out/jac_launches.csv | cut -d, -f5,8,9,13,15 | head -30
