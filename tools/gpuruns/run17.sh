set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t17.log
tail -3 gpurun_out/r2_t17.log
timeout 300 python tools/bench_matvec.py c3 > gpurun_out/r2_c3_17.log 2>&1; cat gpurun_out/r2_c3_17.log
CGGP_PIPE_DEEP=1 timeout 300 python tools/bench_matvec.py c3 > gpurun_out/r2_c3_17_deep.log 2>&1; cat gpurun_out/r2_c3_17_deep.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench17.log 2>&1; tail -1 gpurun_out/r2_bench17.log | cut -c1-200
python tools/prof_case.py pipe1 2000000 > gpurun_out/plain_pipe1w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kfu_pipe_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pipe1_wide python tools/prof_case.py pipe1 2000000 > gpurun_out/ncu_pipe1w.log 2>&1
