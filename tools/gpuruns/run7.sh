set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -4 gpurun_out/r2_t7.log
timeout 600 python tools/bench_matvec.py multi c4 > gpurun_out/r2_multi_c4_wide.log 2>&1; head -4 gpurun_out/r2_multi_c4_wide.log
CGGP_PIPE_WIDE=0 timeout 600 python tools/bench_matvec.py c4 > gpurun_out/r2_c4_narrow.log 2>&1; cat gpurun_out/r2_c4_narrow.log
timeout 600 python tools/bench_matvec.py c4 > gpurun_out/r2_c4_wide.log 2>&1; cat gpurun_out/r2_c4_wide.log
python tools/prof_case.py pipe8 > gpurun_out/plain_pipe8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kfu_pipe8 -s 2 -c 1 -f -o gpurun_out/r2_prof_pipe8 python tools/prof_case.py pipe8 > gpurun_out/ncu_pipe8.log 2>&1
python tools/prof_case.py pipe1 > gpurun_out/plain_pipe1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kfu_pipe_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pipe1 python tools/prof_case.py pipe1 > gpurun_out/ncu_pipe1.log 2>&1
python tools/prof_case.py cg > gpurun_out/plain_cg.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_cg.csv python tools/prof_case.py cg > gpurun_out/ncu_cg.log 2>&1
tail -3 gpurun_out/ncu_pipe8.log gpurun_out/ncu_cg.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench7.log 2>&1; tail -1 gpurun_out/r2_bench7.log | cut -c1-300
