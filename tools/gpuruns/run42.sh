set -x
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke42.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke42.log; tail -2 gpurun_out/r2_smoke42.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t42.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t42.log
tail -3 gpurun_out/r2_t42.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench42.log 2>&1; tail -1 gpurun_out/r2_bench42.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench42_ref.log 2>&1; tail -1 gpurun_out/r2_bench42_ref.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain_bench42.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches42.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench42.log 2>&1
python tools/prof_case.py pipe1 2000000 > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kfu_pipe_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pipe1_c3_final python tools/prof_case.py pipe1 2000000 > gpurun_out/ncu_c3.log 2>&1
du -sh gpurun_out
