set -x
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke20.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke9.log; tail -2 gpurun_out/r2_smoke9.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t20.log
tail -3 gpurun_out/r2_t20.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench20.log 2>&1; tail -1 gpurun_out/r2_bench20.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench20_ref.log 2>&1; tail -1 gpurun_out/r2_bench20_ref.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches_final.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
