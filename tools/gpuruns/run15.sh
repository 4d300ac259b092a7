set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t15.log
tail -3 gpurun_out/r2_t15.log
timeout 300 python tools/bench_matvec.py c2 c1 c4 > gpurun_out/r2_mv15.log 2>&1; cat gpurun_out/r2_mv15.log
timeout 300 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench15_c2.log 2>&1; tail -1 gpurun_out/r2_bench15_c2.log | cut -c1-200
