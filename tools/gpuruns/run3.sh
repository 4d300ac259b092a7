set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -5 gpurun_out/r2_t3.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.log 2>&1; tail -1 gpurun_out/r2_bench3.log
CGGP_FUSED_TAIL=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3_notail.log 2>&1; tail -1 gpurun_out/r2_bench3_notail.log
timeout 600 python bench.py --workload c2 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3_c2.log 2>&1; tail -1 gpurun_out/r2_bench3_c2.log
