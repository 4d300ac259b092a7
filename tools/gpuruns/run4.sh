set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
tail -5 gpurun_out/r2_t4.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench4.log 2>&1; tail -1 gpurun_out/r2_bench4.log
