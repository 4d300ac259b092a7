set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t27.log
tail -3 gpurun_out/r2_t27.log
timeout 300 python tools/bench_matvec.py c3 c2 c4 2>&1 | grep -v simple > gpurun_out/r2_mv27.log; cat gpurun_out/r2_mv27.log
timeout 300 python tools/bench_matvec.py multi c3 2>&1 > gpurun_out/r2_multi27.log; cat gpurun_out/r2_multi27.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench27.log 2>&1; tail -1 gpurun_out/r2_bench27.log | cut -c1-200
