set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t11.log
tail -3 gpurun_out/r2_t11.log
timeout 300 python tools/bench_dense.py step > gpurun_out/r2_step11.log 2>&1; cat gpurun_out/r2_step11.log
timeout 300 python tools/bench_gram.py 16384 131072 2 > gpurun_out/r2_gram11.log 2>&1; cat gpurun_out/r2_gram11.log
timeout 600 python tools/bench_dense.py 2>&1 | tail -3 > gpurun_out/r2_dense11.log; cat gpurun_out/r2_dense11.log
