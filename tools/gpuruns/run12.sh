set -x
mkdir -p gpurun_out
timeout 600 python tools/sanitize_case.py > gpurun_out/r2_san_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r2_san_mem.log 2>&1; echo "rc=$?" >> gpurun_out/r2_san_mem.log
tail -5 gpurun_out/r2_san_plain.log; tail -8 gpurun_out/r2_san_mem.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t12.log
tail -3 gpurun_out/r2_t12.log
timeout 300 python tools/bench_dense.py step > gpurun_out/r2_step12.log 2>&1; cat gpurun_out/r2_step12.log
