set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t50.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t50.log
tail -3 gpurun_out/r2_t50.log
