set -x
mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/r2_bench49.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_bench49.log | cut -c1-260
