set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_covertree.py -q -m gpu -k "not large" > gpurun_out/r2_ct30.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ct30.log
tail -40 gpurun_out/r2_ct30.log | cut -c1-220
timeout 200 python tools/covertree_bench.py 200000 2 0.25 > gpurun_out/r2_ctb30.log 2>&1; cat gpurun_out/r2_ctb30.log | tail -5
