set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_covertree.py -q -m gpu -k "large" > gpurun_out/r2_ct31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ct31.log
tail -5 gpurun_out/r2_ct31.log | cut -c1-220
timeout 300 python tools/covertree_bench.py 2000000 2 0.25 > gpurun_out/r2_ctb31.log 2>&1; tail -4 gpurun_out/r2_ctb31.log
timeout 300 python tools/covertree_bench.py 8000000 2 0.11 > gpurun_out/r2_ctb31b.log 2>&1; tail -4 gpurun_out/r2_ctb31b.log
timeout 300 python tools/covertree_bench.py 434874 3 0.8 > gpurun_out/r2_ctb31c.log 2>&1; tail -4 gpurun_out/r2_ctb31c.log
