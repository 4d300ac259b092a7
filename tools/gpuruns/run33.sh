set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_covertree.py tests/test_gpu_configs.py -q -m gpu -k "covertree or cover_tree" > gpurun_out/r2_ct33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ct33.log
tail -30 gpurun_out/r2_ct33.log | cut -c1-220
CGGP_CT_TRACE=1 timeout 300 python tools/covertree_bench.py 8000000 2 0.11 > gpurun_out/r2_ctb33.log 2>&1; tail -22 gpurun_out/r2_ctb33.log | grep -v "^covertree level [1-8]: .* 0.[0-9]* ms"
timeout 300 python tools/covertree_bench.py 2000000 2 0.25 > gpurun_out/r2_ctb33b.log 2>&1; tail -3 gpurun_out/r2_ctb33b.log
timeout 300 python tools/covertree_bench.py 434874 3 0.8 > gpurun_out/r2_ctb33c.log 2>&1; tail -3 gpurun_out/r2_ctb33c.log
