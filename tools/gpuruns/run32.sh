set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_covertree.py tests/test_gpu_configs.py -q -m gpu -k "covertree or cover_tree" > gpurun_out/r2_ct32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ct32.log
tail -30 gpurun_out/r2_ct32.log | cut -c1-220
CGGP_CT_TRACE=1 timeout 300 python tools/covertree_bench.py 8000000 2 0.11 > gpurun_out/r2_ctb32.log 2>&1; tail -32 gpurun_out/r2_ctb32.log
CGGP_CT_CLUSTER=0 CGGP_CT_TRACE=1 timeout 300 python tools/covertree_bench.py 8000000 2 0.11 > gpurun_out/r2_ctb32b.log 2>&1; grep -v "^covertree level" gpurun_out/r2_ctb32b.log | tail -4
timeout 300 python tools/covertree_bench.py 434874 3 0.8 > gpurun_out/r2_ctb32c.log 2>&1; tail -4 gpurun_out/r2_ctb32c.log
