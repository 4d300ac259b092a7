set -x
mkdir -p gpurun_out
CGGP_CT_TRACE=1 timeout 300 python tools/covertree_bench.py 8000000 2 0.11 > gpurun_out/r2_ctb34.log 2>&1; tail -32 gpurun_out/r2_ctb34.log | cut -c1-200
