set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fourteen or kuf_kfu_matvec" > gpurun_out/r2_t41.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t41.log
tail -8 gpurun_out/r2_t41.log | cut -c1-200
timeout 300 python tools/bench_matvec.py c4 c2 2>&1 | grep -v simple > gpurun_out/r2_mv41.log; cat gpurun_out/r2_mv41.log
CGGP_PIPE_W14=0 timeout 300 python tools/bench_matvec.py c4 2>&1 | grep -v simple > gpurun_out/r2_mv41b.log; cat gpurun_out/r2_mv41b.log
