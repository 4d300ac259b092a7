set -x
mkdir -p gpurun_out
CGGP_PIPE_WIDE=1 timeout 300 python tools/bench_matvec.py c3 > gpurun_out/r2_c3_wide.log 2>&1; cat gpurun_out/r2_c3_wide.log
timeout 300 python tools/bench_matvec.py c3 > gpurun_out/r2_c3_narrow.log 2>&1; cat gpurun_out/r2_c3_narrow.log
CGGP_PIPE_WIDE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kuf_kfu_matvec or fused_matvec or matrix_free_cg or multi_rhs" > gpurun_out/r2_t16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t16.log; tail -3 gpurun_out/r2_t16.log
