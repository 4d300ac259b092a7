set -x
mkdir -p gpurun_out
CGGP_PIPE_XWIDE=1 timeout 300 python tools/bench_matvec.py c4 c2 > gpurun_out/r2_c4_xwide.log 2>&1; cat gpurun_out/r2_c4_xwide.log
timeout 300 python tools/bench_matvec.py c4 c2 > gpurun_out/r2_c4_wide2.log 2>&1; cat gpurun_out/r2_c4_wide2.log
CGGP_PIPE_XWIDE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "many_row_blocks or deterministic or kuf_kfu_matvec" > gpurun_out/r2_t21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t21.log; tail -3 gpurun_out/r2_t21.log
