set -x
mkdir -p gpurun_out
CGGP_PIPE_W12=1 timeout 300 python tools/bench_matvec.py c3 > gpurun_out/r2_c3_w12.log 2>&1; cat gpurun_out/r2_c3_w12.log
timeout 300 python tools/bench_matvec.py c3 > gpurun_out/r2_c3_w16.log 2>&1; cat gpurun_out/r2_c3_w16.log
CGGP_PIPE_W12=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "many_row_blocks or deterministic" > gpurun_out/r2_t19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t19.log; tail -3 gpurun_out/r2_t19.log
