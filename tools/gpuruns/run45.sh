set -x
mkdir -p gpurun_out
CGGP_TF32_EPI=16 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu -k "tf32 or f16x3 or float32 or config5" > gpurun_out/r2_t45.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t45.log
tail -6 gpurun_out/r2_t45.log | cut -c1-220
CGGP_TF32_EPI=16 timeout 300 python tools/bench_matvec.py c5 > gpurun_out/r2_c5_epi16.log 2>&1; cat gpurun_out/r2_c5_epi16.log | cut -c1-200
CGGP_TF32_EPI=8 timeout 300 python tools/bench_matvec.py c5 > gpurun_out/r2_c5_epi8.log 2>&1; cat gpurun_out/r2_c5_epi8.log | cut -c1-200
