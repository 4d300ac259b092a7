set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference_tests.py tests/test_gpu_parity.py -q -m gpu -k "reference or rff or test_cg or log_determinant" > gpurun_out/r2_t44.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t44.log
tail -25 gpurun_out/r2_t44.log | cut -c1-220
