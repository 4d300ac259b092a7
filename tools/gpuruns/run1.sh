set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kuf_kfu_matvec or fused_matvec or matrix_free_cg" > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
timeout 600 python tools/bench_matvec.py multi c3 c2 c4 > gpurun_out/r2_multi.log 2>&1
timeout 600 python tools/bench_matvec.py c3 c2 > gpurun_out/r2_mv.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench1.log 2>&1
tail -5 gpurun_out/r2_t1.log; cat gpurun_out/r2_multi.log; cat gpurun_out/r2_mv.log; tail -2 gpurun_out/r2_bench1.log
