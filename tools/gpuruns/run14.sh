set -x
mkdir -p gpurun_out
CGGP_PIPE_WIDE=1 timeout 300 python tools/bench_matvec.py c2 c1 > gpurun_out/r2_c2_wide.log 2>&1; cat gpurun_out/r2_c2_wide.log
CGGP_PIPE_WIDE=0 timeout 300 python tools/bench_matvec.py c2 c1 > gpurun_out/r2_c2_narrow.log 2>&1; cat gpurun_out/r2_c2_narrow.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench14.log 2>&1; tail -1 gpurun_out/r2_bench14.log | cut -c1-200
