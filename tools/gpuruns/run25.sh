set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_gram.py 4096 2000000 11 > gpurun_out/r2_gram_c3.log 2>&1; cat gpurun_out/r2_gram_c3.log
timeout 600 python tools/bench_elbo.py > gpurun_out/r2_elbo_c3.log 2>&1; cat gpurun_out/r2_elbo_c3.log
