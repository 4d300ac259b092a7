set -x
mkdir -p gpurun_out
for r in 512 2048 4096; do CGGP_GRAM_ROWS=$r timeout 300 python tools/bench_gram.py 16384 131072 2 >> gpurun_out/r2_gram10.log 2>&1; done
timeout 300 python tools/bench_gram.py 16384 131072 2 >> gpurun_out/r2_gram10.log 2>&1
timeout 300 python tools/bench_gram.py 4096 262144 11 >> gpurun_out/r2_gram10.log 2>&1
cat gpurun_out/r2_gram10.log
python tools/prof_case.py pipe1 2000000 > gpurun_out/plain_pipe1full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kfu_pipe_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pipe1_full python tools/prof_case.py pipe1 2000000 > gpurun_out/ncu_pipe1full.log 2>&1
tail -2 gpurun_out/ncu_pipe1full.log | cut -c1-200
timeout 300 python tools/bench_dense.py step > gpurun_out/r2_step10.log 2>&1; cat gpurun_out/r2_step10.log
