set -x
mkdir -p gpurun_out
python tools/prof_case.py pipe1 434874 c2 > gpurun_out/plain_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kfu_pipe_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pipe1_c2 python tools/prof_case.py pipe1 434874 c2 > gpurun_out/ncu_c2.log 2>&1
python tools/prof_case.py pipe1 500000 c4 > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kfu_pipe_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pipe1_c4 python tools/prof_case.py pipe1 500000 c4 > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4; du -sh gpurun_out
