set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_covertree.py tests/test_gpu_configs.py -q -m gpu -k "covertree or cover_tree" > gpurun_out/r2_ct38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ct38.log
tail -5 gpurun_out/r2_ct38.log | cut -c1-220
timeout 300 python tools/covertree_bench.py 8000000 2 0.11 > gpurun_out/r2_ctb38.log 2>&1; tail -3 gpurun_out/r2_ctb38.log
timeout 300 python tools/covertree_bench.py 2000000 2 0.25 > gpurun_out/r2_ctb38b.log 2>&1; tail -3 gpurun_out/r2_ctb38b.log
timeout 300 python tools/covertree_bench.py 434874 3 0.8 > gpurun_out/r2_ctb38c.log 2>&1; tail -3 gpurun_out/r2_ctb38c.log
timeout 300 python tools/covertree_bench.py 200000 2 0.25 > gpurun_out/r2_ctb38d.log 2>&1; tail -3 gpurun_out/r2_ctb38d.log
