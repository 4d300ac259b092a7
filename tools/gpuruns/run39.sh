set -x
mkdir -p gpurun_out
timeout 600 python tools/config2_pipeline.py > gpurun_out/r2_c2pipe39.log 2>&1; tail -3 gpurun_out/r2_c2pipe39.log | cut -c1-1500
