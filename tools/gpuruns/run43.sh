set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2_t43.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t43.log
tail -3 gpurun_out/r2_t43.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2_bench43_g2.log 2>&1; tail -1 gpurun_out/r2_bench43_g2.log | cut -c1-300
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench43_g2_ref.log 2>&1; tail -1 gpurun_out/r2_bench43_g2_ref.log | cut -c1-200
