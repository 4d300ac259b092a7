set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
timeout 110 $TR bench.py --gpus 8 --steps 60 --warmup 5 > gpurun_out/r2_bench47_g8.log 2>&1; tail -1 gpurun_out/r2_bench47_g8.log | cut -c1-300
timeout 60 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2_t47.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t47.log
tail -3 gpurun_out/r2_t47.log
