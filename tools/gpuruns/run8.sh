set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2_t8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t8.log
tail -3 gpurun_out/r2_t8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench8_g2.log 2>&1; tail -1 gpurun_out/r2_bench8_g2.log | cut -c1-200
CGGP_FUSED_TAIL=0 CGGP_PEER_ALLREDUCE=0 timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 3 --no-parity-check > gpurun_out/r2_bench8_g2_old.log 2>&1; tail -1 gpurun_out/r2_bench8_g2_old.log | cut -c1-200
CGGP_TAIL_SHARD=0 timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench8_g2_noshard.log 2>&1; tail -1 gpurun_out/r2_bench8_g2_noshard.log | cut -c1-200
