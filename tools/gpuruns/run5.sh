set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t5.log
tail -5 gpurun_out/r2_t5.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench5_g2.log 2>&1; tail -1 gpurun_out/r2_bench5_g2.log
CGGP_PEER_ALLREDUCE=0 timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 3 --no-parity-check > gpurun_out/r2_bench5_g2_nccl.log 2>&1; tail -1 gpurun_out/r2_bench5_g2_nccl.log
CGGP_FUSED_TAIL=0 CGGP_PEER_ALLREDUCE=0 timeout 600 $TR bench.py --gpus 2 --steps 40 --warmup 3 --no-parity-check > gpurun_out/r2_bench5_g2_old.log 2>&1; tail -1 gpurun_out/r2_bench5_g2_old.log
timeout 900 $TR bench.py --gpus 2 --workload c4 --mode predict --predict-rows 400000 --predict-inducing 4096 --predict-points 8192 > gpurun_out/r2_pred5_g2.log 2>&1; tail -1 gpurun_out/r2_pred5_g2.log
