set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2_bench6_g8.log 2>&1; tail -1 gpurun_out/r2_bench6_g8.log | cut -c1-400
CGGP_PEER_ALLREDUCE=0 timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 5 --no-parity-check > gpurun_out/r2_bench6_g8_nccl.log 2>&1; tail -1 gpurun_out/r2_bench6_g8_nccl.log | cut -c1-200
CGGP_FUSED_TAIL=0 CGGP_PEER_ALLREDUCE=0 timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 5 --no-parity-check > gpurun_out/r2_bench6_g8_old.log 2>&1; tail -1 gpurun_out/r2_bench6_g8_old.log | cut -c1-200
timeout 900 $TR bench.py --gpus 8 --workload c4 --mode predict > gpurun_out/r2_pred6_g8.log 2>&1; tail -1 gpurun_out/r2_pred6_g8.log
timeout 600 $TR bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-parity-check > gpurun_out/r2_bench6_c4_g8.log 2>&1; tail -1 gpurun_out/r2_bench6_c4_g8.log | cut -c1-300
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log; tail -3 gpurun_out/r2_t6.log
