set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2_bench13_g8.log 2>&1; tail -1 gpurun_out/r2_bench13_g8.log | cut -c1-300
timeout 900 $TR bench.py --gpus 8 --workload c4 --mode predict > gpurun_out/r2_pred13_g8.log 2>&1; tail -1 gpurun_out/r2_pred13_g8.log | cut -c1-600
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2_t13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t13.log; tail -3 gpurun_out/r2_t13.log
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29642"
timeout 600 $TR4 bench.py --gpus 4 --steps 60 --warmup 5 > gpurun_out/r2_bench13_g4.log 2>&1; tail -1 gpurun_out/r2_bench13_g4.log | cut -c1-300
