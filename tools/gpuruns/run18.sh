set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651"
timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2_bench18_g8.log 2>&1; tail -1 gpurun_out/r2_bench18_g8.log | cut -c1-300
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29652"
timeout 600 $TR4 bench.py --gpus 4 --steps 60 --warmup 5 > gpurun_out/r2_bench18_g4.log 2>&1; tail -1 gpurun_out/r2_bench18_g4.log | cut -c1-300
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29653"
timeout 600 $TR2 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2_bench18_g2.log 2>&1; tail -1 gpurun_out/r2_bench18_g2.log | cut -c1-300
timeout 600 $TR bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-parity-check > gpurun_out/r2_bench18_c4_g8.log 2>&1; tail -1 gpurun_out/r2_bench18_c4_g8.log | cut -c1-300
