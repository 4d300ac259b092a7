set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651"
timeout 85 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-parity-check --balance > gpurun_out/r2_bench48_bal.log 2>&1; tail -1 gpurun_out/r2_bench48_bal.log | cut -c1-200
timeout 85 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-parity-check > gpurun_out/r2_bench48_even.log 2>&1; tail -1 gpurun_out/r2_bench48_even.log | cut -c1-200
