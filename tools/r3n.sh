#!/bin/bash
# final multi-GPU record: the bench line at all GPUs of the box (weak scaling, one 1 M-A-scan shard per rank)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --cpu-seconds 0 > gpurun_out/r3n_bench_$N.log 2>&1
echo done
