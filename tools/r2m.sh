#!/bin/bash
# multi-GPU: H2D ceiling and the bench at N ranks (N = number of visible GPUs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/r2m_topo_$N.log 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name" > gpurun_out/r2m_cpu_$N.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py > gpurun_out/r2m_h2d_$N.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m_bench_$N.log 2>&1
PAUT_BENCH_LANES=4 PAUT_BENCH_CHUNK_ASCANS=38400 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/r2m_bench_${N}_lanes4.log 2>&1
echo done
