#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3d_pytest.log
timeout 300 python bench.py --model enhanced --steps 5 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3d_bench_enhanced.log 2>&1
echo done
