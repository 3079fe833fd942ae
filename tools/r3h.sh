#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r3h_pytest.log
for m in two_stage enhanced ssd conv1d_msc; do
  timeout 300 python bench.py --model $m --steps 5 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3h_bench_$m.log 2>&1
done
echo done
