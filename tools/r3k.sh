#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "two_stage or ssd" 2>&1 | tail -4 > gpurun_out/r3k_pytest.log
for m in two_stage ssd; do
  timeout 300 python bench.py --model $m --steps 8 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3k_bench_$m.log 2>&1
done
PAUT_LRL_UNFUSED=1 timeout 300 python bench.py --model two_stage --steps 8 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3k_bench_two_stage_unfused.log 2>&1
echo done
