#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "two_stage or fused" 2>&1 | tail -6 > gpurun_out/r3g_pytest.log
timeout 300 python bench.py --model two_stage --steps 8 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3g_bench_two_stage.log 2>&1
PAUT_TS_DEBUG=1 timeout 120 python tools/run_stage.py --stage 1 --sets 20000 --reps 2 > gpurun_out/r3g_probe.log 2>&1
echo done
