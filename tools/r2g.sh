#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -q -k "two_stage" 2>&1 | tail -30 > gpurun_out/r2g_fused.log
PAUT_TS_DEBUG=1 timeout 200 python tools/run_stage.py --stage 1 --sets 2000 > gpurun_out/r2g_probe.log 2>&1
timeout 300 python bench.py --model two_stage --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/r2g_bench_ts.log 2>&1
echo done
