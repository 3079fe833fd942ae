#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -q -k "two_stage" 2>&1 | tail -5 > gpurun_out/r2j_fused.log
PAUT_TS_DEBUG=1 timeout 200 python tools/run_stage.py --stage 1 --sets 2000 > gpurun_out/r2j_probe.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench_full.log 2>&1
echo done
