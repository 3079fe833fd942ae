#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest "tests/test_gpu_fused.py::test_two_stage_fused_encoder_features[1-16-320]" -x -q 2>&1 | tail -12 > gpurun_out/r2i_fused.log
timeout 120 python -m pytest "tests/test_gpu_fused.py::test_two_stage_fused_encoder_features[3-50-320]" -x -q 2>&1 | tail -12 >> gpurun_out/r2i_fused.log
timeout 600 python -m pytest tests/test_gpu_fused.py -q -k "two_stage" 2>&1 | tail -30 >> gpurun_out/r2i_fused.log
PAUT_TS_DEBUG=1 timeout 200 python tools/run_stage.py --stage 1 --sets 2000 > gpurun_out/r2i_probe.log 2>&1
timeout 300 python bench.py --model two_stage --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/r2i_bench_ts.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -q -k "two_stage" 2>&1 | tail -8 >> gpurun_out/r2i_fused.log
echo done
