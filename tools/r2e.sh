#!/bin/bash
# round 2, GPU call E: ts encoder v2 (register-resident stem weights, 8-warp epilogue) + ncu of the tcgen05 attention block
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest "tests/test_gpu_fused.py::test_two_stage_fused_encoder_features[1-16-320]" -x -q 2>&1 | tail -12 > gpurun_out/r2e_fused.log
timeout 600 python -m pytest tests/test_gpu_fused.py -q -k "two_stage" 2>&1 | tail -30 >> gpurun_out/r2e_fused.log
timeout 300 python bench.py --model two_stage --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/r2e_bench_ts.log 2>&1
timeout 200 python tools/run_stage.py --stage 1 --sets 2000 > gpurun_out/r2e_stage1.log 2>&1
timeout 200 python tools/run_stage.py --stage 2 --sets 592 > gpurun_out/r2e_stage2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msc_attn_tc -c 1 -o gpurun_out/r2e_attn_tc python tools/run_stage.py --stage 2 --sets 592 --reps 1 > gpurun_out/r2e_ncu.log 2>&1
echo done
