#!/bin/bash
# round 2, GPU call A: fused two-stage encoder -- first light
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
# each parametrised case in its own process: a trapped kernel kills the CUDA context of its process only
for t in "1-1-320" "1-16-320" "1-17-320" "3-50-320" "7-37-320" "40-50-320" "2-50-128" "2-33-256" "1-40-512" "2-50-336"; do
  echo "=== $t" >> gpurun_out/r2a_fused.log
  timeout 180 python -m pytest "tests/test_gpu_fused.py::test_two_stage_fused_encoder_features[$t]" -x -q 2>&1 | tail -15 >> gpurun_out/r2a_fused.log
done
timeout 300 python -m pytest tests/test_gpu_fused.py -q -k "not encoder_features" 2>&1 | tail -15 >> gpurun_out/r2a_fused.log
timeout 300 python bench.py --model two_stage --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/r2a_bench_ts.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/r2a_bench_msc.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
echo done
