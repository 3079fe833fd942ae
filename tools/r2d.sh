#!/bin/bash
# round 2, GPU call D: tcgen05 attention block first light + ncu of the fused two-stage encoder
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in "False-1-300" "True-1-300" "False-2-37" "True-2-128" "False-2-129" "True-1-16" "False-300-300" "True-300-300"; do
  echo "=== $t" >> gpurun_out/r2d_attn.log
  timeout 180 python -m pytest "tests/test_gpu_fused.py::test_msc_attention_block_tcgen05[$t]" -x -q 2>&1 | tail -12 >> gpurun_out/r2d_attn.log
done
timeout 300 python -m pytest tests/test_gpu_fused.py -q -k "attention" 2>&1 | tail -30 >> gpurun_out/r2d_attn.log
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/r2d_bench_msc.log 2>&1
timeout 200 python tools/run_stage.py --stage 1 --sets 2000 > gpurun_out/r2d_stage1.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ts_encoder -c 1 -o gpurun_out/r2d_ts_encoder python tools/run_stage.py --stage 1 --sets 2000 --reps 1 > gpurun_out/r2d_ncu.log 2>&1
echo done
