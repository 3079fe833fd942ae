#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in "False-1-300" "True-3-300" "False-2-37" "True-2-129" "False-1-320"; do
  echo "=== $t" >> gpurun_out/r2k_attn.log
  timeout 120 python -m pytest "tests/test_gpu_fused.py::test_msc_attention_block_tcgen05[$t]" -x -q 2>&1 | tail -12 >> gpurun_out/r2k_attn.log
done
timeout 300 python -m pytest tests/test_gpu_fused.py -q -k "attention" 2>&1 | tail -12 >> gpurun_out/r2k_attn.log
timeout 200 python tools/run_stage.py --stage 2 --sets 3334 > gpurun_out/r2k_stage2.log 2>&1
timeout 200 python tools/run_stage.py --stage 4 --sets 3334 >> gpurun_out/r2k_stage2.log 2>&1
PAUT_ATTN=tc timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2k_bench_tc.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2k_bench_mma.log 2>&1
echo done
