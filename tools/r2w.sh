#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2x_variants.log
for v in 0 30 31 0 30 31; do
  echo "== PAUT_ATTN_VARIANT=$v" >> gpurun_out/r2x_variants.log
  PAUT_ATTN_VARIANT=$v timeout 120 python tools/run_stage.py --stage 4 --sets 6660 --reps 8 >> gpurun_out/r2x_variants.log 2>&1
done
echo done
