#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r3i_skew.log
for s in 0 3000 6000 12000 24000 36000 60000; do
  echo "== PAUT_ATTN_SKEW_NS=$s" >> gpurun_out/r3i_skew.log
  PAUT_ATTN_SKEW_NS=$s timeout 120 python tools/run_stage.py --stage 4 --sets 6660 --reps 8 >> gpurun_out/r3i_skew.log 2>&1
done
echo done
