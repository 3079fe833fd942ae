#!/bin/bash
# attention block: where do the cycles go?  timing variants of the persistent kernel, then one ncu capture with source
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2t_variants.log
for v in 0 1 2 3 4 5 6; do
  echo "== PAUT_ATTN_VARIANT=$v" >> gpurun_out/r2t_variants.log
  PAUT_ATTN_VARIANT=$v timeout 120 python tools/run_stage.py --stage 4 --sets 6660 --reps 6 >> gpurun_out/r2t_variants.log 2>&1
done
echo "== PAUT_ATTN=v1" >> gpurun_out/r2t_variants.log
PAUT_ATTN=v1 timeout 120 python tools/run_stage.py --stage 4 --sets 6660 --reps 6 >> gpurun_out/r2t_variants.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msc_attn_block_p \
  -o gpurun_out/r2t_attn_p -f python tools/run_stage.py --stage 4 --sets 1480 --reps 1 > gpurun_out/r2t_ncu.log 2>&1
echo done
