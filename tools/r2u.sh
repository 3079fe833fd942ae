#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2u_pytest.log
: > gpurun_out/r2u_variants.log
for v in 0 1 2 3 4 5 6; do
  echo "== PAUT_ATTN_VARIANT=$v" >> gpurun_out/r2u_variants.log
  PAUT_ATTN_VARIANT=$v timeout 120 python tools/run_stage.py --stage 4 --sets 6660 --reps 6 >> gpurun_out/r2u_variants.log 2>&1
done
echo "== cross (stage 5)" >> gpurun_out/r2u_variants.log
timeout 120 python tools/run_stage.py --stage 5 --sets 6660 --reps 6 >> gpurun_out/r2u_variants.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2u_bench_msc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msc_attn_block_p \
  -o gpurun_out/r2u_attn_p -f python tools/run_stage.py --stage 4 --sets 1480 --reps 1 > gpurun_out/r2u_ncu.log 2>&1
echo done
