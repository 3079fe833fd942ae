#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2v_pytest.log
: > gpurun_out/r2v_variants.log
for v in 0 10 12 6 1 3; do
  echo "== PAUT_ATTN_VARIANT=$v" >> gpurun_out/r2v_variants.log
  PAUT_ATTN_VARIANT=$v timeout 120 python tools/run_stage.py --stage 4 --sets 6660 --reps 6 >> gpurun_out/r2v_variants.log 2>&1
done
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2v_bench_msc.log 2>&1
echo done
