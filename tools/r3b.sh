#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r3b_probe.log
for n in 1 4; do
  echo "== PAUT_MSCN_CTAS=$n" >> gpurun_out/r3b_probe.log
  PAUT_MSCN_CTAS=$n PAUT_MSCN_DEBUG=1 timeout 120 python tools/run_stage.py --stage 6 --sets 3334 --reps 2 >> gpurun_out/r3b_probe.log 2>&1
done
echo done
