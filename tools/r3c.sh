#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
true
: > gpurun_out/r3c_sub.log
for s in 4096 8192 16384 32768; do
  echo "== PAUT_CONV_SUBCHUNK=$s" >> gpurun_out/r3c_sub.log
  PAUT_CONV_SUBCHUNK=$s timeout 300 python bench.py --model enhanced --steps 3 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3c_bench_enh_$s.log 2>&1
  python tools/bench_summary.py gpurun_out/r3c_bench_enh_$s.log | sed -n 2,4p | cut -c1-330 >> gpurun_out/r3c_sub.log
done
echo done
