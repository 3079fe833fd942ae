#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/ncu_capture.py --kinds two_stage --n 50 --sets 4000 > gpurun_out/r3j_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc' -c 12 \
  -o gpurun_out/r3j_gemm -f python tools/ncu_capture.py --kinds two_stage --n 50 --sets 4000 > gpurun_out/r3j_ncu.log 2>&1
echo done
