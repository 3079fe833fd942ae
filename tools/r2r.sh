#!/bin/bash
# MSC at the bench's set size (300 A-scans per set): ncu --set full of the three fused kernels, with source
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/ncu_capture.py --kinds msc --n 300 --sets 1480 > gpurun_out/r2r_capture_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'k_msc_encoder_tc|k_msc_attn_block|k_msc_ffn_head' \
  -o gpurun_out/r2r_msc300 -f python tools/ncu_capture.py --kinds msc --n 300 --sets 1480 > gpurun_out/r2r_ncu.log 2>&1
echo done
