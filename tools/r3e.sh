#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/mscn_check.py > gpurun_out/r3e_mscn_check.log 2>&1
echo done
