#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2y_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2y_bench_msc.log 2>&1
PAUT_ENC_DEBUG=1 timeout 120 python tools/enc_probe.py > gpurun_out/r2y_enc_probe.log 2>&1
echo done
