#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "msc or MSC or full_size or scanner or alternate" 2>&1 | tail -5 > gpurun_out/r3l_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3l_bench_msc.log 2>&1
PAUT_ATTN=notail timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3l_bench_msc_notail.log 2>&1
echo done
