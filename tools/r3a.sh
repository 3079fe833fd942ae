#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "msc_n or mscn or MSC_N" 2>&1 | tail -15 > gpurun_out/r3a_pytest.log
timeout 300 python tools/mscn_check.py --reps 2 > gpurun_out/r3a_check.log 2>&1
timeout 300 python bench.py --model msc_n --steps 10 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3a_bench_msc_n.log 2>&1
echo done
