#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 0 1 0 1; do
PAUT_ENC_ROT=$i timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2z3_bench_msc_rot$i.log 2>&1
python tools/bench_summary.py gpurun_out/r2z3_bench_msc_rot$i.log | sed -n 2,4p >> gpurun_out/r2z3_rot_ab.log
done
timeout 600 python -m pytest tests -m gpu -q -x -k "msc or MSC or full_size" 2>&1 | tail -5 > gpurun_out/r2z3_pytest.log
echo done
