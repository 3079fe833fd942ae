#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in "1-16-320" "3-170-320" "2-50-128"; do
  echo "=== $t" >> gpurun_out/r2l_mscn.log
  timeout 120 python -m pytest "tests/test_gpu_fused.py::test_msc_n_fused_front_end[$t]" -x -q 2>&1 | tail -12 >> gpurun_out/r2l_mscn.log
done
timeout 300 python -m pytest tests/test_gpu_fused.py -q -k "msc_n" 2>&1 | tail -20 >> gpurun_out/r2l_mscn.log
timeout 300 python bench.py --model msc_n --steps 10 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2l_bench_mscn.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2l_pytest.log
echo done
