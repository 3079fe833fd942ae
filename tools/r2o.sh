#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py tests/test_gpu_next.py -q -x -k "enhanced or ssd or conv1d or hybrid or complex or improved" 2>&1 | tail -8 > gpurun_out/r2o_pytest.log
for m in enhanced ssd conv1d_msc; do
  timeout 300 python bench.py --model $m --steps 5 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2o_bench_$m.log 2>&1
done
for n in 1 2 3 4; do PAUT_MSCN_CTAS=$n timeout 100 python tools/run_stage.py --stage 6 --sets 3334 >> gpurun_out/r2o_mscn_ctas.log 2>&1; done
echo done
