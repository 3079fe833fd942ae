#!/bin/bash
# round-2 end measurement: GPU tests, smoke, default bench (all extras), per-model lines, reference arm, ncu launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r3f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3f_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r3f_bench_default.log 2> gpurun_out/r3f_bench_default.err
for m in two_stage enhanced ssd msc_n conv1d_msc msc_legacy improved hybrid complex; do
  timeout 300 python bench.py --model $m --steps 5 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3f_bench_$m.log 2>&1
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r3f_bench_ref.log 2>&1
timeout 200 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3f_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3f_launches.csv \
  python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r3f_ncu_launches.log 2>&1
echo done
