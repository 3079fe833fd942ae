# GPU tests + one bench line per conv model (development loop)
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/rm_pytest.log 2>&1; tail -3 gpurun_out/rm_pytest.log
for m in ssd two_stage enhanced conv1d_msc; do timeout 300 python bench.py --model $m --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/rm_bench_$m.log 2>&1; done
