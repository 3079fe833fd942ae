set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/r4_pytest.log
for m in ssd two_stage enhanced conv1d_msc; do timeout 300 python bench.py --model $m --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/r4_bench_$m.log 2>&1; done
PAUT_CONV_DEBUG=1 timeout 120 python tools/conv_probe.py ssd 2>&1 | grep "conv probe" | tail -2
timeout 120 python tools/mma_probe.py 2>&1 | tail -14
