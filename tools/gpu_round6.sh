set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r8_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r8_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 > gpurun_out/r8_bench_msc.log 2>&1
for m in ssd two_stage; do timeout 300 python bench.py --model $m --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/r8_bench_$m.log 2>&1; done
PAUT_ENC_DEBUG=1 timeout 120 python tools/enc_probe.py 2>&1 | grep "enc probe" | tail -8
