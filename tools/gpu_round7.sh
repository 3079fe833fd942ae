set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r9_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 > gpurun_out/r9_bench_msc.log 2>&1
PAUT_ENC_DEBUG=1 timeout 120 python tools/enc_probe.py 2>&1 | grep "enc probe" | tail -8
