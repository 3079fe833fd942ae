set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "msc and not legacy" > gpurun_out/s3c_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/s3c_pytest.log
timeout 200 python bench.py --cpu-seconds 0 > gpurun_out/s3c_bench_msc.log 2>gpurun_out/s3c_bench_msc.err; echo "bench rc $?"
PAUT_ENC_DEBUG=1 timeout 100 python bench.py --cpu-seconds 0 --steps 2 --warmup 3 > /dev/null 2>gpurun_out/s3c_probe.err; grep "enc probe" gpurun_out/s3c_probe.err | tail -8
