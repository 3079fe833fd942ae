set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3d_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/s3d_pytest.log
timeout 300 python bench.py --model improved --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/s3d_bench_improved.log 2>&1; echo "improved rc $?"
