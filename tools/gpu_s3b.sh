set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "legacy or improved or hybrid or complex or next" > gpurun_out/s3b_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/s3b_pytest.log
timeout 200 python tools/post_probe.py > gpurun_out/s3b_post_probe.log 2>&1; cat gpurun_out/s3b_post_probe.log | tail -5
for m in improved hybrid complex; do timeout 300 python bench.py --model $m --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/s3b_bench_$m.log 2>&1; echo "$m rc $?"; done
