# Session 3: full GPU suite, bench lines of the default config and of the section-8 "next" models
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/s3_pytest.log
timeout 300 python bench.py > gpurun_out/s3_bench_msc.log 2>gpurun_out/s3_bench_msc.err; echo "bench rc $?"
for m in msc_legacy improved hybrid complex; do timeout 300 python bench.py --model $m --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/s3_bench_$m.log 2>&1; echo "$m rc $?"; done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; tail -1 gpurun_out/s3_smoke.log
