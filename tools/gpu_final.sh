# Round-end measurement: GPU tests, bench lines of every model, reference arm, ncu launch list of the default bench
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/rf_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/rf_pytest.log
timeout 300 python bench.py > gpurun_out/rf_bench_msc.log 2>gpurun_out/rf_bench_msc.err; echo "bench rc $?"
for m in two_stage ssd conv1d_msc enhanced msc_n; do timeout 300 python bench.py --model $m --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/rf_bench_$m.log 2>&1; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/rf_bench_ref.log 2>&1
timeout 120 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/rf_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rf_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/rf_ncu_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rf_smoke.log 2>&1; tail -1 gpurun_out/rf_smoke.log
