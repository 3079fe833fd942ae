# Session-3 last measurement: full GPU suite, default bench (+ reference arm), ncu launch list and encoder capture
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3z_pytest.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/s3z_pytest.log
timeout 300 python bench.py > gpurun_out/s3z_bench_msc.log 2>gpurun_out/s3z_bench_msc.err; echo "bench rc $?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3z_bench_ref.log 2>&1
timeout 120 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/s3z_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3z_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/s3z_ncu_launches.log 2>&1
timeout 120 python bench.py --sets 1200 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/s3z_plain_enc.log 2>&1 && timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_msc_encoder_tc -s 3 -c 1 -o gpurun_out/s3z_prof_enc python bench.py --sets 1200 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/s3z_ncu_enc.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3z_smoke.log 2>&1; tail -1 gpurun_out/s3z_smoke.log
