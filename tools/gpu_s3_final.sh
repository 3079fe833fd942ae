# Session-3 round-end measurement: GPU suite, bench lines of all ten models, reference arm, ncu launch list of the
# default bench, ncu --set full of the (changed) MSC encoder and of the hybrid model's conv + front-end kernels.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3f_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/s3f_pytest.log
timeout 300 python bench.py > gpurun_out/s3f_bench_msc.log 2>gpurun_out/s3f_bench_msc.err; echo "bench rc $?"
for m in msc_n two_stage ssd conv1d_msc enhanced msc_legacy improved hybrid complex; do timeout 300 python bench.py --model $m --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/s3f_bench_$m.log 2>&1; echo "$m rc $?"; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3f_bench_ref.log 2>&1
timeout 120 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/s3f_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3f_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/s3f_ncu_launches.log 2>&1
timeout 120 python bench.py --sets 1200 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/s3f_plain_enc.log 2>&1 && timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_msc_encoder_tc -s 3 -c 1 -o gpurun_out/s3f_prof_enc python bench.py --sets 1200 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/s3f_ncu_enc.log 2>&1
timeout 120 python bench.py --model improved --sets 600 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/s3f_plain_imp.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_bgsub_chanmean|k_chanmean_resample|k_conv_tc" -s 6 -c 2 -o gpurun_out/s3f_prof_imp python bench.py --model improved --sets 600 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/s3f_ncu_imp.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3f_smoke.log 2>&1; tail -1 gpurun_out/s3f_smoke.log
ls -la gpurun_out/*.ncu-rep | tail -3
