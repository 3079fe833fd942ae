set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r3_pytest.log
for m in ssd conv1d_msc two_stage; do python bench.py --model $m --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/r3_bench_$m.log 2>&1; done
PAUT_CONV_DEBUG=1 python tools/conv_probe.py ssd 2>&1 | grep "conv probe" | tail -3
python bench.py --model ssd --sets 4000 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 10 -c 2 -o gpurun_out/prof_conv3 python bench.py --model ssd --sets 4000 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r3_ncu.log 2>&1
tail -3 gpurun_out/r3_ncu.log | cut -c1-300
