set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "msc" > gpurun_out/r11_pytest.log 2>&1; tail -2 gpurun_out/r11_pytest.log
timeout 120 python bench.py --sets 1200 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r11_plain_msc.log 2>&1 && timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_msc_encoder_tc -s 3 -c 1 -o gpurun_out/prof_enc5 python bench.py --sets 1200 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r11_ncu_msc.log 2>&1
tail -3 gpurun_out/r11_ncu_msc.log | cut -c1-200
timeout 120 python bench.py --model ssd --sets 2000 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r11_plain_ssd.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 6 -c 2 -o gpurun_out/prof_conv5 python bench.py --model ssd --sets 2000 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r11_ncu_ssd.log 2>&1
tail -3 gpurun_out/r11_ncu_ssd.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3
