set -x
python bench.py --model ssd --sets 4000 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 4 -c 2 -o gpurun_out/prof_conv4 python bench.py --model ssd --sets 4000 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r5_ncu.log 2>&1
tail -2 gpurun_out/r5_ncu.log | cut -c1-200
