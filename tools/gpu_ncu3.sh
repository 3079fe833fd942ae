set -x
timeout 120 python bench.py --model conv1d_msc --sets 400 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r19_plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 12 -c 4 -o gpurun_out/prof_gemm5 python bench.py --model conv1d_msc --sets 400 --steps 1 --warmup 3 --cpu-seconds 0 > gpurun_out/r19_ncu.log 2>&1
tail -2 gpurun_out/r19_ncu.log | cut -c1-200
