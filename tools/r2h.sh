#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -q -k "attention" 2>&1 | tail -8 > gpurun_out/r2h_attn.log
timeout 200 python tools/run_stage.py --stage 2 --sets 592 > gpurun_out/r2h_stage2.log 2>&1
for v in "0,0,0,0" "0,500,0,0" "0,500,50,0" "20,500,50,20" "50,500,100,50" "0,500,100,50" "50,500,100,0"; do
  echo "== PAUT_ENC_SLEEP=$v" >> gpurun_out/r2h_enc_sleep.log
  PAUT_ENC_SLEEP=$v PAUT_ATTN_MMA=1 timeout 200 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['kernels_ms_per_step'])" >> gpurun_out/r2h_enc_sleep.log
done
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 > gpurun_out/r2h_bench_msc_tc.log 2>&1
timeout 200 python tools/run_stage.py --stage 1 --sets 2000 > gpurun_out/r2h_stage1.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ts_encoder -c 1 -o gpurun_out/r2h_ts_encoder python tools/run_stage.py --stage 1 --sets 2000 --reps 1 > gpurun_out/r2h_ncu.log 2>&1
echo done
