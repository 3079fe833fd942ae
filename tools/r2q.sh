#!/bin/bash
# round-2 verification pass: workspace fix, whole GPU suite, smoke, default bench + per-model lines, one ncu --set full
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/dbg_ws.py > gpurun_out/r2q_dbg.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2q_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2q_bench_default.log 2>&1
for m in two_stage enhanced ssd msc_n; do
  timeout 300 python bench.py --model $m --steps 5 --warmup 3 --cpu-seconds 0 --no-extra > gpurun_out/r2q_bench_$m.log 2>&1
done
timeout 300 python tools/ncu_capture.py > gpurun_out/r2q_capture_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'k_msc_encoder_tc|k_msc_attn_block|k_msc_ffn_head|k_ts_encoder|k_mscn_front|k_conv_tc|k_stem_flat' \
  -o gpurun_out/r2q_full -f python tools/ncu_capture.py > gpurun_out/r2q_ncu.log 2>&1
echo done
