set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc $?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_msc.log 2>gpurun_out/r1_bench_msc.err; echo "bench rc $?"
for m in two_stage ssd conv1d_msc enhanced msc_n; do python bench.py --model $m --steps 2 --warmup 3 --cpu-seconds 4 > gpurun_out/r1_bench_$m.log 2>&1; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.log 2>&1
tail -c 600 gpurun_out/r1_pytest.log
cut -c1-900 gpurun_out/r1_bench_msc.log
