#!/bin/bash
# Round-1 second measurement pass (run under gpurun on ONE GPU): parity tests, bench lines for
# every workload, launch list, and full ncu captures of the kernels that now carry the path
# (fir_os4096_kernel for cf32 L=M=1, fft4096_kernel for /comms/fft, fir_tile_kernel for the rest).
# Numbers printed by runs under ncu are never bench values; they only feed profiles/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_headline.log 2>&1; tail -1 $O/bench_headline.log
for w in c1 c1_real c2 c3 c5 c4 c4_i16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w > $O/bench_$w.log 2>&1; tail -1 $O/bench_$w.log
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.log 2>&1; tail -1 $O/bench_reference.log
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_headline.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_os4096 -s 3 -c 1 -f -o $O/prof_fir_os_headline \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 > $O/ncu_os.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft4096 -s 6 -c 1 -f -o $O/prof_fft4096_c4 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload c4 > $O/ncu_fft.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_tile -s 3 -c 1 -f -o $O/prof_fir_c3 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload c3 > $O/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_tile -s 3 -c 1 -f -o $O/prof_fir_c2 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload c2 > $O/ncu_c2.log 2>&1
ls -la $O
