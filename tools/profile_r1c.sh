#!/bin/bash
# Round-1 evidence pass (run under gpurun on ONE GPU): bench lines for every workload, launch
# list of the default bench command, full ncu captures of the kernels that carry the path:
#   fir_os32_kernel (headline / C1), fir_os64_kernel (C5), fir_os32g_kernel (C3 resampler),
#   fft4096_kernel (C4), fir_tile_kernel (C2, bit-exact int16).
# Numbers printed by runs under ncu are never bench values; they only feed profiles/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_headline.log 2>&1; tail -1 $O/bench_headline.log
for w in c1 c1_real c2 c3 c5 c5_bank short short_cx resamp_short real64 real64_i16 c4 c4_i16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w > $O/bench_$w.log 2>&1; tail -1 $O/bench_$w.log
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.log 2>&1; tail -1 $O/bench_reference.log
kill $SMI
./tools/peak_fma > $O/peak_fma.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_headline.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
cap() { # name kernel-regex skip workload
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/$1 \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload $4 > $O/ncu_$1.log 2>&1
}
cap prof_os32_headline fir_os32_kernel 3 headline
cap prof_os64_c5 fir_os64 3 c5
cap prof_os32g_c3 fir_os32g 3 c3
cap prof_fft4096_c4 fft4096 6 c4
cap prof_tile_c2 fir_tile 3 c2
ls -la $O
