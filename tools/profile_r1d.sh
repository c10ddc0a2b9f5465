#!/bin/bash
# Round-1 (session d) evidence pass on ONE GPU: bench lines for every workload, the launch list of
# the default bench command, full ncu captures of the int16 tensor-core kernels and the resampler.
# Numbers printed by runs under ncu are never bench values; they only feed profiles/.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r01d_bench_headline.log 2>&1; tail -1 $O/r01d_bench_headline.log
for w in c1 c1_real c2 c3 c3_i16 c5 c5_bank short short_cx resamp_short real64 real64_i16 c4 c4_i16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w > $O/r01d_bench_$w.log 2>&1; tail -1 $O/r01d_bench_$w.log
done
B200C_FIR_ALGO=imma timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r01d_bench_c2_imma.log 2>&1; tail -1 $O/r01d_bench_c2_imma.log
B200C_FIR_ALGO=umma timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r01d_bench_c2_umma16.log 2>&1; tail -1 $O/r01d_bench_c2_umma16.log
B200C_FIR_ALGO=direct timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu --no-e2e --workload c2 > $O/r01d_bench_c2_direct.log 2>&1; tail -1 $O/r01d_bench_c2_direct.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r01d_bench_reference.log 2>&1; tail -1 $O/r01d_bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r01d_launches_headline.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
./tools/ncu_cap.sh r01d_prof_umma32_c2 fir_umma32 c2
./tools/ncu_cap.sh r01d_prof_umma32_real64 fir_umma32 real64_i16
./tools/ncu_cap.sh r01d_prof_ummap_c3i16 fir_ummap c3_i16
ls -la $O | tail -30
