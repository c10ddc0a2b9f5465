#!/bin/bash
# Round-1 measurement pass (run under gpurun on ONE GPU).  Numbers printed by runs under ncu
# are never bench values; they only feed profiles/.
set -u
O=gpurun_out
mkdir -p $O
./tools/peak_fma > $O/peak_fma.jsonl 2>&1
for w in c4 c4_i16; do timeout 300 python bench.py --steps 5 --warmup 3 --workload $w > $O/bench_$w.log 2>&1; tail -1 $O/bench_$w.log; done
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_headline_full.log 2>&1; tail -1 $O/bench_headline_full.log
# launch list of the default bench command (every launch with its device time)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_headline.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
# full capture of the dominant kernel (one launch, after warm-up)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_tile_kernel -s 3 -c 1 -f -o $O/prof_fir_headline \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 > $O/ncu_fir.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_tile_kernel -s 3 -c 1 -f -o $O/prof_fir_c1 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload c1 > $O/ncu_fir_c1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_staged -s 6 -c 1 -f -o $O/prof_fft_c4 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload c4 > $O/ncu_fft.log 2>&1
ls -la $O
