#!/bin/bash
# Round-1 (session e) evidence pass on ONE GPU: the full GPU test suite, bench lines for every workload,
# the reference arm, the launch list of the default bench command, full ncu captures of the kernels
# this session changed.  Numbers printed by runs under ncu are never bench values; they only feed profiles/.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/r01e_pytest_gpu.log 2>&1; tail -6 $O/r01e_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r01e_bench_headline.log 2>&1; tail -1 $O/r01e_bench_headline.log
for w in c2 c3 c4; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w > $O/r01e_bench_$w.log 2>&1; tail -1 $O/r01e_bench_$w.log
done
for w in c1 c3_i16 c5 c5_bank short resamp_short real64 real64_i16 c4_i16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/r01e_bench_$w.log 2>&1; tail -1 $O/r01e_bench_$w.log
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r01e_bench_reference.log 2>&1; tail -1 $O/r01e_bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r01e_launches_headline.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
./tools/ncu_cap.sh r01e_prof_os32x_c3 fir_os32x c3
python tools/ncu_summary.py $O/r01e_prof_os32x_c3.ncu-rep > $O/r01e_prof_os32x_c3.txt
./tools/ncu_cap.sh r01e_prof_ummap_c3i16 fir_ummap c3_i16
python tools/ncu_summary.py $O/r01e_prof_ummap_c3i16.ncu-rep > $O/r01e_prof_ummap_c3i16.txt
cat $O/r01e_bench_*.log | grep '^{' > $O/r01e_bench_lines.jsonl
ls -la $O | tail -12
