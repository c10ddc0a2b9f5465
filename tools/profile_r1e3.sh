#!/bin/bash
# Round-1 (session e, final pass): defaults = persistent shared-table kernels with early prefetch.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/r01e3_pytest_gpu.log 2>&1; tail -6 $O/r01e3_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r01e3_bench_headline.log 2>&1; tail -1 $O/r01e3_bench_headline.log | cut -c1-300
for w in c1 short real64; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/r01e3_bench_$w.log 2>&1; ./tools/benchval.sh $w
done
B200C_OS32R_CFG=1012 ./tools/benchval.sh real64
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r01e_launches_headline.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
./tools/ncu_cap.sh r01e_prof_os32_headline fir_os32_kernel headline
python tools/ncu_summary.py $O/r01e_prof_os32_headline.ncu-rep > $O/r01e_prof_os32_headline.txt
cat $O/r01e3_bench_*.log | grep '^{' > $O/r01e3_bench_lines.jsonl
