#!/bin/bash
# Round-1 (session e, second pass) after the persistent shared-table forms of fir_os32_kernel / fir_os32r_kernel became
# the default: GPU suite, default bench line, the workloads those kernels serve, launch list, one ncu capture.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/r01e2_pytest_gpu.log 2>&1; tail -6 $O/r01e2_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r01e2_bench_headline.log 2>&1; tail -1 $O/r01e2_bench_headline.log
for w in c1 short real64 c5_bank; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/r01e2_bench_$w.log 2>&1; tail -1 $O/r01e2_bench_$w.log | cut -c1-400
done
B200C_OS32R_CFG=112 ./tools/benchval.sh real64
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r01e_launches_headline.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/ncu_launches.log 2>&1
./tools/ncu_cap.sh r01e_prof_os32_headline fir_os32_kernel headline
python tools/ncu_summary.py $O/r01e_prof_os32_headline.ncu-rep > $O/r01e_prof_os32_headline.txt
cat $O/r01e2_bench_*.log | grep '^{' > $O/r01e2_bench_lines.jsonl
