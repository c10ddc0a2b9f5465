#!/bin/bash
# Round 2, pass k (ONE GPU): evidence pass -- launch list of the default bench command, full ncu captures of the kernels
# behind the headline and the configs that changed this round, racecheck + memcheck of the kernel families.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02k_launches_default.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1; tail -2 $O/ncu_launches.log | cut -c1-200
./tools/ncu_cap.sh r02k_prof_os32_headline fir_os32_kernel headline
python tools/ncu_summary.py $O/r02k_prof_os32_headline.ncu-rep > $O/r02k_prof_os32_headline.txt; head -4 $O/r02k_prof_os32_headline.txt
./tools/ncu_cap.sh r02k_prof_os32x_c3 fir_os32x c3
python tools/ncu_summary.py $O/r02k_prof_os32x_c3.ncu-rep > $O/r02k_prof_os32x_c3.txt; head -4 $O/r02k_prof_os32x_c3.txt
./tools/ncu_cap.sh r02k_prof_umma32_c2 fir_umma32 c2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_os64p -s 3 -c 1 -f -o $O/r02k_prof_os64p_bank \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload c5_bank --channels 296 --log2-samples 20 > $O/ncu_r02k_prof_os64p_bank.log 2>&1
python tools/ncu_summary.py $O/r02k_prof_os64p_bank.ncu-rep > $O/r02k_prof_os64p_bank.txt; head -12 $O/r02k_prof_os64p_bank.txt
(time timeout 1200 python -m pytest tests -m gpu -q -rf 2>&1 | tail -5) > $O/r02k_pytest_gpu.log 2>&1; cat $O/r02k_pytest_gpu.log | cut -c1-300
(time timeout 600 python bench.py --steps 20 --warmup 5) > $O/r02k_bench_default.log 2>&1; tail -4 $O/r02k_bench_default.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02k_bench_reference.log 2>&1; tail -1 $O/r02k_bench_reference.log | cut -c1-300
python tools/ncu_summary.py $O/r02k_prof_umma32_c2.ncu-rep > $O/r02k_prof_umma32_c2.txt; head -4 $O/r02k_prof_umma32_c2.txt
timeout 1500 compute-sanitizer --tool racecheck --target-processes all python tools/sanitize_small.py > $O/r02k_sanitize_racecheck.log 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/r02k_sanitize_racecheck.log | sort | uniq -c | head
timeout 1500 compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_small.py > $O/r02k_sanitize_memcheck.log 2>&1; grep -E "ERROR SUMMARY" $O/r02k_sanitize_memcheck.log | sort | uniq -c
