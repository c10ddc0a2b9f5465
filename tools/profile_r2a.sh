#!/bin/bash
# Round 2, pass a (ONE GPU): GPU suite, the default bench line (headline + every BASELINE config), the reference
# arm, and the ncu capture round 1 never took (int16 FFT).  Numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/r02a_pytest_gpu.log 2>&1; tail -8 $O/r02a_pytest_gpu.log
(time timeout 600 python bench.py --steps 20 --warmup 5) > $O/r02a_bench_default.log 2>&1; tail -5 $O/r02a_bench_default.log | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02a_bench_reference.log 2>&1; tail -1 $O/r02a_bench_reference.log | cut -c1-600
./tools/ncu_cap.sh r02a_prof_fft4096_i16 fft4096 c4_i16
python tools/ncu_summary.py $O/r02a_prof_fft4096_i16.ncu-rep > $O/r02a_prof_fft4096_i16.txt; cat $O/r02a_prof_fft4096_i16.txt
ls -la $O | tail -8
