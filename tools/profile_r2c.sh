#!/bin/bash
# Round 2, pass c (ONE GPU): full GPU suite with failure summary, default bench line, Toeplitz-GEMM probe, ncu of the lazy-wrap Q15 FFT.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 1200 python -m pytest tests -m gpu -q -rf --deselect tests/test_multigpu_gpu.py 2>&1 | tail -60) > $O/r02c_pytest_gpu.log 2>&1; tail -40 $O/r02c_pytest_gpu.log | cut -c1-400
(time timeout 600 python bench.py --steps 20 --warmup 5) > $O/r02c_bench_default.log 2>&1; tail -4 $O/r02c_bench_default.log | cut -c1-400
timeout 600 python tools/probe_toeplitz_gemm.py > $O/r02_probe_toeplitz_gemm.jsonl 2> $O/r02_probe_toeplitz_gemm.err; cat $O/r02_probe_toeplitz_gemm.jsonl | cut -c1-400; tail -3 $O/r02_probe_toeplitz_gemm.err
./tools/ncu_cap.sh r02c_prof_fft4096_i16_lazy fft4096 c4_i16
python tools/ncu_summary.py $O/r02c_prof_fft4096_i16_lazy.ncu-rep > $O/r02c_prof_fft4096_i16_lazy.txt; head -12 $O/r02c_prof_fft4096_i16_lazy.txt
