#!/bin/bash
# Round 2, pass h (ONE GPU): full GPU suite after the pruning, default bench line, memcheck of the default kernels.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 1200 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8) > $O/r02h_pytest_gpu.log 2>&1; cat $O/r02h_pytest_gpu.log | cut -c1-300
(time timeout 600 python bench.py --steps 20 --warmup 5) > $O/r02h_bench_default.log 2>&1; tail -4 $O/r02h_bench_default.log | cut -c1-300
timeout 1500 compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_small.py > $O/r02h_sanitize_memcheck.log 2>&1; tail -15 $O/r02h_sanitize_memcheck.log | cut -c1-300
