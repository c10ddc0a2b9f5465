#!/bin/bash
# Round 2, pass l (ONE GPU): fir_os128_kernel (128 threads x 32 points per 4096-point transform) against fir_os64_kernel.
set -u
O=gpurun_out
mkdir -p $O
for v in 4 5; do
(B200C_OS128=$v timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -k "overlap_save or filter_bank or tap_counts" 2>&1 | tail -4 | cut -c1-300) > $O/r02l_pytest_os128_$v.log 2>&1; cat $O/r02l_pytest_os128_$v.log
done
for v in 0 3 4 5; do
  for w in c5 c5_bank; do
    B200C_OS128=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/r02l_bench_${w}_v$v.log 2>&1
    python - <<PY
import json
for l in open("$O/r02l_bench_${w}_v$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("os128=$v $w", round(d["value"]), "%.4f" % d["roofline"]["frac"], d["roofline"]["kernel"])
PY
  done
done
B200C_OS128=4 ./tools/ncu_cap.sh r02l_prof_os128_c5 fir_os128 c5
python tools/ncu_summary.py $O/r02l_prof_os128_c5.ncu-rep > $O/r02l_prof_os128_c5.txt; cat $O/r02l_prof_os128_c5.txt
