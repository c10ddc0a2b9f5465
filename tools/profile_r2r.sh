#!/bin/bash
# Round 2, pass r (ONE GPU): spectral resampler with 13 / 14 warps per SM at 128 registers.
set -u
O=gpurun_out
mkdir -p $O
for v in 13 14; do
  (B200C_OSX_WARPS=$v timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -k "spectral_resampler" 2>&1 | tail -2 | cut -c1-300) > $O/r02r_pytest_$v.log 2>&1; cat $O/r02r_pytest_$v.log
done
for v in 12 13 14; do
  for args in "--workload c3" "--workload resamp_short" "--workload c3 --log2-samples 30"; do
    B200C_OSX_WARPS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $args > $O/r02r_bench.log 2>&1
    python - <<PY
import json
for l in open("$O/r02r_bench.log"):
    if l.startswith("{"):
        d = json.loads(l); print("warps=$v $args", round(d["value"]), "%.4f" % d["roofline"]["frac"])
PY
  done
done
