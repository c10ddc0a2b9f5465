#!/bin/bash
# Round 2, pass g (ONE GPU): lane-pair split of the spectral resampler's second inverse round: parity, then rate.
set -u
O=gpurun_out
mkdir -p $O
for v in 312 308; do
  (B200C_OSX_MINB=$v timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -k "spectral_resampler" 2>&1 | tail -3 | cut -c1-300) > $O/r02g_pytest_osx_$v.log 2>&1; echo "osx $v:"; cat $O/r02g_pytest_osx_$v.log
done
for v in 112 312 308; do
  for w in c3 resamp_short; do
    B200C_OSX_MINB=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/r02g_bench_${w}_osx$v.log 2>&1
    python - <<PY
import json
for l in open("$O/r02g_bench_${w}_osx$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("osx $v $w", round(d["value"]), "%.4f" % d["roofline"]["frac"], d["roofline"]["kernel"])
PY
  done
done
(time timeout 900 python -m pytest tests -m gpu -q -rf --deselect tests/test_multigpu_gpu.py 2>&1 | tail -8) > $O/r02g_pytest_gpu.log 2>&1; cat $O/r02g_pytest_gpu.log | cut -c1-300
