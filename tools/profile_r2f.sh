#!/bin/bash
# Round 2, pass f (ONE GPU): spectral resampler variants (12 warps vs 8 / 9 warps with landing buffers) -- parity then rate;
# host-path sweep (chunk size x pipeline depth); dispatch sweep (which kernel family wins where).
set -u
O=gpurun_out
mkdir -p $O
for v in 208 209; do
  (B200C_OSX_MINB=$v timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -k "spectral_resampler" 2>&1 | tail -3 | cut -c1-300) > $O/r02f_pytest_osx_$v.log 2>&1; echo "osx $v:"; cat $O/r02f_pytest_osx_$v.log
done
for v in 112 208 209; do
  for w in c3 resamp_short; do
    B200C_OSX_MINB=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/r02f_bench_${w}_osx$v.log 2>&1
    python - <<PY
import json
for l in open("$O/r02f_bench_${w}_osx$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("osx $v $w", round(d["value"]), "%.4f" % d["roofline"]["frac"], d["roofline"]["kernel"])
PY
  done
done
timeout 900 python tools/sweep_host_path.py > $O/r02_sweep_host_path.jsonl 2> $O/r02_sweep_host_path.err; cat $O/r02_sweep_host_path.jsonl | cut -c1-200; tail -2 $O/r02_sweep_host_path.err
timeout 900 python tools/sweep_dispatch.py > $O/r02_sweep_dispatch.jsonl 2> $O/r02_sweep_dispatch.err; cat $O/r02_sweep_dispatch.jsonl | cut -c1-330; tail -2 $O/r02_sweep_dispatch.err
