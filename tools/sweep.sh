#!/bin/bash
# tap-count sweep, cf32 L=M=1: direct kernel vs fused overlap-save (1024 / 4096)
for k in "$@"; do
  for algo in direct fft; do
    for n in 1024 4096; do
      [ $algo = direct ] && [ $n = 4096 ] && continue
      v=$(B200C_FIR_ALGO=$algo B200C_OS_N=$n timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --ntaps $k 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(round(d['value']), d['roofline']['kernel'])")
      echo "K=$k algo=$algo N=$n $v"
    done
  done
done
