#!/bin/bash
# tap-count sweep, complex int16 L=M=1: direct IMAD kernel vs int8 tensor-core kernel
for k in "$@"; do
  for algo in direct imma; do
    v=$(B200C_FIR_ALGO=$algo timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload c2 --ntaps $k 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(round(d['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3))")
    echo "K=$k algo=$algo $v"
  done
done
