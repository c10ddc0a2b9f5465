#!/bin/bash
# quick GPU check: FIR parity tests + a few bench lines (args: workloads)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$w', round(d['value']), 'Msamp/s', 'frac', round(r['frac'],3), 'kernel_ms', round(r.get('kernel_ms',0),3))"
done
