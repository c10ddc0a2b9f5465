#!/bin/bash
# GPU check of the tcgen05 FIR under a short timeout (a wrong descriptor must not hang the box)
# usage: quick_umma.sh <variant> workloads...
O=gpurun_out; mkdir -p $O
export B200C_UMMA_V=$1; shift
timeout 240 python -m pytest tests/test_fir_gpu.py -m gpu -x -q -k "umma" 2>&1 | tail -25
for w in "$@"; do
  B200C_FIR_ALGO=umma timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$w', round(d['value']), 'Msamp/s', 'frac', round(r['frac'],3), r.get('kernel'), 'kernel_ms', round(r.get('kernel_ms',0),3))"
done
