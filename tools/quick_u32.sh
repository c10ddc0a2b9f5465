#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 240 python -m pytest tests/test_fir_gpu.py -m gpu -x -q -k "umma32" 2>&1 | tail -25
for w in "$@"; do
  B200C_UMMA_DBG=1 B200C_FIR_ALGO=umma32 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w 2>&1 | grep -E "umma32:|metric" | tail -3 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('umma32:'): print(l.strip()); continue
    d=json.loads(l); r=d['roofline']
    print('$w', round(d['value']), 'Msamp/s', 'frac', round(r['frac'],3), r.get('kernel'), 'kernel_ms', round(r.get('kernel_ms',0),3))"
done
