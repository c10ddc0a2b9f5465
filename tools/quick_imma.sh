#!/bin/bash
# GPU check of the int8 tensor-core FIR: parity tests, then bench lines (args: workloads)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_fir_gpu.py -m gpu -x -q -k "imma or tap_counts or wrapping or factory or baseline or streaming or host_buffer or fixtures" 2>&1 | tail -15
for w in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$w', round(d['value']), 'Msamp/s', 'frac', round(r['frac'],3), r.get('kernel'), 'kernel_ms', round(r.get('kernel_ms',0),3))"
done
