#!/bin/bash
# round 2, pass ak: programmatic dependent launch of fir_os32_kernel's persistent form: parity, the block-layer stream rows and the headline with and without
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_fir_gpu.py tests/test_blocks_gpu.py -x -q -m gpu -k "back_to_back or chunk or stream or seam or split or blocks or history or burst" > $O/r02ak_pytest.log 2>&1
tail -2 $O/r02ak_pytest.log
for pdl in 0 1; do
B200C_PDL=$pdl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload headline_blocks > $O/r02ak_blocks_pdl$pdl.log 2>&1
grep '^{' $O/r02ak_blocks_pdl$pdl.log | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('pdl $pdl', d['value'], json.dumps(d.get('rows') or d['config'].get('rows') or d.get('blocks') or '')[:600])"
B200C_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --configs none > $O/r02ak_headline_pdl$pdl.log 2>&1
grep '^{' $O/r02ak_headline_pdl$pdl.log | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('pdl $pdl headline', d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
