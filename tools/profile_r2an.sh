#!/bin/bash
# round 2, pass an: fir_os32x_kernel as a programmatic dependent: parity, C3 line
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_fir_gpu.py tests/test_blocks_gpu.py tests/test_entry_gpu.py -x -q -m gpu -k "back_to_back or resampler or spectral or chunk or stream or smoke or baseline_configs" > $O/r02an_pytest.log 2>&1
tail -2 $O/r02an_pytest.log
for pdl in 0 1; do
B200C_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --workload c3 > $O/r02an_c3_pdl$pdl.log 2>&1
grep '^{' $O/r02an_c3_pdl$pdl.log | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('pdl $pdl c3', d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
