#!/bin/bash
# round 2, pass aq: fir_umma32t_kernel on real int16 (64-output windows, SWIZZLE_64B planes): parity, real64_i16 against the original kernel
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "umma32t or unaligned or limb" > $O/r02aq_pytest.log 2>&1
tail -12 $O/r02aq_pytest.log | cut -c1-300
for algo in umma32 umma32t; do
B200C_FIR_ALGO=$algo timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload real64_i16 > $O/r02aq_real64_$algo.log 2>&1
grep '^{' $O/r02aq_real64_$algo.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
B200C_UMMA_DBG=1 B200C_FIR_ALGO=umma32t timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload real64_i16 > $O/r02aq_real64_dbg.log 2>&1
grep -i "umma32:" $O/r02aq_real64_dbg.log | tail -2 | cut -c1-400
