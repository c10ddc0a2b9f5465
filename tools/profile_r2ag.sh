#!/bin/bash
# round 2, pass ag: fir_umma32t_kernel: six stager warps, untimed fast-path waits, accumulators loaded at once
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "umma or imma or int16 or i16 or baseline_configs" > $O/r02ag_pytest.log 2>&1
tail -2 $O/r02ag_pytest.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02ag_c2.log 2>&1
grep '^{' $O/r02ag_c2.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
B200C_UMMA_DBG=1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02ag_c2_dbg.log 2>&1
grep -i "umma32:" $O/r02ag_c2_dbg.log | tail -2 | cut -c1-400
for k in 32 64 96 160 200 225; do
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 --ntaps $k > $O/r02ag_c2_t_$k.log 2>&1
grep '^{' $O/r02ag_c2_t_$k.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('K=$k', d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
