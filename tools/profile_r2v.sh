#!/bin/bash
# round 2, pass v: the operand-swapped tcgen05 kernel (fir_umma32t_kernel): parity, then C2 against the original
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "umma32t or unaligned" > $O/r02v_pytest.log 2>&1
tail -5 $O/r02v_pytest.log
for algo in umma32 umma32t; do
B200C_FIR_ALGO=$algo timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02v_c2_$algo.log 2>&1
grep '^{' $O/r02v_c2_$algo.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
B200C_UMMA_DBG=1 B200C_FIR_ALGO=umma32t timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02v_c2_dbg.log 2>&1
grep -i "umma32:" $O/r02v_c2_dbg.log | tail -2 | cut -c1-400
for k in 32 64 200; do
B200C_FIR_ALGO=umma32t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 --ntaps $k > $O/r02v_c2_t_$k.log 2>&1
grep '^{' $O/r02v_c2_t_$k.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('K=$k', d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
B200C_FIR_ALGO=umma32 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 --ntaps $k > $O/r02v_c2_o_$k.log 2>&1
grep '^{' $O/r02v_c2_o_$k.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('K=$k', d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
