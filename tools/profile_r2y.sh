#!/bin/bash
# round 2, pass y: fir_umma32t_kernel with the 16-lane tensor-memory loads in the epilogue
set -u
O=gpurun_out
mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_tmem_ld tools/probe_tmem_ld_shapes.cu && timeout 60 /tmp/probe_tmem_ld > $O/r02y_probe_tmem_ld_shapes.txt 2>&1
tail -1 $O/r02y_probe_tmem_ld_shapes.txt
timeout 300 python tools/probe_u32t_tiles.py > $O/r02y_probe_tiles.log 2>&1
grep -c "equal=True" $O/r02y_probe_tiles.log; grep "TIMEOUT\|rc=1\|equal=False" $O/r02y_probe_tiles.log | head
if grep -q "TIMEOUT\|rc=1\|equal=False" $O/r02y_probe_tiles.log; then exit 1; fi
timeout 600 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "umma32 or unaligned" > $O/r02y_pytest.log 2>&1
tail -3 $O/r02y_pytest.log
for algo in umma32 umma32t; do
B200C_FIR_ALGO=$algo timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02y_c2_$algo.log 2>&1
grep '^{' $O/r02y_c2_$algo.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
B200C_UMMA_DBG=1 B200C_FIR_ALGO=umma32t timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02y_c2_dbg.log 2>&1
grep -i "umma32:" $O/r02y_c2_dbg.log | tail -2 | cut -c1-400
for k in 32 64 200; do
for algo in umma32 umma32t; do
B200C_FIR_ALGO=$algo timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 --ntaps $k > $O/r02y_c2_${algo}_$k.log 2>&1
grep '^{' $O/r02y_c2_${algo}_$k.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('K=$k', d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
done
