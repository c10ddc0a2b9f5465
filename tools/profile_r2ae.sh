#!/bin/bash
# round 2, pass ae: four stager warps with compile-time sizes, unrolled epilogue: C2 A/B + one full ncu capture
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_fir_gpu.py -x -q -m gpu -k "umma32" > $O/r02ae_pytest.log 2>&1
tail -2 $O/r02ae_pytest.log
for algo in umma32 umma32t; do
B200C_FIR_ALGO=$algo timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02ae_c2_$algo.log 2>&1
grep '^{' $O/r02ae_c2_$algo.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
B200C_UMMA_DBG=1 B200C_FIR_ALGO=umma32t timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02ae_c2_dbg.log 2>&1
grep -i "umma32:" $O/r02ae_c2_dbg.log | tail -2 | cut -c1-400
bash tools/ncu_cap_env.sh umma32t r02ae_umma32t_c2 fir_umma32t c2
python tools/ncu_summary.py $O/r02ae_umma32t_c2.ncu-rep > $O/r02ae_prof_umma32t_c2.txt 2>&1
head -50 $O/r02ae_prof_umma32t_c2.txt
for k in 32 64 200; do
B200C_FIR_ALGO=umma32t timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 --ntaps $k > $O/r02ae_c2_t_$k.log 2>&1
grep '^{' $O/r02ae_c2_t_$k.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('K=$k', d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
