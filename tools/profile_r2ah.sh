#!/bin/bash
# round 2, pass ah: fir_umma32t_kernel variants at C2: stager warps (4 / 6 / 8) x epilogue (0: three chunks, 1: accumulators loaded at once)
set -u
O=gpurun_out
mkdir -p $O
for sw in 4 6 8; do for epi in 0 1; do
B200C_U32T_SW=$sw B200C_U32T_EPI=$epi timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02ah_c2_${sw}_${epi}.log 2>&1
grep '^{' $O/r02ah_c2_${sw}_${epi}.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('stager warps $sw epilogue $epi', d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done; done
for sw in 4 8; do
B200C_UMMA_DBG=1 B200C_U32T_SW=$sw B200C_U32T_EPI=0 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02ah_c2_dbg_$sw.log 2>&1
grep -i "umma32:" $O/r02ah_c2_dbg_$sw.log | tail -2 | cut -c1-400
done
