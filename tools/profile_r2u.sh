#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
B200C_UMMA_DBG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c2 > $O/r02u_c2_dbg.log 2>&1
grep -i "umma32" $O/r02u_c2_dbg.log | tail -4 | cut -c1-400
for k in 32 64 256; do
B200C_UMMA_DBG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c2 --ntaps $k > $O/r02u_c2_dbg_$k.log 2>&1
grep -i "umma32" $O/r02u_c2_dbg_$k.log | tail -1 | cut -c1-400
done
