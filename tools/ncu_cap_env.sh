#!/bin/bash
# one full ncu capture with B200C_FIR_ALGO set: ncu_cap_env.sh <algo> <out name> <kernel regex> <workload>
O=gpurun_out; mkdir -p $O
B200C_FIR_ALGO=$1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -f -o $O/$2 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload $4 > $O/ncu_$2.log 2>&1
tail -2 $O/ncu_$2.log
