#!/bin/bash
# one full ncu capture: ncu_cap.sh <out name> <kernel regex> <workload> [skip]
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-3} -c 1 -f -o $O/$1 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload $3 > $O/ncu_$1.log 2>&1
tail -2 $O/ncu_$1.log
