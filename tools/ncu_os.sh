#!/bin/bash
# one full ncu capture of the overlap-save FIR kernel (2^26 samples), output name = $1
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_os -s 3 -c 1 -f -o $O/${1:-prof_os} \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload ${2:-headline} > $O/ncu_os.log 2>&1
tail -2 $O/ncu_os.log
