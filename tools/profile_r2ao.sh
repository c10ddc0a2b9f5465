#!/bin/bash
# round 2, pass ao: fir_ummap_kernel (C3 on complex int16) after the issue / wait changes: role timers and one full ncu capture
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c3_i16 > $O/r02ao_c3i16.log 2>&1
grep '^{' $O/r02ao_c3i16.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
B200C_UMMA_DBG=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --workload c3_i16 > $O/r02ao_c3i16_dbg.log 2>&1
grep -i "ummap" $O/r02ao_c3i16_dbg.log | grep -v '^{' | tail -3 | cut -c1-500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_ummap -s 3 -c 1 -f -o $O/r02ao_ummap_c3i16 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --log2-samples 26 --workload c3_i16 > $O/ncu_r02ao.log 2>&1
python tools/ncu_summary.py $O/r02ao_ummap_c3i16.ncu-rep > $O/r02ao_prof_ummap_c3i16.txt 2>&1
cat $O/r02ao_prof_ummap_c3i16.txt | head -60
