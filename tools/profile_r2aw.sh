#!/bin/bash
# round 2, pass aw: full ncu captures of the final swapped tcgen05 kernels: complex (C2) and real int16
set -u
O=gpurun_out
mkdir -p $O
bash tools/ncu_cap_env.sh umma32t r02aw_umma32t_c2 fir_umma32t_kernel c2
python tools/ncu_summary.py $O/r02aw_umma32t_c2.ncu-rep > $O/r02aw_prof_umma32t_c2.txt 2>&1
bash tools/ncu_cap_env.sh umma32t r02aw_umma32tr_real64 fir_umma32tr real64_i16
python tools/ncu_summary.py $O/r02aw_umma32tr_real64.ncu-rep > $O/r02aw_prof_umma32tr_real64.txt 2>&1
grep -h "Kernel Name\|gpu__time\|dram__bytes\|dram_throughput\|issue_active\|tensor\|l1tex__throughput\|cycles_elapsed" $O/r02aw_prof_umma32t_c2.txt $O/r02aw_prof_umma32tr_real64.txt
