#!/bin/bash
# Round 2, pass p (ONE GPU): os32x with the next block fetched after the first inverse round's reads; os64p parts penalty.
set -u
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -k "spectral_resampler or filter_bank or overlap_save" 2>&1 | tail -3 | cut -c1-300) > $O/r02p_pytest.log 2>&1; cat $O/r02p_pytest.log
for args in "--workload c3" "--workload resamp_short" "--workload c3 --log2-samples 30" "--workload c5_bank" "--workload c5_bank --channels 128"; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $args > $O/r02p_bench.log 2>&1
    python - <<PY
import json
for l in open("$O/r02p_bench.log"):
    if l.startswith("{"):
        d = json.loads(l); print("$args", round(d["value"]), "%.4f" % d["roofline"]["frac"], d["roofline"]["kernel"])
PY
    grep -i "error" $O/r02p_bench.log | tail -2
done
./tools/ncu_cap.sh r02p_prof_os32x_c3 fir_os32x c3
python tools/ncu_summary.py $O/r02p_prof_os32x_c3.ncu-rep > $O/r02p_prof_os32x_c3.txt; cat $O/r02p_prof_os32x_c3.txt | grep -E "duration|pipe_fma_cycles|l1tex__throughput|stall"
