#!/bin/bash
# Round 2, pass m (ONE GPU): fir_os64p_kernel v2 (channel per CTA, tap spectrum + twiddles in shared memory) against fir_os64_kernel.
set -u
O=gpurun_out
mkdir -p $O
(B200C_OS64P=1 timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -k "overlap_save or filter_bank or tap_counts" 2>&1 | tail -4 | cut -c1-300) > $O/r02m_pytest_os64p.log 2>&1; cat $O/r02m_pytest_os64p.log
for v in 0 1; do
  for w in c5 c5_bank; do
    B200C_OS64P=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/r02m_bench_${w}_p$v.log 2>&1
    python - <<PY
import json
for l in open("$O/r02m_bench_${w}_p$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("os64p=$v $w", round(d["value"]), "%.4f" % d["roofline"]["frac"], d["roofline"]["kernel"])
PY
  done
done
B200C_OS64P=1 ./tools/ncu_cap.sh r02m_prof_os64p_c5 fir_os64p c5
python tools/ncu_summary.py $O/r02m_prof_os64p_c5.ncu-rep > $O/r02m_prof_os64p_c5.txt; cat $O/r02m_prof_os64p_c5.txt
