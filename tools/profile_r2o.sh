#!/bin/bash
# Round 2, pass o (ONE GPU): fir_os64p_kernel with (channel, part) tasks: parity, the full bank, a 128-channel bank (what one of 8 GPUs holds).
set -u
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -k "overlap_save or filter_bank or tap_counts" 2>&1 | tail -3 | cut -c1-300) > $O/r02o_pytest.log 2>&1; cat $O/r02o_pytest.log
for args in "--workload c5" "--workload c5_bank" "--workload c5_bank --channels 128" "--workload c5_bank --channels 256" "--workload c5_bank --channels 512"; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $args > $O/r02o_bench.log 2>&1
    python - <<PY
import json
for l in open("$O/r02o_bench.log"):
    if l.startswith("{"):
        d = json.loads(l); print("$args", round(d["value"]), "%.4f" % d["roofline"]["frac"], d["roofline"]["kernel"])
PY
    grep -i "error" $O/r02o_bench.log | tail -2
done
