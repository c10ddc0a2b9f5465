#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for v in 0 1; do
    B200C_OS64P=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload c5_bank > $O/r02n_bench_c5_bank_p$v.log 2>&1
    python - <<PY
import json
for l in open("$O/r02n_bench_c5_bank_p$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("os64p=$v c5_bank", round(d["value"]), "%.4f" % d["roofline"]["frac"], d["roofline"]["kernel"], d["parity"])
PY
done
tail -3 $O/r02n_bench_c5_bank_p1.log | cut -c1-300
B200C_OS64P=1 timeout 600 python -m pytest tests/test_fir_gpu.py tests/test_blocks_gpu.py -m gpu -q 2>&1 | tail -3
