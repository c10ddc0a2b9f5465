#!/bin/bash
# Round 2, pass s (ONE GPU): host path with page-aligned chunks: parity of the host entry points, e2e of headline / c2 / c3 / c4.
set -u
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests -m gpu -q -k "host" 2>&1 | tail -2 | cut -c1-300) > $O/r02s_pytest.log 2>&1; cat $O/r02s_pytest.log
for w in headline c1 c2 c3 c4; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $w > $O/r02s_bench.log 2>&1
    python - <<PY
import json
for l in open("$O/r02s_bench.log"):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]; print("$w e2e", round(e["value"]), "h2d %.1f d2h %.1f GB/s" % (e["h2d_gbs"], e["d2h_gbs"]), "peak", e.get("memcpy_peak_gbs_each_way"))
PY
    grep -i "error" $O/r02s_bench.log | tail -2
done
