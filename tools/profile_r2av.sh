#!/bin/bash
# round 2, pass av: last single-GPU pass over the final binary: full GPU suite (incl. smoke), default bench line
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02av_pytest_gpu.log 2>&1
tail -2 $O/r02av_pytest_gpu.log
timeout 600 python bench.py > $O/r02av_bench_default.log 2>&1
grep '^{' $O/r02av_bench_default.log > $O/r02av_bench_default.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02av_bench_default.jsonl').readline())
print('headline', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for v in d.get('configs',[]):
    r=v.get('roofline') or {}
    print(v.get('workload','')[:40], round(v.get('value')), round(r.get('frac'),4), r.get('kernel'), (v.get('e2e') or {}).get('value'))
PY
