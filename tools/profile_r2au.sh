#!/bin/bash
# round 2, pass au: final single-GPU evidence of the round (after the misaligned-stream bulk path): full GPU suite (incl. smoke), sanitizers, default bench line
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02au_pytest_gpu.log 2>&1
tail -2 $O/r02au_pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --target-processes all python tools/sanitize_small.py > $O/r02au_sanitize_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY" $O/r02au_sanitize_racecheck.log | sort | uniq -c | head -3
timeout 600 compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_small.py > $O/r02au_sanitize_memcheck.log 2>&1
grep -E "ERROR SUMMARY" $O/r02au_sanitize_memcheck.log | sort | uniq -c
timeout 600 python bench.py > $O/r02au_bench_default.log 2>&1
grep '^{' $O/r02au_bench_default.log > $O/r02au_bench_default.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02au_bench_default.jsonl').readline())
print('headline', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for v in d.get('configs',[]):
    r=v.get('roofline') or {}
    print(v.get('workload','')[:40], round(v.get('value')), round(r.get('frac'),4), r.get('kernel'), (v.get('e2e') or {}).get('value'))
PY
