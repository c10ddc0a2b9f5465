#!/bin/bash
# round 2, pass ai: full GPU suite, default bench line, int16 workloads after the tcgen05 issue / wait changes, ncu capture of fir_umma32t_kernel
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02ai_pytest_gpu.log 2>&1
tail -3 $O/r02ai_pytest_gpu.log
timeout 600 python bench.py > $O/r02ai_bench_default.log 2>&1
grep '^{' $O/r02ai_bench_default.log > $O/r02ai_bench_default.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ai_bench_default.jsonl').readline())
print('headline', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d.get('configs',{}).items():
    r=v.get('roofline') or {}
    print(k, v.get('value'), r.get('frac'), r.get('kernel'), v.get('parity'))
PY
for wl in c3_i16 real64_i16 c1_real; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $wl > $O/r02ai_$wl.log 2>&1
grep '^{' $O/r02ai_$wl.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$wl', d['roofline']['kernel'], d['value'], d['roofline']['frac'], d['parity'])"
done
bash tools/ncu_cap_env.sh umma32t r02ai_umma32t_c2 fir_umma32t c2
python tools/ncu_summary.py $O/r02ai_umma32t_c2.ncu-rep > $O/r02ai_prof_umma32t_c2.txt 2>&1
head -12 $O/r02ai_prof_umma32t_c2.txt
