#!/bin/bash
# round 2, pass al: final single-GPU evidence: full GPU suite, smoke(), default bench line, reference arm, ncu launch list of the default bench
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02al_pytest_gpu.log 2>&1
tail -2 $O/r02al_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02al_smoke.log 2>&1; tail -1 $O/r02al_smoke.log
timeout 600 python bench.py > $O/r02al_bench_default.log 2>&1
grep '^{' $O/r02al_bench_default.log > $O/r02al_bench_default.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02al_bench_default.jsonl').readline())
print('headline', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for v in d.get('configs',[]):
    r=v.get('roofline') or {}
    print(v.get('workload','')[:40], round(v.get('value')), round(r.get('frac'),4), r.get('kernel'), (v.get('e2e') or {}).get('value'))
    for row in v.get('per_buffer_size') or []: print('   ', row['work_buffer_mib'], 'MiB', round(row['value']), row['us_per_work'], round(row['hbm_frac'],3))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02al_bench_reference.log 2>&1; grep '^{' $O/r02al_bench_reference.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02al_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02al_ncu_bench.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02al_launches_default.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    k=r[ik].split('(')[0]; agg[k][0]+=1; agg[k][1]+=v
with open('gpurun_out/r02al_launches_default_summary.csv','w') as f:
    f.write('"kernel","launches","total_us"\n')
    for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): f.write(f'"{k}",{n},{t/1000:.1f}\n')
print(open('gpurun_out/r02al_launches_default_summary.csv').read()[:1500])
PY
