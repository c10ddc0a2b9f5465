#!/bin/bash
# round 2, pass az: ncu launch list of the default bench command, our kernels only (every BASELINE config of the line)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir_|fft|halo_|map_kernel|probe_kernel|table_source" -c 2000 --csv --log-file $O/r02az_launches_default_ours.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02az_ncu_bench.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02az_launches_default_ours.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    k=r[ik].split('(')[0]
    a=agg.setdefault(k,[0,0.0,1e30,0.0]); a[0]+=1; a[1]+=v; a[2]=min(a[2],v); a[3]=max(a[3],v)
with open('gpurun_out/r02az_launches_default_summary.csv','w') as f:
    f.write('"kernel","launches","total_us","min_us","max_us"\n')
    for k,(n,t,lo,hi) in agg.items(): f.write(f'"{k}",{n},{t/1000:.1f},{lo/1000:.1f},{hi/1000:.1f}\n')
print(open('gpurun_out/r02az_launches_default_summary.csv').read()[:3000])
PY
