#!/bin/bash
# Round 2, pass ay (FOUR GPUs): the N = 4 line with the final code
set -u
O=gpurun_out
mkdir -p $O
n=4
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29744 bench.py --gpus $n --steps 20 --warmup 5) > $O/r02ay_bench_n$n.log 2>&1
grep '^{' $O/r02ay_bench_n$n.log > $O/r02ay_bench_n$n.jsonl
python - $n <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02ay_bench_n{n}.jsonl').readline())
print('N',n,'headline', round(d['value']), d['roofline']['frac'], 'ms', d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), d.get('halo'))
for v in d.get('configs',[]):
    r=v.get('roofline') or {}
    print('  ', v.get('workload','')[:40], round(v.get('value')), r.get('frac'), v.get('halo'), str(v.get('parity'))[:80])
PY
