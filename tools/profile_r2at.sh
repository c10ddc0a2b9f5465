#!/bin/bash
# Round 2, pass at (EIGHT GPUs): final multi-GPU lines with the final code (dependent launches of fir_os32_kernel and fir_os32x_kernel): N = 8 and 2, two-GPU tests
set -u
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -3) > $O/r02at_pytest_multigpu.log 2>&1; cat $O/r02at_pytest_multigpu.log
for n in 8 2; do
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 20 --warmup 5) > $O/r02at_bench_n$n.log 2>&1
grep '^{' $O/r02at_bench_n$n.log > $O/r02at_bench_n$n.jsonl
python - $n <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02at_bench_n{n}.jsonl').readline())
print('N',n,'headline', round(d['value']), d['roofline']['frac'], 'ms', d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), d.get('halo'))
for v in d.get('configs',[]):
    r=v.get('roofline') or {}
    print('  ', v.get('workload','')[:40], round(v.get('value')), r.get('frac'), v.get('halo'), str(v.get('parity'))[:80])
PY
done
