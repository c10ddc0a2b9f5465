#!/bin/bash
# usage: benchval.sh workload  -> prints "workload value frac"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$1', round(d['value']), 'frac', round(r['frac'],3))"
