#!/bin/bash
# Round 2, pass q (EIGHT GPUs): final multi-GPU lines with the round-2 kernels (N = 8, 4, 2, 1 on one box), smoke().
set -u
O=gpurun_out
mkdir -p $O
(time python -c "import __graft_entry__ as g; g.smoke()") > $O/r02q_smoke.log 2>&1; tail -3 $O/r02q_smoke.log | cut -c1-200
for n in 8 4 2; do
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5) > $O/r02q_bench_n$n.log 2>&1; tail -4 $O/r02q_bench_n$n.log | cut -c1-150
done
(time timeout 600 python bench.py --steps 20 --warmup 5) > $O/r02q_bench_n1.log 2>&1; tail -4 $O/r02q_bench_n1.log | cut -c1-150
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3) > $O/r02q_pytest_gpu.log 2>&1; cat $O/r02q_pytest_gpu.log
