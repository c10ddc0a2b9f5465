#!/bin/bash
# Round 2, pass b (TWO GPUs): the whole GPU suite without -x (incl. tests/test_multigpu_gpu.py), then the N=2 bench line
# (headline weak scaling with the peer-memory halo, c3 and c5_bank strong scaling) and the NCCL-halo form for comparison.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > $O/r02b_pytest_gpu.log 2>&1; tail -45 $O/r02b_pytest_gpu.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > $O/r02b_bench_n2.log 2>&1; tail -4 $O/r02b_bench_n2.log | cut -c1-1500
(B200C_BENCH_HALO=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --configs none --no-e2e) > $O/r02b_bench_n2_nccl.log 2>&1; tail -2 $O/r02b_bench_n2_nccl.log | cut -c1-600
