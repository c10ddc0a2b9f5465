#!/bin/bash
# Round 2, pass i (EIGHT GPUs): the bench line at N = 8 and N = 4 as the driver launches it, the reference arm,
# and the multi-GPU test.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02i_topo.txt 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5) > $O/r02i_bench_n8.log 2>&1; tail -4 $O/r02i_bench_n8.log | cut -c1-300
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5) > $O/r02i_bench_n4.log 2>&1; tail -4 $O/r02i_bench_n4.log | cut -c1-300
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 8 --steps 5 --warmup 1) > $O/r02i_bench_reference_n8.log 2>&1; tail -3 $O/r02i_bench_reference_n8.log | cut -c1-300
(timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -3) > $O/r02i_pytest_multigpu.log 2>&1; cat $O/r02i_pytest_multigpu.log
