#!/bin/bash
# Round 2, pass e (TWO GPUs): peer-halo probe after the event fix, previously failing tests, multi-GPU test, N=2 bench
# with the halo pulled by the one-CTA peer-load kernel (and the copy-engine form for comparison).
set -u
O=gpurun_out
mkdir -p $O
timeout 120 ./tools/probe_peer_halo > $O/r02e_probe_peer_halo.txt 2>&1; tail -12 $O/r02e_probe_peer_halo.txt
(timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -rf -k "regression_fixtures or short_float or untuned" 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-600) > $O/r02e_pytest_failing.log 2>&1; cat $O/r02e_pytest_failing.log
(timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rf 2>&1 | tail -8 | cut -c1-400) > $O/r02e_pytest_multigpu.log 2>&1; cat $O/r02e_pytest_multigpu.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > $O/r02e_bench_n2.log 2>&1; tail -4 $O/r02e_bench_n2.log | cut -c1-400
(B200C_HALO_MEMCPY=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --configs none --no-e2e) > $O/r02e_bench_n2_memcpy.log 2>&1; tail -1 $O/r02e_bench_n2_memcpy.log | cut -c1-400
