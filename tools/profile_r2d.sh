#!/bin/bash
# Round 2, pass d (TWO GPUs): same-process peer-halo probe, the three failing tests with diagnostics, the multi-GPU test,
# the N=2 bench (peer halo without events), Toeplitz probe rerun (fp32 outputs for the bf16 variants).
set -u
O=gpurun_out
mkdir -p $O
timeout 120 ./tools/probe_peer_halo > $O/r02d_probe_peer_halo.txt 2>&1; tail -30 $O/r02d_probe_peer_halo.txt
(timeout 600 python -m pytest tests/test_fir_gpu.py -m gpu -q -rf -k "regression_fixtures or short_float" 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-600) > $O/r02d_pytest_failing.log 2>&1; cat $O/r02d_pytest_failing.log
(timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rf 2>&1 | tail -15 | cut -c1-400) > $O/r02d_pytest_multigpu.log 2>&1; cat $O/r02d_pytest_multigpu.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > $O/r02d_bench_n2.log 2>&1; tail -4 $O/r02d_bench_n2.log | cut -c1-700
timeout 600 python tools/probe_toeplitz_gemm.py > $O/r02_probe_toeplitz_gemm.jsonl 2> $O/r02_probe_toeplitz_gemm.err; cat $O/r02_probe_toeplitz_gemm.jsonl | cut -c1-330; tail -3 $O/r02_probe_toeplitz_gemm.err
