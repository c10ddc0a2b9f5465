#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
(time timeout 1200 python -m pytest tests -m gpu -q -rf 2>&1 | tail -6) > $O/r02t_pytest_gpu.log 2>&1; cat $O/r02t_pytest_gpu.log | cut -c1-300
