#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_fir_gpu.py -m gpu -x -q -k "ummap" 2>&1 | tail -25
for w in "$@"; do ./tools/benchval.sh $w; done
