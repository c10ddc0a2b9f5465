#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_fir_gpu.py -m gpu -x -q -k "polyphase or baseline or streaming or burst" 2>&1 | tail -15
for w in "$@"; do ./tools/benchval.sh $w; done
