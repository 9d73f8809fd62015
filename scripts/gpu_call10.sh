#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -40; }
run python -m pytest tests/test_gpu_crnn.py -m gpu -q --timeout 600 -p no:cacheprovider
run python __graft_entry__.py smoke
