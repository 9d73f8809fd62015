#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -40; }
for t in test_eval_forward_vs_golden 'test_train_forward_backward_vs_oracle[False-4]' 'test_train_forward_backward_vs_oracle[True-64]' test_trainer_steps_and_full_size; do
  run python -m pytest "tests/test_gpu_tsrn.py::$t" -m gpu -q --timeout 900 -p no:cacheprovider
  cp gpurun_out/tsrn_parity.json "gpurun_out/tsrn_parity_$t.json" 2>/dev/null
done
