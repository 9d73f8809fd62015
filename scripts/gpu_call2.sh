#!/bin/bash
# GPU call: op tests, whole-network parity (one process per test so a faulting kernel cannot poison the rest), bench
mkdir -p gpurun_out
nvidia-smi -L
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -40; }
run python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 300 -p no:cacheprovider
run python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider
for t in test_eval_forward_vs_golden 'test_train_forward_backward_vs_oracle[False]' 'test_train_forward_backward_vs_oracle[True]' test_reference_loop_and_fused_trainer_agree test_dropout_on_matches_oracle_with_same_masks test_full_size_properties; do
  run python -m pytest "tests/test_gpu_tbsrn.py::$t" -m gpu -q --timeout 600 -p no:cacheprovider
  cp gpurun_out/tbsrn_parity.json gpurun_out/parity_$t.json 2>/dev/null
done
echo "=== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
