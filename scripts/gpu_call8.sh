#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 "$@" 2>&1 | tail -30; }
run python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 300 -p no:cacheprovider
run python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider
for t in test_eval_forward_vs_golden 'test_train_forward_backward_vs_oracle[False]' 'test_train_forward_backward_vs_oracle[True]' test_reference_loop_and_fused_trainer_agree test_dropout_on_matches_oracle_with_same_masks test_full_size_properties; do
  run python -m pytest "tests/test_gpu_tbsrn.py::$t" -m gpu -q --timeout 900 -p no:cacheprovider
  cp gpurun_out/tbsrn_parity.json "gpurun_out/parity_$t.json" 2>/dev/null
done
echo "=== bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
echo "=== ncu full on hot kernels"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"attn_|tc_gemm|linear_wgrad_kernel|conv_wgrad_kernel" -s 6 -c 6 -o gpurun_out/hot_r01 -f python scripts/prof_ops.py > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
