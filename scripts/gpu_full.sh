#!/bin/bash
# what the driver runs at round end, plus the profile captures
mkdir -p gpurun_out
echo "=== pytest -m gpu (single process, -x)"
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"
timeout 900 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1
echo "=== ncu launch list (one step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01.csv python scripts/one_step.py 256 2 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_r01.csv
echo "=== ncu full on hot kernels"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"attn_|tc_gemm|linear_wgrad_kernel|conv_wgrad_kernel" -s 7 -c 7 -o gpurun_out/hot_r01 -f python scripts/prof_ops.py > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
