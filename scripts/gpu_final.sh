#!/bin/bash
# round-end evidence: tests, smoke, bench (both arms), launch list of the bench command, ncu --set full of the hot kernels
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cut -c1-400 gpurun_out/bench_final.json
echo "=== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
echo "=== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches_r01h.csv
echo "=== ncu full: attention"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_" -c 6 -o gpurun_out/hot_attn_r01h -f python scripts/one_step.py 256 1 > gpurun_out/ncu_full1.log 2>&1; tail -1 gpurun_out/ncu_full1.log
echo "=== ncu full: gemm / wgrad"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_gemm|linear_wgrad_tc|conv_wgrad" -s 6 -c 10 -o gpurun_out/hot_gemm_r01h -f python scripts/one_step.py 256 1 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
ls -la gpurun_out/*.ncu-rep | tail -3
