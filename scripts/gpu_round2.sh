#!/bin/bash
# round-2 evidence: tests, smoke, bench (both arms), launch list of the bench command, ncu --set full of the hot kernels
mkdir -p gpurun_out
T=${1:-r02}
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
echo "=== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; tail -1 gpurun_out/${T}_ncu_bench.log | cut -c1-200; wc -l gpurun_out/${T}_launches.csv
echo "=== ncu full: attention"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_" -c 4 -o gpurun_out/${T}_hot_attn -f python scripts/one_step.py 256 1 > gpurun_out/${T}_ncu_full1.log 2>&1; tail -1 gpurun_out/${T}_ncu_full1.log
echo "=== ncu full: wgrad / bn"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv3x3_wgrad_tc_kernel|bn_bwd_reduce|bn_bwd_apply|bn_stats_kernel|bn_finalize|linear_wgrad_tc" -s 8 -c 12 -o gpurun_out/${T}_hot_misc -f python scripts/one_step.py 256 1 > gpurun_out/${T}_ncu_full2.log 2>&1; tail -1 gpurun_out/${T}_ncu_full2.log
ls -la gpurun_out/${T}_*.ncu-rep
