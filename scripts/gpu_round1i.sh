#!/bin/bash
# round-1 (pass i) evidence: full GPU test suite, smoke, headline bench (both arms), recogniser step timings with the weight
# gradients on the streaming kernel vs on tcgen05, ncu launch list of one recogniser step
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1100 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -25
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 600 python bench.py > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; tail -2 gpurun_out/bench_r01i.err; cut -c1-600 gpurun_out/bench_r01i.json
echo "=== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
echo "=== cfg4 sld, streaming wgrad"; FOCR_TC_WGRAD=0 timeout 200 python scripts/bench_cfg4.py --batch 64 --steps 5 2>&1 | tail -1 | tee gpurun_out/cfg4_sld_wgrad_legacy.json | cut -c1-900
echo "=== cfg4 sld, tcgen05 wgrad"; FOCR_TC_WGRAD=1 timeout 200 python scripts/bench_cfg4.py --batch 64 --steps 5 2>&1 | tail -1 | tee gpurun_out/cfg4_sld_wgrad_tc.json | cut -c1-900
echo "=== ids, tcgen05 wgrad"; FOCR_TC_WGRAD=1 timeout 200 python scripts/bench_cfg4.py --model ids --batch 64 --steps 5 2>&1 | tail -1 | tee gpurun_out/ids_wgrad_tc.json | cut -c1-900
echo "=== sld tests with tcgen05 wgrad"; FOCR_TC_WGRAD=1 timeout 300 python -m pytest tests/test_gpu_sld.py tests/test_gpu_ids.py -q -p no:cacheprovider 2>&1 | tail -3
echo "=== ncu launch list of the sld step"
FOCR_TC_WGRAD=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01i_sld.csv python scripts/bench_cfg4.py --batch 64 --steps 1 --warmup 3 > gpurun_out/ncu_sld.log 2>&1; tail -1 gpurun_out/ncu_sld.log | cut -c1-200; wc -l gpurun_out/launches_r01i_sld.csv
