#!/bin/bash
# quick iteration: targeted tests + bench
mkdir -p gpurun_out
echo "=== tests: $1"
timeout 900 python -m pytest $1 -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15
echo "=== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['value'])
print(d['roofline']['breakdown_ms_per_step'])
PY
