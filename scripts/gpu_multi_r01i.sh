#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "=== bench N=$N"
[ -n "$SKIP_HEADLINE" ] || timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r01i_n$N.json 2> gpurun_out/bench_r01i_n$N.err; tail -3 gpurun_out/bench_r01i_n$N.err; cut -c1-400 gpurun_out/bench_r01i_n$N.json
echo "=== sld N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/bench_cfg4.py --batch 64 --steps 5 2>&1 | grep '^{"metric' | tee gpurun_out/cfg4_sld_n$N.json | cut -c1-700
echo "=== ids N=$N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 scripts/bench_cfg4.py --model ids --batch 64 --steps 5 2>&1 | grep '^{"metric' | tee gpurun_out/ids_n$N.json | cut -c1-700
