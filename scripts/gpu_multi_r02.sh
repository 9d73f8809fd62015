#!/bin/bash
# round-2 two-GPU pass: the headline bench (weak + its strong point), the other BASELINE configs under torchrun, the reference arm
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, args...
  local name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" > gpurun_out/r02e_${name}_n$N.json 2> gpurun_out/r02e_${name}_n$N.err
  echo "=== $name rc=$? lines=$(wc -l < gpurun_out/r02e_${name}_n$N.json)"; cut -c1-260 gpurun_out/r02e_${name}_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version\|^$" gpurun_out/r02e_${name}_n$N.err | tail -3
}
run cfg2 --steps 10 --warmup 3
run cfg3 --config 3 --steps 5 --warmup 3 --no-cpu-baseline
run cfg4 --config 4 --steps 5 --warmup 3 --no-cpu-baseline
run cfg4_w32 --config 4 --width 32 --steps 5 --warmup 3 --no-cpu-baseline
run cfg5 --config 5 --steps 5 --warmup 3 --no-cpu-baseline
run ref --impl reference --steps 1 --warmup 0
