#!/bin/bash
# Multi-GPU check of the default bench command on ONE box: gpurun --gpus N -- bash scripts/scaling_round.sh "1 2"
# (each N runs the full default line: config 2 headline + configs 3 weak / 3 strong / 4 / 5 as secondary)
mkdir -p gpurun_out
for n in ${1:-1 2}; do
  if [ "$n" = 1 ]; then timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline
  fi 2> gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json
  echo "N=$n rc=$?"; cut -c1-260 gpurun_out/scale_$n.json; tail -3 gpurun_out/scale_$n.err
done
