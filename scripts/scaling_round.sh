#!/bin/bash
# 1 -> 8 GPU weak-scaling of the bench on ONE 8-GPU box (run with: gpurun --gpus 8 -- bash scripts/scaling_round.sh)
mkdir -p gpurun_out
run() {  # n workload
  if [ "$1" = 1 ]; then timeout 600 python bench.py --gpus 1 --workload $2 --no-cpu-baseline
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --workload $2 --no-cpu-baseline
  fi
}
for n in 1 2 4 8; do run $n pp_sweep 2>gpurun_out/scale_pp_$n.err | tail -1 > gpurun_out/scale_pp_$n.json; echo "pp N=$n rc=$?"; cut -c1-220 gpurun_out/scale_pp_$n.json; done
for n in 8; do run $n stanley_lqg 2>gpurun_out/scale_stanley_$n.err | tail -1 > gpurun_out/scale_stanley_$n.json; echo "stanley N=$n rc=$?"; cut -c1-220 gpurun_out/scale_stanley_$n.json; done
