#!/bin/bash
# ncu captures of the config-5 kinds: launch #6 = pure pursuit, #4 = stanley+lqg+dynamic, #5 = stanley+pid (first timed step)
mkdir -p gpurun_out
for spec in "mc 5" "mc_stlqg 3" "mc_stpid 4"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel_f32 -s $2 -c 1 -o gpurun_out/prof_$1 -f \
    python bench.py --workload skidpad_mc --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
done
