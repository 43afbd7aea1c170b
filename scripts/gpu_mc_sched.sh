#!/bin/bash
# Config 5: schedules for the Stanley kinds (heavy-tailed episode lengths) next to the static default
mkdir -p gpurun_out
run() {  # <label> <env...>
  label=$1; shift
  env "$@" timeout 300 python bench.py --workload skidpad_mc --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_mc_$label.log 2>&1
  echo "$label rc=$?"; tail -1 gpurun_out/bench_mc_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   value %.4e  e2e %.4e (%.2f) launches %d' % (d['value'], d['e2e']['value'], d['e2e']['value']/d['value'], d['gpu_launches']))
except Exception as e: print('   (no json)', e)"
}
run static X=1
run dyn_stanley BGR_MC_DYNAMIC=1 BGR_MC_KINDS=1,2
run dyn_all BGR_MC_DYNAMIC=1
run split160_stanley BGR_MC_SPLIT=160 BGR_MC_KINDS=1,2
run split256_stanley BGR_MC_SPLIT=256 BGR_MC_KINDS=1,2
run serial BGR_MC_SERIAL=1
run serial_dyn_stanley BGR_MC_SERIAL=1 BGR_MC_DYNAMIC=1 BGR_MC_KINDS=1,2
