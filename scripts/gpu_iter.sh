#!/bin/bash
# One kernel-iteration session on the GPU box: parity tests, short bench lines for every library variant present,
# then one `ncu --set full` capture (with source counters) of the two headline step kernels.
#   usage: gpurun -- bash scripts/gpu_iter.sh [tests|notests] [variant ...]      (NCU_VARIANTS="base cta3" to profile more)
mkdir -p gpurun_out
mode=${1:-tests}; shift
variants="base $*"
if [ "$mode" = tests ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --tb=short > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
  tail -40 gpurun_out/pytest.log | cut -c1-300
fi
for v in $variants; do
  lib=""; [ "$v" != base ] && lib="$PWD/bgr_pathplanning_control_b200/libbgr_rollout_$v.so"
  for w in pp_sweep stanley_lqg; do
    BGR_ROLLOUT_LIB=$lib timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary \
      > gpurun_out/bench_${w}_$v.log 2>&1
    echo "bench $w $v rc=$?"; tail -1 gpurun_out/bench_${w}_$v.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   value %.4e  e2e %.4e  kernel_ms %.3f  clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_avg'], d['clocks']))
except Exception as e: print('   (no json)', e)"
  done
done
for v in ${NCU_VARIANTS:-base}; do
  [ "$v" = none ] && continue
  lib=""; [ "$v" != base ] && lib="$PWD/bgr_pathplanning_control_b200/libbgr_rollout_$v.so"
  for w in pp_sweep stanley_lqg; do
    BGR_ROLLOUT_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 \
      -o gpurun_out/prof_${w}_$v -f python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-secondary \
      > gpurun_out/ncu_${w}_$v.log 2>&1
    echo "ncu $w $v rc=$?"
  done
done
ls -la gpurun_out | head -40
