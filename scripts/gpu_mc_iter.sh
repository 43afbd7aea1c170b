#!/bin/bash
# Config-5 iteration on the GPU box: parity tests, per-kind kernel times with the direct search on / off, bench lines.
mkdir -p gpurun_out
if [ "${1:-tests}" = tests ]; then
  timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 --tb=short > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
  tail -15 gpurun_out/pytest.log | cut -c1-300
fi
for sd in 0 1; do
  BGR_SEARCH_DIRECT=$sd timeout 300 python scripts/diag_mc.py > gpurun_out/diag_mc_sd$sd.log 2>&1; echo "diag sd=$sd rc=$?"
  grep "noise+cones" gpurun_out/diag_mc_sd$sd.log | cut -c1-200
done
for w in skidpad_mc pp_replan; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$w.log 2>&1
  echo "bench $w rc=$?"; tail -1 gpurun_out/bench_$w.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   value %.4e  e2e %.4e (%.2f)' % (d['value'], d['e2e']['value'], d['e2e']['value']/d['value']))
except Exception as e: print('   (no json)', e)"
done
