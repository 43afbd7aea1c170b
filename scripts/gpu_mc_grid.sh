#!/bin/bash
mkdir -p gpurun_out
for pct in 100 50 25 100 50 25 12; do
  BGR_DYN_GRID_PCT=$pct python bench.py --workload skidpad_mc --steps 10 --warmup 3 --no-cpu-baseline --no-secondary | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pct $pct value %.4e  e2e %.4e' % (d['value'], d['e2e']['value']))"
done
