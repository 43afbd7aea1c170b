#!/bin/bash
# e2e (host buffers through bgr_rollout_host) against the host chunk size of the copy / compute pipeline
for c in 65536 131072 262144 524288; do
  BGR_HOST_CHUNK=$c timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/chunk_$c.json
  python -c "import json; d=json.load(open('/tmp/chunk_$c.json')); print($c, '%.4g' % d['value'], '%.4g' % d['e2e']['value'])"
done
