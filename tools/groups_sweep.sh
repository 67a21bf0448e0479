#!/bin/bash
# Tuning experiment: bench.py with different numbers of stream groups (B2GPU_STREAM_GROUPS).
for g in "$@"; do
  B2GPU_STREAM_GROUPS=$g python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null > /tmp/g.json
  python - "$g" <<'PY'
import json, sys
d = json.load(open('/tmp/g.json'))
print('groups', sys.argv[1], '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], '%.3f ms' % d['e2e']['ms_per_step'])
PY
done
