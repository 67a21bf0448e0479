#!/bin/bash
# tuning sweep behind profiles/r02_stream_groups.md: stream groups x stagger x CUDA graphs, headline workload
for cfg in "8 1 none" "8 1 no_graph" "4 1 no_graph" "16 1 no_graph" "8 0 no_graph"; do
  set -- $cfg
  extra=""; [ "$3" != "none" ] && extra="--solver $3"
  B2GPU_STREAM_GROUPS=$1 B2GPU_STAGGER=$2 timeout 200 python bench.py --no-cpu --no-t0 --no-single-world $extra > gpurun_out/sweep_$1_$2_$3.json 2>gpurun_out/sweep_err.log
  python - <<PY
import json
d=json.load(open('gpurun_out/sweep_$1_$2_$3.json'))
print("groups $1 stagger $2 $3: value %.4g ms %.4f | e2e %.4g ms %.4f" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
PY
done
