#!/bin/bash
# compute-sanitizer passes over a small batched run (memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards in the island, velocity and position kernels).  Run on a GPU box:
#   bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
cd "$(dirname "$0")/.."
cat > /tmp/b2g_sanitize_case.py <<'PY'
import sys
sys.path.insert(0, ".")
from box2d_rs_b200 import scenes, world
for name, build, g in (("pyramid", scenes.pyramid, (0.0, -10.0)), ("variety", scenes.variety, (0.0, -10.0))):
    w = world.B2world(g)
    build(w)
    b = w.batch(40)
    b.step(scenes.DT, 8, 3, 45)
    st = b.stats()
    print(name, "contacts", int(st["contacts"][0]), "status", set(st["status"].tolist()))
    b.close()
    w.close()
PY
for tool in memcheck racecheck; do
  echo "==== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 200 python /tmp/b2g_sanitize_case.py 2>&1 | tail -60
done
