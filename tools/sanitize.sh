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
  [ "${1:-}" = "large" ] && break
  echo "==== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 200 python /tmp/b2g_sanitize_case.py 2>&1 | tail -60
done

# large-world mode (b2g_large.h): global-memory kernels only, so memcheck + initcheck (racecheck covers shared memory).
# Run alone with:  bash tools/sanitize.sh large
if [ "${1:-}" = "large" ] || [ "${1:-}" = "all" ]; then
  for tool in memcheck initcheck; do
    for sc in "pile 400 40 1" "addpair 1500 60 1" "variety 0 120 2"; do
      set -- $sc
      echo "==== compute-sanitizer --tool $tool  large-world mode $4: $1 n=$2 steps=$3"
      timeout 600 compute-sanitizer --tool $tool --print-limit 50 python tools/large_smoke.py --scene $1 --n $2 --steps $3 --mode $4 2>&1 | tail -8
    done
  done
fi
