#!/bin/bash
# Runs bench.py the way the driver does at N = 1, 2, 4, 8 on one box (plus the reference arm) and leaves the
# JSON lines under gpurun_out/.  Usage (8-GPU box): gpurun --gpus 8 -- bash tools/scale_run.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
port=29511
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py \
    --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  port=$((port + 1))
done
python - <<'PY'
import json
for n in ("ref", "1gpu", "2gpu", "4gpu", "8gpu"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d.get("clocks"))
    except Exception as e:
        print(n, "failed:", e)
PY
