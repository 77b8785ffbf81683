#!/bin/bash
mkdir -p gpurun_out
for per in 4 8 16 32; do
  VLOAM_SR_CURV_TILES=$per timeout 300 python bench.py --steps 12 --warmup 4 --legs device > gpurun_out/curv_$per.json 2> gpurun_out/curv_$per.err
  PER=$per python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/curv_{os.environ['PER']}.json").read().strip().splitlines()[-1])
k = d["kernels"]["sr_curvature"]
print("tiles/CTA", os.environ["PER"], "value", round(d["value"]), "sr_curvature us", round(k["avg_us"], 1), "GB/s", round(k["gbs"]))
PY
done
