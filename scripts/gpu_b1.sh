#!/bin/bash
mkdir -p gpurun_out
tag=${1:-b1}
timeout 300 python bench.py --steps 20 --warmup 5 --legs device --batch 1 --handles 1 > gpurun_out/${tag}_b1.json 2> gpurun_out/${tag}_b1.err
tail -2 gpurun_out/${tag}_b1.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_b1.json").read().strip().splitlines()[-1])
ks = d["kernels"]
tot = 0
for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["avg_us"] * kv[1]["launches"]):
    per_scan = v["avg_us"] * v["launches"] / 20
    tot += per_scan
    print(f"{k:22s} avg {v['avg_us']:7.1f} us x {v['launches']/20:.0f} = {per_scan:7.1f} us/scan")
print("sum", round(tot), "us/scan; ms_per_step", d["ms_per_step"])
PY
