#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02f}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
show() { NAME=$1 TAG=$tag python - <<'PY'
import json, os
try:
    d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_{os.environ['NAME']}.json").read().strip().splitlines()[-1])
    ks = d["kernels"]
    print(os.environ["NAME"], "value", round(d["value"]), "e2e", round(d.get("e2e", {}).get("value", 0)), "lat", d.get("single_stream_latency"),
          {k: round(ks[k]["avg_us"], 1) for k in ("sr_curvature", "lm_associate", "lm_place") if k in ks})
except Exception as e:
    print(os.environ["NAME"], "unreadable", e)
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_base.json 2> gpurun_out/${tag}_base.err; show base
timeout 600 python bench.py --steps 20 --warmup 5 --graphs 1 > gpurun_out/${tag}_graphs.json 2> gpurun_out/${tag}_graphs.err; tail -3 gpurun_out/${tag}_graphs.err; show graphs
VLOAM_SR_CURV_TMA=1 timeout 300 python bench.py --steps 20 --warmup 5 --legs device > gpurun_out/${tag}_curvtma.json 2> gpurun_out/${tag}_curvtma.err; show curvtma
timeout 300 python bench.py --steps 20 --warmup 5 --legs device --solver-mode 2 --handles 1 --batch 192 > gpurun_out/${tag}_wide192.json 2>&1; show wide192
