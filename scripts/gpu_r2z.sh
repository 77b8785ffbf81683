#!/bin/bash
# r02z: end-of-round check of the committed state: full GPU suite, smoke(), default bench line
mkdir -p gpurun_out
tag=${1:-r02z}
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_bench.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"]), "e2e", round(e["value"]), "h2d", round(e["h2d_gbs_per_gpu"], 1), "ceiling", round(e["h2d_ceiling_gbs_per_gpu"], 1), "lat", d["single_stream_latency_ms"], "cpu", d["cpu_baseline"]["value"])
print(d["roofline"]["kernel"], d["roofline"]["frac"], d["clocks"])
print(d.get("vo_frontend"))
PY
