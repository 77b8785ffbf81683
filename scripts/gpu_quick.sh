#!/bin/bash
# quick GPU check: mapping + golden parity tests, then the default bench; prints per-kernel times
mkdir -p gpurun_out
tag=${1:-quick}
sel=${2:-"tests/test_gpu_mapping.py tests/test_golden.py"}
timeout 900 python -m pytest $sel -m gpu -q --maxfail=5 > gpurun_out/${tag}_pytest.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceiling", round(d["e2e"]["h2d_ceiling_scans_per_s"]), "lat", d["single_stream_latency_ms"])
ks = d["kernels"]
print({k: round(v["avg_us"], 1) for k, v in ks.items()})
print("sum per scan per handle us:", round(sum(v["ms_total"] for v in ks.values()) * 1e3 / (d["steps"] * d["config"]["handles"])))
print(d["laser_mapping_work"])
PY
