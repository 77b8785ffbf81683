#!/bin/bash
# r02k: validates the per-voxel centroid pass + the published clouds, default bench, fresh ncu launch list + full capture
mkdir -p gpurun_out
tag=${1:-r02k}
timeout 1200 python -m pytest tests/test_gpu_mapping.py tests/test_golden.py tests/test_adapter_stubs.py tests/test_gpu_lidar.py -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceiling", round(d["e2e"]["h2d_ceiling_scans_per_s"]), "lat", d["single_stream_latency_ms"])
ks = d["kernels"]
print({k: round(v["avg_us"], 1) for k, v in ks.items()})
print("sum per scan per handle us:", round(sum(v["ms_total"] for v in ks.values()) * 1e3 / (d["steps"] * d["config"]["handles"])))
PY
bash scripts/gpu_ncu.sh $tag
