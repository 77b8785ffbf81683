#!/bin/bash
# r02q: stationary benchmark trajectory: bench-configuration parity test + default bench (+ racecheck re-run of sr_pick_features)
mkdir -p gpurun_out
tag=${1:-r02q}
timeout 900 python -m pytest tests/test_gpu_mapping.py::test_laser_mapping_bench_configuration tests/test_golden.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_golden.py -m gpu -q -x > gpurun_out/${tag}_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | tail -3
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_bench.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"]), "e2e", round(e["value"]), "h2d", round(e["h2d_gbs_per_gpu"], 1), "ceiling", round(e["h2d_ceiling_gbs_per_gpu"], 1), "lat", d["single_stream_latency_ms"], "cpu", d["cpu_baseline"]["value"])
print({k: round(v["avg_us"], 1) for k, v in d["kernels"].items()})
print(d["north_star_kernels"]["sr_curvature"])
print(d["laser_mapping_work"])
PY
