#!/bin/bash
# r02p: full GPU suite, detector timing, compute-sanitizer (memcheck + racecheck) on small parity tests, default bench, ncu
mkdir -p gpurun_out
tag=${1:-r02p}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log
timeout 200 python scripts/gpu_detect_timing.py 32 > gpurun_out/${tag}_detect.json 2> gpurun_out/${tag}_detect.err; cat gpurun_out/${tag}_detect.json
SEL="tests/test_golden.py tests/test_vo_frontend.py tests/test_gpu_mapping.py::test_laser_mapping_sequence tests/test_gpu_lidar.py::test_motion_distortion_path"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest $SEL -m gpu -q -x > gpurun_out/${tag}_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_golden.py tests/test_vo_frontend.py -m gpu -q -x > gpurun_out/${tag}_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | tail -4
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_bench.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"]), "e2e", round(e["value"]), "h2d", round(e["h2d_gbs_per_gpu"], 1), "ceiling", round(e["h2d_ceiling_gbs_per_gpu"], 1), "lat", d["single_stream_latency_ms"], "cpu", d["cpu_baseline"]["value"])
print({k: round(v["avg_us"], 1) for k, v in d["kernels"].items()})
print(d["roofline"])
PY
bash scripts/gpu_ncu.sh $tag
