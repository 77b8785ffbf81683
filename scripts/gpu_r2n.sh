#!/bin/bash
# r02n: batched per-stream upload: tests + default bench A/B (VLOAM_UPLOAD_BATCH=0 = one cudaMemcpyAsync per stream)
mkdir -p gpurun_out
tag=${1:-r02n}
timeout 900 python -m pytest tests/test_gpu_lidar.py tests/test_gpu_host_api.py -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
for mode in batch percopy; do
  if [ $mode = percopy ]; then export VLOAM_UPLOAD_BATCH=0; fi
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_${mode}.json 2> gpurun_out/${tag}_${mode}.err
  tail -2 gpurun_out/${tag}_${mode}.err
  MODE=$mode TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_{os.environ['MODE']}.json").read().strip().splitlines()[-1])
e = d["e2e"]
print(os.environ["MODE"], "value", round(d["value"]), "e2e", round(e["value"]), "h2d GB/s", round(e["h2d_gbs_per_gpu"], 1), "ceiling", round(e["h2d_ceiling_gbs_per_gpu"], 1),
      "per-copy", round(e["h2d_one_memcpy_per_stream_gbs_per_gpu"], 1), "same poses", e["poses_identical_to_device_leg"], "lat", d["single_stream_latency_ms"])
PY
done
