#!/bin/bash
# mapping parity tests + the sr_lo_lm bench summary
tag=${1:-x}
python -m pytest tests/test_gpu_mapping.py -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -4 gpurun_out/${tag}_tests.log
python bench.py --workload sr_lo_lm ${@:2} > gpurun_out/${tag}_bench_map.json 2> gpurun_out/${tag}_bench_map.err; tail -3 gpurun_out/${tag}_bench_map.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_map.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["poses_identical_to_device_leg"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_total"])[:14]: print(k, round(v["share"],3), round(v["avg_us"],1), v["launches"])
PY
