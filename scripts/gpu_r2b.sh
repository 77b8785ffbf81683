#!/bin/bash
# Round-2 GPU session B: parity, bench, ncu launch list + full capture of one steady-state scan (64 streams, one handle).
mkdir -p gpurun_out
tag=${1:-r02b}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -25 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ceiling", d["e2e"]["h2d_ceiling_scans_per_s"], "lat", d["single_stream_latency_ms"])
print({k: round(v["avg_us"], 1) for k, v in d["kernels"].items()})
print(d["laser_mapping_work"])
PY
CMD="python bench.py --workload sr_lo_lm --legs device --batch 64 --handles 1 --steps 3 --warmup 4"
# per scan: 10 SR + 6 LO + 18 LM launches = 34; skip the handle set-up launches and 4 warm-up scans, list 3 scans
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sr_|lo_|lm_" -s 140 -c 102 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1
tail -2 gpurun_out/${tag}_launches.log
ncu --set full --clock-control none --import-source on -k regex:"sr_|lo_|lm_" -s 174 -c 34 -o gpurun_out/${tag}_full $CMD > gpurun_out/${tag}_full.log 2>&1
tail -2 gpurun_out/${tag}_full.log
ls -la gpurun_out/${tag}*
