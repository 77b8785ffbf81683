#!/bin/bash
# Round-2 GPU session A: parity first (fail fast), then the bench, then the ncu artefacts.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
tag=${1:-r02a}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1 || { tail -30 gpurun_out/${tag}_smoke.log; echo SMOKE FAILED; }
tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python -m pytest tests/test_gpu_mapping.py -q -x -k "sequence or full_size" > gpurun_out/${tag}_first.log 2>&1
tail -15 gpurun_out/${tag}_first.log
if ! grep -q "passed" gpurun_out/${tag}_first.log || grep -q "failed" gpurun_out/${tag}_first.log; then echo "BASIC MAPPING PARITY FAILED - stopping"; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=15 > gpurun_out/${tag}_pytest.log 2>&1
tail -40 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
VLOAM_LM_KNN_OCC=3 timeout 300 python bench.py --steps 20 --warmup 5 --legs device > gpurun_out/${tag}_bench_knnocc3.json 2>&1
TAG=$tag python - <<'PY'
import json, os
for f in ("bench", "bench_knnocc3"):
    try:
        d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
# memcheck of the new index / k-NN / write-back paths on a small sequence (slow under the sanitizer: bounded)
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mapping.py -q -x -k "incremental_refilter" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | tail -3
