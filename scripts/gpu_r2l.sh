#!/bin/bash
# r02l: validates lm_prepare / lm_insert_keys / lo_build_grid / lm_knn list changes, then A/B runs of the device leg
mkdir -p gpurun_out
tag=${1:-r02l}
timeout 1200 python -m pytest tests/test_gpu_mapping.py tests/test_golden.py tests/test_adapter_stubs.py tests/test_gpu_lidar.py -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --legs device ${BARGS} > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  NAME=$name TAG=$tag python - <<'PY'
import json, os
try:
    d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_{os.environ['NAME']}.json").read().strip().splitlines()[-1])
    ks = d["kernels"]
    sel = ("lm_associate", "lm_solve", "lo_solve", "lm_prepare", "lm_insert", "lo_build_grid", "lm_refilter", "lm_voxel", "lo_associate")
    print(os.environ["NAME"], "value", round(d["value"]), {k: round(ks[k]["avg_us"], 1) for k in sel if k in ks},
          "sum", round(sum(v["ms_total"] for v in ks.values()) * 1e3 / (d["steps"] * d["config"]["handles"])))
except Exception as e:
    print(os.environ["NAME"], "unreadable", e)
PY
}
BARGS="" run base X=1
BARGS="" run knn_direct VLOAM_LM_KNN_LIST=0
BARGS="" run solve128 VLOAM_LM_SOLVE_REGS=128 VLOAM_LO_SOLVE_REGS=128
BARGS="--handles 4" run h4 X=1
BARGS="--handles 6" run h6 X=1
BARGS="--handles 2" run h2 X=1
