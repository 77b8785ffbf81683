#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02i}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --legs device ${BARGS} > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  NAME=$name TAG=$tag python - <<'PY'
import json, os
try:
    d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_{os.environ['NAME']}.json").read().strip().splitlines()[-1])
    ks = d["kernels"]
    sel = ("lm_associate", "lm_solve", "lo_solve", "lm_accumulate", "lm_step", "lm_place", "lm_fit")
    print(os.environ["NAME"], "value", round(d["value"]), {k: round(ks[k]["avg_us"], 1) for k in sel if k in ks})
except Exception as e:
    print(os.environ["NAME"], "unreadable", e)
PY
}
BARGS="" run base X=1
BARGS="" run solve128 VLOAM_LM_SOLVE_REGS=128
BARGS="--solver-mode 2" run wide X=1
