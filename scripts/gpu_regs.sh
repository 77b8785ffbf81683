#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --legs device > gpurun_out/regs_$name.json 2> gpurun_out/regs_$name.err
  NAME=$name python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/regs_{os.environ['NAME']}.json").read().strip().splitlines()[-1])
k = d["kernels"]
print(os.environ["NAME"], "value", round(d["value"]), "lm_solve", round(k["lm_solve"]["avg_us"], 1), "lo_solve", round(k["lo_solve"]["avg_us"], 1))
PY
}
run auto X=1
run big VLOAM_LM_SOLVE_REGS=256 VLOAM_LO_SOLVE_REGS=256
run auto2 X=1
run big2 VLOAM_LM_SOLVE_REGS=256 VLOAM_LO_SOLVE_REGS=256
run lm_only VLOAM_LO_SOLVE_REGS=256
