#!/bin/bash
N=${1:-2}
tag=${2:-r02_pt$N}
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" \
     > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  echo "== $name rc=$?"; grep -i "error" gpurun_out/${tag}_${name}.err | tail -3 | cut -c1-300
  NAME=$name TAG=$tag python - <<'PY'
import json, os
try:
    d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_{os.environ['NAME']}.json").read().strip().splitlines()[-1])
    ks = d["kernels"]
    print(os.environ["NAME"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/rank", [round(x, 2) for x in d["ms_per_step_per_rank"]],
          {k: round(ks[k]["avg_us"], 1) for k in ("lo_accumulate", "lo_step", "lm_accumulate", "lm_step", "lo_solve", "lm_solve", "lo_associate", "lm_associate") if k in ks})
except Exception as e:
    print(os.environ["NAME"], "unreadable", e)
PY
}
run point_b8 --parallelism point --batch 8
run point_b64 --parallelism point --batch 64
