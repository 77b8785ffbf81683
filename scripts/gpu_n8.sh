#!/bin/bash
# 8 GPUs, final code: stream-sharded (weak scaling) and the NCCL point-sharded layout at 64 streams
N=8; tag=${1:-r02u_n8}
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" \
     > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  echo "== $name rc=$?"
  NAME=$name TAG=$tag python - <<'PY'
import json, os
try:
    d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_{os.environ['NAME']}.json").read().strip().splitlines()[-1])
    print(os.environ["NAME"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceiling", round(d["e2e"]["h2d_ceiling_scans_per_s"]),
          "ms/rank", [round(x, 2) for x in d["ms_per_step_per_rank"]], "e2e ms/rank", [round(x, 2) for x in d["e2e"]["ms_per_step_per_rank"]],
          "h2d", round(d["e2e"]["h2d_gbs_per_gpu"], 1), round(d["e2e"]["h2d_ceiling_gbs_per_gpu"], 1), round(d["e2e"]["h2d_one_memcpy_per_stream_gbs_per_gpu"], 1))
except Exception as e:
    print(os.environ["NAME"], "unreadable", e)
PY
}
run stream
run point_b64 --parallelism point --batch 64
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1; lscpu | grep -E "NUMA|^CPU\(s\)|Model name" >> gpurun_out/${tag}_topo.txt
