"""Worker of tests/test_gpu_shard.py::test_point_sharded_two_processes — one process per GPU under torchrun.
Every rank feeds the same scans to a point-sharded handle (--mode peer: normal equations summed inside the solve kernel through
CUDA-IPC peer memory; --mode nccl: ncclAllReduce of the tile partials, odometry and mapping); rank 0 also runs an unsharded
handle on the same scans.  Writes <out>/rank<r>.npz with the poses of every scan (and the unsharded ones on rank 0)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["peer", "nccl"], required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--scans", type=int, default=4)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import vloam_b200 as V
    from vloam_b200 import dist as D, synth
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=device)
    n_cols, B = 512, 2
    streams = [synth.ScanStream(91 + i, n_cols=n_cols) for i in range(B)]
    ctx = V.Context(device=rank)
    lom = V.LidarOdometryMapping(ctx, batch=B, max_points=64 * n_cols)
    ref = V.LidarOdometryMapping(ctx, batch=B, max_points=64 * n_cols) if rank == 0 else None
    if a.mode == "peer":
        D.enable_point_sharding(lom, dist, device)
    else:
        D.enable_point_sharding_nccl(lom, dist, device)
    out = {}
    for k in range(a.scans):
        scans = np.stack([s.scan(k) for s in streams])
        lom.reset(); lom.scanRegistrationIO(scans)
        p = lom.laserOdometryIO()
        m = lom.laserMappingIO() if a.mode == "nccl" else None
        for key in ("q_last_curr", "t_last_curr", "q_w_curr", "t_w_curr", "corner_correspondence", "plane_correspondence"):
            out[f"lo_{key}_{k}"] = np.asarray(p[key])
        if m is not None:
            out[f"lm_q_{k}"], out[f"lm_t_{k}"] = np.asarray(m["q_w_curr"]), np.asarray(m["t_w_curr"])
        if ref is not None:
            ref.reset(); ref.scanRegistrationIO(scans)
            pr = ref.laserOdometryIO()
            for key in ("q_last_curr", "t_last_curr", "q_w_curr", "t_w_curr", "corner_correspondence", "plane_correspondence"):
                out[f"ref_lo_{key}_{k}"] = np.asarray(pr[key])
            if a.mode == "nccl":
                mr = ref.laserMappingIO()
                out[f"ref_lm_q_{k}"], out[f"ref_lm_t_{k}"] = np.asarray(mr["q_w_curr"]), np.asarray(mr["t_w_curr"])
    out["shard_status"] = np.array(lom.shard_status() if a.mode == "peer" else 0)
    np.savez(os.path.join(a.out, f"rank{rank}.npz"), **out)
    dist.barrier()
    if a.mode == "nccl":
        lom.shard_nccl_destroy()
    lom.close()
    if ref is not None:
        ref.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
