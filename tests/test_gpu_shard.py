"""Point-sharded solve (SURVEY.md section 8e layout (ii)): the laser-odometry normal equations summed across ranks inside
the solve kernel.  Single-GPU form of the test: two handles of one process play rank 0 and rank 1 on two CUDA streams and
exchange through each other's buffers exactly as two processes would through IPC-mapped peer memory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_point_sharded_odometry_matches_unsharded(synth, oracle):
    import torch
    import vloam_b200 as V
    n_cols, B = 512, 2
    streams = [synth.ScanStream(91 + i, n_cols=n_cols) for i in range(B)]
    s0, s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    c0, c1, c2 = (V.Context(cuda_stream=s.cuda_stream) for s in (s0, s1, s2))
    r0 = V.LidarOdometryMapping(c0, batch=B, max_points=64 * n_cols)
    r1 = V.LidarOdometryMapping(c1, batch=B, max_points=64 * n_cols)
    ref = V.LidarOdometryMapping(c2, batch=B, max_points=64 * n_cols)
    ptrs = [r0.shard_buffer(), r1.shard_buffer()]
    r0.shard_enable(0, 2, ptrs)
    r1.shard_enable(1, 2, ptrs)
    olo = [oracle.LaserOdometry() for _ in range(B)]
    for k in range(4):
        scans = np.stack([s.scan(k) for s in streams])
        for h in (r0, r1, ref):
            h.reset()
            h.scanRegistrationIO(scans)
        # both "ranks" must be in flight together: enqueue without waiting, then read
        r0.laserOdometryIO(fetch=False)
        r1.laserOdometryIO(fetch=False)
        p0, p1 = r0.lo_pose(), r1.lo_pose()
        pr = ref.laserOdometryIO()
        assert r0.shard_status() == 0 and r1.shard_status() == 0
        for key in ("q_last_curr", "t_last_curr", "q_w_curr", "t_w_curr"):
            assert np.array_equal(p0[key], p1[key]), f"scan {k}: ranks disagree on {key}"          # bit-identical by construction
            np.testing.assert_allclose(p0[key], pr[key], atol=1e-10, err_msg=f"scan {k}: {key}")   # summation order only
        assert np.array_equal(p0["corner_correspondence"], pr["corner_correspondence"])
        assert np.array_equal(p0["plane_correspondence"], pr["plane_correspondence"])
        for b in range(B):
            olo[b].solve(oracle.scan_registration(scans[b]))
            np.testing.assert_allclose(p0["t_last_curr"][b], olo[b].state["t_last_curr"], atol=1e-4)
    # the slices are disjoint and together they are the unsharded correspondence set of the last pass
    for b in range(B):
        t0, t1, tr = r0.lo_trace(1, b), r1.lo_trace(1, b), ref.lo_trace(1, b)
        for kind in ("corner", "plane"):
            q0, q1 = set(t0[kind][:, 0].tolist()), set(t1[kind][:, 0].tolist())
            assert q0 and q1 and not (q0 & q1)
        # (the reference handle associates with its own, marginally different pose; compare sizes only)
        assert abs(len(t0["plane"]) + len(t1["plane"]) - len(tr["plane"])) <= 2
        assert t0["n_plane"] == t1["n_plane"] == tr["n_plane"]          # the counts are exchanged too
    for h in (r0, r1, ref):
        h.close()
    for c in (c0, c1, c2):
        c.close()


def test_unanswered_peer_is_reported_not_hung(synth):
    """A rank whose peer never shows up must come back with an error bit, not spin forever."""
    import torch
    import vloam_b200 as V
    n_cols = 256
    st = synth.ScanStream(5, n_cols=n_cols)
    a = V.LidarOdometryMapping(batch=1, max_points=64 * n_cols)
    b = V.LidarOdometryMapping(a.ctx, batch=1, max_points=64 * n_cols)   # never runs: its slots stay at sequence 0
    a.shard_enable(0, 2, [a.shard_buffer(), b.shard_buffer()])
    for k in range(2):
        a.reset()
        a.scanRegistrationIO(st.scan(k))
        a.laserOdometryIO(fetch=False)
    a.ctx.synchronize()
    assert a.shard_status() != 0
    b.close(); a.close()        # a owns the context b borrows


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_point_sharded_two_processes(tmp_path, mode):
    """The same split across two PROCESSES on two GPUs (torchrun): `peer` maps the other rank's exchange slots through CUDA IPC
    and sums inside the running solve kernel over NVLink; `nccl` all-reduces the tile partials of the wide solve (odometry and
    mapping).  Both ranks must end on bit-identical poses, equal to an unsharded handle's up to summation order."""
    import os
    import subprocess
    import sys
    if not _two_gpus():
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531" if mode == "peer" else "29532", os.path.join(root, "tests", "workers", "shard_two_process.py"),
           "--mode", mode, "--out", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert int(r0["shard_status"]) == 0 and int(r1["shard_status"]) == 0
    for k in range(4):
        for key in ("q_last_curr", "t_last_curr", "q_w_curr", "t_w_curr"):
            a, b, ref = r0[f"lo_{key}_{k}"], r1[f"lo_{key}_{k}"], r0[f"ref_lo_{key}_{k}"]
            assert np.array_equal(a, b), f"scan {k}: ranks disagree on {key}"
            np.testing.assert_allclose(a, ref, atol=1e-9, err_msg=f"scan {k}: {key}")
        for key in ("corner_correspondence", "plane_correspondence"):
            assert np.array_equal(r0[f"lo_{key}_{k}"], r0[f"ref_lo_{key}_{k}"])
        if mode == "nccl":
            assert np.array_equal(r0[f"lm_q_{k}"], r1[f"lm_q_{k}"]) and np.array_equal(r0[f"lm_t_{k}"], r1[f"lm_t_{k}"])
            np.testing.assert_allclose(r0[f"lm_t_{k}"], r0[f"ref_lm_t_{k}"], atol=1e-6)
            np.testing.assert_allclose(np.abs(r0[f"lm_q_{k}"]), np.abs(r0[f"ref_lm_q_{k}"]), atol=1e-6)
