"""CPU tests of the synthetic HDL-64 generator (SURVEY.md section 8d "Synthetic input") and of bench.py's host helpers."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_generator_is_deterministic_and_shaped_like_an_hdl64_sweep(synth):
    a = synth.ScanStream(5, n_cols=128).scan(2)
    b = synth.ScanStream(5, n_cols=128).scan(2)
    assert a.shape == (64 * 128, 3) and a.dtype == np.float32
    assert np.array_equal(np.nan_to_num(a, nan=-1.0), np.nan_to_num(b, nan=-1.0))          # seeded: bit-identical
    assert not np.array_equal(np.nan_to_num(a, nan=-1.0), np.nan_to_num(synth.ScanStream(6, n_cols=128).scan(2), nan=-1.0))
    ok = np.isfinite(a).all(1)
    assert 0.3 < ok.mean() < 1.0                                                           # no-return rays are NaN (exercises removeNaN)
    r = np.linalg.norm(a[ok], axis=1)
    assert r.max() <= 80.5                                                                 # ranges are clipped at 80 m
    # ring-major emission, elevations at the bin centres of scan_registration.cpp:215-218: every finite point of ring r has
    # elevation close to ring_elevations_deg()[r] (range noise moves it by well under a bin)
    el = synth.ring_elevations_deg()
    assert el.shape == (64,) and np.all(np.diff(el) < 0)
    ang = np.degrees(np.arctan2(a[:, 2], np.hypot(a[:, 0], a[:, 1]))).reshape(64, 128)
    fin = ok.reshape(64, 128)
    for ring in range(64):
        if fin[ring].any():
            assert np.max(np.abs(ang[ring][fin[ring]] - el[ring])) < 0.12, ring


def test_generator_ring_ids_are_unambiguous_for_the_reference_formula(synth, oracle):
    """The oracle's ring classification (the reference's formula) recovers exactly the emitting ring for rings 0..50 and drops
    51..63."""
    s = synth.ScanStream(8, n_cols=256)
    sc = s.scan(0)
    ref = oracle.scan_registration(sc)
    rings = ref.laserCloud[:, 3].astype(int)
    assert rings.min() >= 0 and rings.max() <= 50
    counts = np.bincount(rings, minlength=51)
    emitted = np.isfinite(sc).all(1).reshape(64, 256)
    kept = (np.linalg.norm(np.nan_to_num(sc), axis=1) >= 5.0).reshape(64, 256) & emitted
    assert np.array_equal(counts, kept[:51].sum(1))


def test_relative_pose_composes_to_absolute_pose(synth):
    s = synth.ScanStream(3, n_cols=64)
    R0, t0 = s.pose(2)
    R1, t1 = s.pose(3)
    Rl, tl = s.relative_pose(3)
    # pose(k) = pose(k-1) o relative_pose(k)
    np.testing.assert_allclose(R0 @ Rl, R1, atol=1e-12)
    np.testing.assert_allclose(R0 @ tl + t0, t1, atol=1e-12)
    assert 0.3 < np.linalg.norm(tl) < 2.0        # 5..15 m/s at 10 Hz


def test_bench_helpers():
    b = _bench()
    assert [b.pingpong(i, 4) for i in range(8)] == [0, 1, 2, 3, 2, 1, 0, 1]
    cubes = b.synth_map_cubes(20000, 7)
    surf = np.concatenate([v for (k, c), v in cubes.items() if k == 1])
    corner = np.concatenate([v for (k, c), v in cubes.items() if k == 0])
    assert surf.shape == (20000, 4) and corner.shape[1] == 4 and len(corner) >= 1000
    # one point per 0.8 m voxel, every point inside the cube it is filed under (vloam_map_set_cube refuses anything else)
    vox = np.floor(surf[:, :3] * np.float32(1 / 0.8)).astype(np.int64)
    assert len(np.unique(vox, axis=0)) == len(surf)
    for (kind, cube), pts in cubes.items():
        i, j, k = cube % 21, (cube // 21) % 21, cube // 441
        lo = np.array([(i - 10) * 50 - 25, (j - 10) * 50 - 25, (k - 5) * 50 - 25], np.float64)
        assert np.all(pts[:, :3] >= lo) and np.all(pts[:, :3] < lo + 50), (kind, cube)
    tot = {"N": 100, "Np": 80, "nSharp": 3, "nLS": 20, "nFlat": 6, "nLF": 40, "nLSlast": 20, "nLFlast": 40, "M": 1000, "S": 50, "Mw": 100}
    for name in ("sr_curvature", "sr_classify", "lo_associate", "lm_associate", "lm_refilter", "lm_place", "lm_fit"):
        assert b.algorithmic_bytes(name, tot) > 0, name
    assert b.algorithmic_bytes("sr_curvature", tot) == 80 * 21                      # 16 B read + 4 + 1 written per kept point
    assert set(b.NCU_KERNELS["lm_place"]) == {"lm_place", "lm_compact_copy", "lm_write_back"}
    # configs[3]: the camera frames of the image front end — KITTI-sized, deterministic, distinct per stream and step parity
    f00, f01, f10 = b.frame_of(0, 0), b.frame_of(0, 1), b.frame_of(1, 0)
    assert f00.shape == (376, 1241) and f00.dtype == np.uint8 and np.array_equal(f00, b.frame_of(0, 2))
    assert np.array_equal(f01, np.roll(f00, 4, axis=1)) and not np.array_equal(f00, f10)
    fe = b.CpuFrontEnd()                       # the reference's processImage calls (OpenCV), when cv2 is importable
    if fe.cv2 is not None:
        fe.process(f00); fe.process(f01)
        assert fe.prev is not None and fe.prev.shape[1] == 32 and len(fe.prev) > 500


def test_c_ray_caster_reproduces_the_numpy_ray_caster_bit_for_bit(synth):
    """csrc_host/synth_raycast.c restates Scene.raycast_numpy operation by operation (no FMA, no BLAS): identical bits."""
    assert synth._raycast_lib() is not None, "lib/libsynth_raycast.so did not build"
    for seed in (7, 1234, 31):
        s = synth.ScanStream(seed, n_cols=256)
        for k in (0, 5, 23):
            R, t = s.pose(k)
            dw = synth._mm(s.dirs, np.ascontiguousarray(R.T))
            a, b = s.scene.raycast(t, dw), s.scene.raycast_numpy(t, dw)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (seed, k)
