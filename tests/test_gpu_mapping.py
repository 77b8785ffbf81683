"""GPU parity tests for laserMapping: CUDA path through the C-ABI vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-4


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _quat_angle(q1, q2):
    return 2.0 * np.arccos(min(1.0, abs(float(np.dot(q1, q2)))))


def _same_xyz(a, b, name):
    assert a.shape == b.shape, f"{name}: {a.shape} vs {b.shape}"
    if a.shape[0]:
        assert np.array_equal(_bits(a[:, :3]), _bits(b[:, :3])), f"{name}: xyz not bit-identical"
        assert np.max(np.abs(a[:, 3] - b[:, 3])) < 0.14, f"{name}: intensity"


def _run_sequence(V, oracle, scans, check_cubes=True, map_capacity=1 << 18, stats_out=None, before_scan=None):
    lom = V.LidarOdometryMapping(batch=1, max_points=scans[0].shape[0], map_capacity_points=map_capacity, debug_keep_submap=1)
    pipe = oracle.Pipeline()
    worst_t = 0.0
    for k, scan in enumerate(scans):
        if before_scan is not None:
            before_scan(k, lom, pipe)
        lom.reset()
        lom.scanRegistrationIO(scan)
        lom.laserOdometryIO()
        mp = lom.laserMappingIO()
        assert pipe.process(scan, do_mapping=True) == 0
        ost = pipe.lm.state
        info = lom.lm_info()[0]
        assert list(info[:4]) == list(ost["cen"]) + [ost["validNum"]], (k, info, ost)
        _same_xyz(lom.cloud(V.CLOUD_CORNER_STACK), pipe.lm.cloud(0), f"scan {k} corner stack")
        _same_xyz(lom.cloud(V.CLOUD_SURF_STACK), pipe.lm.cloud(1), f"scan {k} surf stack")
        _same_xyz(lom.cloud(V.CLOUD_CORNER_MAP), pipe.lm.cloud(2), f"scan {k} corner from map")
        _same_xyz(lom.cloud(V.CLOUD_SURF_MAP), pipe.lm.cloud(3), f"scan {k} surf from map")
        otr = pipe.lm.trace()
        if k == 0:
            assert len(otr) == 0                      # empty map: the gate at laser_mapping.cpp:448 fails
        for p, t in enumerate(otr):
            g = lom.lm_trace(p)
            assert (g["n_corner"], g["n_plane"]) == (len(t["corner"]), len(t["plane"])), (k, p)
            assert g["termination"] == t["termination"]
            n = t["iterations"].shape[0]
            assert g["n_records"] == n
            np.testing.assert_allclose(g["iterations"][:n, 0], t["iterations"][:, 0], rtol=1e-8, atol=1e-12)
            np.testing.assert_array_equal(g["iterations"][:n, 5:7], t["iterations"][:, 5:7])
            np.testing.assert_allclose(g["para"], t["para"], atol=1e-8)
        dt = float(np.max(np.abs(mp["t_w_curr"][0] - ost["t_w_curr"])))
        worst_t = max(worst_t, dt)
        assert dt < POSE_TOL_M
        assert _quat_angle(mp["q_w_curr"][0], ost["q_w_curr"]) < POSE_TOL_RAD
        assert np.max(np.abs(mp["t_wmap_wodom"][0] - ost["t_wmap_wodom"])) < POSE_TOL_M
        assert _quat_angle(mp["q_wmap_wodom"][0], ost["q_wmap_wodom"]) < POSE_TOL_RAD
        if check_cubes:
            occupied = 0
            for cube in range(4851):
                for kind in (0, 1):
                    n_or = pipe.lm.cube_count(kind, cube)
                    if n_or or k == 0 and cube % 97 == 0:
                        _same_xyz(lom.map_get_cube(kind, cube), pipe.lm.cube(kind, cube), f"scan {k} cube {cube} kind {kind}")
                        occupied += n_or > 0
            assert occupied > 0
        ms = lom.map_stats()[0]
        for kind in (0, 1):      # storage bookkeeping: the tables account for exactly the oracle's map
            assert ms[kind, 0] == sum(pipe.lm.cube_count(kind, c) for c in range(4851))
            assert ms[kind, 0] <= ms[kind, 7] <= ms[kind, 1] <= map_capacity
        if stats_out is not None:
            stats_out.append(ms.copy())
    lom.close()
    return worst_t


def test_laser_mapping_sequence(synth, oracle):
    import vloam_b200 as V
    s = synth.ScanStream(31, n_cols=1024)
    scans = [s.scan(k) for k in range(5)]
    worst = _run_sequence(V, oracle, scans)
    print("max |t_w_curr - oracle| =", worst)


def test_laser_mapping_full_size(synth, oracle):
    """BASELINE configs[2] shape: full 64 x 2048 scans through scan registration, odometry and mapping (the map is built
    by the scans themselves); every cube, trace and pose against the oracle after every scan."""
    import vloam_b200 as V
    s = synth.ScanStream(33, n_cols=2048)
    scans = [s.scan(k) for k in range(3)]
    worst = _run_sequence(V, oracle, scans, map_capacity=1 << 17)
    print("max |t_w_curr - oracle| =", worst)


def test_laser_mapping_incremental_refilter_and_repack(synth, oracle):
    """The per-scan re-filter of every valid cube (laser_mapping.cpp:689-702) is skipped for cubes that are fixed points
    of their voxel filter and received nothing; rewritten cubes keep their slab, move to a fresh one or trigger a re-pack
    of the whole map into the other pool.  A tight map capacity forces all three placements; the map must stay identical
    to the oracle's (which re-filters everything, like the reference) cube by cube after every scan."""
    import vloam_b200 as V
    s = synth.ScanStream(31, n_cols=1024)
    scans = [s.scan(k) for k in range(8)]
    stats = []
    _run_sequence(V, oracle, scans, map_capacity=9000, stats_out=stats)
    st = np.stack(stats)                                  # (scan, kind, 8)
    assert st[-1, :, 6].max() >= 1, st[:, :, 6]           # at least one re-pack happened
    assert (st[1:, :, 4] > 0).any()                       # some cubes were found in fixed-point form ...
    assert (st[2:, :, 5] < st[2:, :, 3]).any(), st[:, :, [3, 5]]   # ... so later scans rewrote fewer cubes than the map holds


def test_laser_mapping_batched_streams_on_seeded_map(synth, oracle):
    """Three streams in one handle (the bench's layout: blockIdx.z = stream), each on its own pre-built map, against three
    oracle pipelines: poses, solver traces and every occupied cube after every scan.  The seeded map is mostly untouched
    by the scans, so most of its cubes must be skipped by the re-filter once they are in fixed-point form."""
    import vloam_b200 as V
    B = 3
    streams = [synth.ScanStream(40 + b, n_cols=512) for b in range(B)]
    cap = 64 * 512
    lom = V.LidarOdometryMapping(batch=B, max_points=cap, map_capacity_points=1 << 17)
    pipes = [oracle.Pipeline() for _ in range(B)]
    rng = np.random.default_rng(5)
    for b in range(B):          # a coarse ground lattice + poles in the 3 x 3 cubes around the origin, different per stream
        g = np.arange(-60.0, 60.0, 0.45 + 0.05 * b)
        gx, gy = np.meshgrid(g, g)
        surf = np.c_[gx.ravel(), gy.ravel(), np.full(gx.size, -1.73)] + rng.uniform(-0.2, 0.2, (gx.size, 3)) * [1, 1, 0.05]
        cx, cy = rng.uniform(-60, 60, 40), rng.uniform(-60, 60, 40)
        corner = np.c_[np.repeat(cx, 12), np.repeat(cy, 12), np.tile(np.arange(12) * 0.45 - 1.5, 40)]
        for kind, cloud in ((0, corner), (1, surf)):
            cloud = np.c_[cloud, np.zeros(len(cloud))].astype(np.float32)
            ci = (np.floor((cloud[:, 0] + 25.0) / 50.0).astype(int) + 10) + 21 * (np.floor((cloud[:, 1] + 25.0) / 50.0).astype(int) + 10) \
                + 441 * (np.floor((cloud[:, 2] + 25.0) / 50.0).astype(int) + 5)
            for c in np.unique(ci):
                lom.map_set_cube(kind, int(c), cloud[ci == c], stream=b)
                pipes[b].lm.set_cube(kind, int(c), cloud[ci == c])
    with pytest.raises(V.VloamError):      # content outside the named cube is refused
        lom.map_set_cube(1, 10 + 21 * 10 + 441 * 5, np.array([[100.0, 0, 0, 0]], np.float32))
    for k in range(4):
        buf = np.stack([s.scan(k) for s in streams])
        lom.reset()
        lom.scanRegistrationIO(buf)
        lom.laserOdometryIO()
        mp = lom.laserMappingIO()
        for b in range(B):
            assert pipes[b].process(buf[b], do_mapping=True) == 0
            ost = pipes[b].lm.state
            assert np.max(np.abs(mp["t_w_curr"][b] - ost["t_w_curr"])) < POSE_TOL_M, (k, b)
            assert _quat_angle(mp["q_w_curr"][b], ost["q_w_curr"]) < POSE_TOL_RAD
            for p, t in enumerate(pipes[b].lm.trace()):
                g = lom.lm_trace(p, b)
                assert (g["n_corner"], g["n_plane"]) == (len(t["corner"]), len(t["plane"])), (k, b, p)
                assert g["n_corner"] + g["n_plane"] > 100
            for kind in (0, 1):
                for cube in range(4851):
                    if pipes[b].lm.cube_count(kind, cube):
                        _same_xyz(lom.map_get_cube(kind, cube, stream=b), pipes[b].lm.cube(kind, cube), f"scan {k} stream {b} cube {cube} kind {kind}")
        ms = lom.map_stats()
        if k >= 2:      # steady state: fewer cubes rewritten than the map holds, none re-packed
            assert (ms[:, 1, 5] < ms[:, 1, 3]).all(), ms[:, 1, :]
            assert (ms[:, :, 4] > 0).all()
    lom.close()


def test_laser_mapping_column_table_recycling(synth, oracle, monkeypatch):
    """The per-cube column tables come from a fixed pool of slots; cubes dropped by grid shifts leak theirs until the pool
    runs out, then every index is dropped and rebuilt.  With only 14 slots per kind that happens several times in 8
    scans; map and poses must stay identical to the oracle's."""
    import vloam_b200 as V
    monkeypatch.setenv("VLOAM_LM_TAB_SLOTS", "16")
    s = synth.ScanStream(31, n_cols=1024)
    scans = [s.scan(k) for k in range(8)]
    rng = np.random.default_rng(9)
    cube = 10 + 21 * 10 + 441 * 6            # the cube above the sensor (z in [25, 75)): valid, never reached by the scans

    def reseed(k, lom, pipe):                # replacing a cube's content drops its column index: its table slot leaks
        if k == 0:
            return                           # (keep the first scan's empty-map case, laser_mapping.cpp:448)
        pts = np.c_[rng.uniform(-20, 20, (80, 2)), rng.uniform(30, 70, 80), np.zeros(80)].astype(np.float32)
        for kind in (0, 1):
            lom.map_set_cube(kind, cube, pts)
            pipe.lm.set_cube(kind, cube, pts)

    stats = []
    _run_sequence(V, oracle, scans, stats_out=stats, before_scan=reseed)
    st = np.stack(stats)                     # (scan, kind, 10): [9] = table slots handed out
    assert st[:, :, 9].max() <= 16, st[:, :, 9]
    assert (np.diff(st[:, 1, 9]) < 0).any(), st[:, 1, 9]      # the slot counter started over at least once


def test_laser_mapping_seeded_map_and_cube_shift(synth, oracle):
    """Map cubes seeded through vloam_map_set_cube; a large odometry offset forces the rolling grid to shift."""
    import vloam_b200 as V
    s = synth.ScanStream(32, n_cols=512)
    scans = [s.scan(k) for k in range(3)]
    lom = V.LidarOdometryMapping(batch=1, max_points=scans[0].shape[0], map_capacity_points=1 << 18, debug_keep_submap=1)
    pipe = oracle.Pipeline()
    # seed one far-away cube in both maps: it must survive the grid shift at its shifted index or be dropped identically
    rng = np.random.default_rng(0)
    seed = np.c_[rng.uniform(280, 320, (500, 2)), rng.uniform(-2, 2, 500), rng.uniform(0, 50, 500)].astype(np.float32)
    cube = (10 + 6) + 21 * (10 + 6) + 441 * 5
    lom.map_set_cube(1, cube, seed)
    pipe.lm.set_cube(1, cube, seed)
    assert np.array_equal(lom.map_get_cube(1, cube), seed)
    for k, scan in enumerate(scans):
        # shift both odometries by 400 m in x so that centerCubeI >= laserCloudWidth - 3 (laser_mapping.cpp:249)
        sc = scan
        lom.reset(); lom.scanRegistrationIO(sc); lom.laserOdometryIO()
        assert pipe.process(sc, do_mapping=False) == 0
        if k == 0:
            off = np.array([[0, 0, 0, 1, 400.0, 0, 0]])
        lom_pose = lom.lo_pose()
        mp = lom.laserMappingIO()
        pipe.lm.reset(); pipe.lm.input_from_lo(pipe.lo); pipe.lm.solve()
        ost = pipe.lm.state
        assert list(lom.lm_info()[0][:3]) == list(ost["cen"])
        assert np.max(np.abs(mp["t_w_curr"][0] - ost["t_w_curr"])) < POSE_TOL_M
    lom.close()
