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
        assert np.array_equal(_bits(a[:, 3]), _bits(b[:, 3])), f"{name}: intensity not bit-identical"


def _run_sequence(V, oracle, scans, check_cubes=True, map_capacity=1 << 18, stats_out=None, before_scan=None, solver_mode=0, nccl=False):
    lom = V.LidarOdometryMapping(batch=1, max_points=scans[0].shape[0], map_capacity_points=map_capacity, debug_keep_submap=1,
                                 solver_mode=solver_mode)
    if nccl:
        lom.shard_nccl_init(0, 1, V.shard_nccl_unique_id())
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
            # the same queries produced the factors (laser_mapping.cpp:472-581), not just the same number of them
            assert np.array_equal(lom.lm_queries(p, 0), t["corner"].ravel()), (k, p, "corner queries")
            assert np.array_equal(lom.lm_queries(p, 1), t["plane"].ravel()), (k, p, "surf queries")
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
        if check_cubes:
            # what LaserMapping::publish sends: /laser_cloud_map (every cube, corner then surf, laser_mapping.cpp:778-790) bit
            # for bit; /velodyne_cloud_registered (:797-805) through the device's own mapping pose, which agrees with the
            # oracle's to ~1e-8, so the float-rounded coordinates may differ in the last place
            _same_xyz(lom.cloud(V.CLOUD_MAP), pipe.lm.map_cloud(), f"scan {k} /laser_cloud_map")
            reg, oreg = lom.cloud(V.CLOUD_FULL_REGISTERED), pipe.lm.publish_registered()
            assert reg.shape == oreg.shape and reg.shape[0] > 1000
            assert np.array_equal(_bits(reg[:, 3]), _bits(oreg[:, 3]))
            np.testing.assert_allclose(reg[:, :3], oreg[:, :3], atol=3e-5, rtol=0)
            assert np.mean(_bits(reg[:, :3]) == _bits(oreg[:, :3])) > 0.99
        ms = lom.map_stats()[0]
        for kind in (0, 1):      # storage bookkeeping: the tables account for exactly the oracle's map
            assert ms[kind, 0] == sum(pipe.lm.cube_count(kind, c) for c in range(4851))
            assert ms[kind, 0] <= ms[kind, 7] <= ms[kind, 1] <= map_capacity
        assert not lom.lm_status()[0].any()
        if stats_out is not None:
            stats_out.append(ms.copy())
    lom.close()
    return worst_t


@pytest.mark.parametrize("solver_mode", [1, 2])
def test_laser_mapping_sequence(synth, oracle, solver_mode):
    """solver_mode 1: the whole ceres::Solve of a pass in one launch, a CTA (cluster) per stream; 2: one wide accumulate launch
    per Levenberg-Marquardt evaluation + a warp-per-stream step (gn_split.cuh).  Same traces, same poses, same map."""
    import vloam_b200 as V
    s = synth.ScanStream(31, n_cols=1024)
    scans = [s.scan(k) for k in range(5)]
    worst = _run_sequence(V, oracle, scans, solver_mode=solver_mode)
    print("max |t_w_curr - oracle| =", worst)


def test_laser_mapping_sequence_nccl_exchange_single_rank(synth, oracle):
    """The point-sharded layout of BASELINE configs[4] with a group of one rank: wide accumulate, ncclAllReduce of the partial
    normal equations (the identity here), step — odometry and mapping.  Exercises the run-time NCCL binding and the
    collective path on one GPU; the multi-GPU run is scripts/run_point_sharded.sh."""
    import vloam_b200 as V
    s = synth.ScanStream(31, n_cols=1024)
    scans = [s.scan(k) for k in range(4)]
    _run_sequence(V, oracle, scans, nccl=True)


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


def test_laser_mapping_32_streams_per_launch(synth, oracle):
    """From 32 streams per launch on the mapping solve runs in its 128-register variant (lm_solve_r128, two CTAs per SM) — the
    one the benchmark's 64-stream handles use: 32 different streams against 32 oracle pipelines, poses, traces and query sets
    after every scan."""
    import vloam_b200 as V
    B = 32
    streams = [synth.ScanStream(500 + b, n_cols=256) for b in range(B)]
    lom = V.LidarOdometryMapping(batch=B, max_points=64 * 256, map_capacity_points=1 << 16)
    pipes = [oracle.Pipeline() for _ in range(B)]
    solved = 0
    for k in range(4):
        buf = np.stack([s.scan(k) for s in streams])
        lom.reset(); lom.scanRegistrationIO(buf); lom.laserOdometryIO()
        mp = lom.laserMappingIO()
        assert not lom.lm_status().any()
        for b in range(B):
            assert pipes[b].process(buf[b], do_mapping=True) == 0
            ost = pipes[b].lm.state
            assert np.max(np.abs(mp["t_w_curr"][b] - ost["t_w_curr"])) < POSE_TOL_M, (k, b)
            assert _quat_angle(mp["q_w_curr"][b], ost["q_w_curr"]) < POSE_TOL_RAD
            for p, t in enumerate(pipes[b].lm.trace()):
                g = lom.lm_trace(p, b)
                assert np.array_equal(lom.lm_queries(p, 0, b), t["corner"].ravel()) and np.array_equal(lom.lm_queries(p, 1, b), t["plane"].ravel())
                n = t["iterations"].shape[0]
                assert g["n_records"] == n and g["termination"] == t["termination"]
                np.testing.assert_allclose(g["iterations"][:n, 0], t["iterations"][:, 0], rtol=1e-8, atol=1e-12)
                np.testing.assert_allclose(g["para"], t["para"], atol=1e-8)
                solved += 1
    assert solved >= 3 * 2 * B
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


def _cube_of(p, cen=(10, 10, 5)):
    c = np.floor((np.asarray(p, np.float64) + 25.0) / 50.0).astype(int) + np.asarray(cen)
    return c[..., 0] + 21 * c[..., 1] + 441 * c[..., 2]


def _compare_all_cubes(lom, olm, tag, stream=0):
    """Every cube of both kinds: identical content (so also: dropped by the same shifts)."""
    ms = lom.map_stats()[stream]
    occupied = 0
    for kind in (0, 1):
        total = 0
        for cube in range(4851):
            n_or = olm.cube_count(kind, cube)
            total += n_or
            if n_or:
                _same_xyz(lom.map_get_cube(kind, cube, stream=stream), olm.cube(kind, cube), f"{tag} cube {cube} kind {kind}")
                occupied += 1
        assert ms[kind, 0] == total, (tag, kind, ms[kind, 0], total)      # ... and nothing else anywhere
    return occupied


def test_laser_mapping_cube_grid_shift_all_directions(synth, oracle):
    """laser_mapping.cpp:218-402: jumps of the odometry pose (vloam_set_lo_pose on both sides, before laserMapping) drive the
    centre cube out of [3, W-3) in +x, +y, -x / +z, -y / -z (single and double steps) and back; the rolling grid must shift
    by the same steps, seeded far-away cubes must survive at their shifted index or be dropped exactly like the oracle's,
    and around the jump targets a seeded ground lattice + poles makes the solve run on the shifted tables."""
    import vloam_b200 as V
    s = synth.ScanStream(32, n_cols=512)
    scans = [s.scan(k) for k in range(6)]
    lom = V.LidarOdometryMapping(batch=1, max_points=scans[0].shape[0], map_capacity_points=1 << 18, debug_keep_submap=1)
    pipe = oracle.Pipeline()
    rng = np.random.default_rng(0)
    seeded = {}

    def seed(kind, pts):
        pts = np.c_[pts, np.zeros(len(pts))].astype(np.float32)
        ci = _cube_of(pts[:, :3])
        for c in np.unique(ci):
            seeded[(kind, int(c))] = pts[ci == c]
    # far cubes that only exist to be carried around / dropped by the shifts (at the grid's border planes too)
    seed(1, np.c_[rng.uniform(280, 320, (500, 2)), rng.uniform(-2, 2, 500)])
    seed(1, np.c_[rng.uniform(-520, -480, 300), rng.uniform(-20, 20, 300), rng.uniform(-2, 2, 300)])      # i = 0: dropped by a shift towards -x
    seed(0, np.c_[rng.uniform(480, 520, 300), rng.uniform(-20, 20, 300), rng.uniform(-2, 2, 300)])        # i = 20
    seed(1, np.c_[rng.uniform(-20, 20, (300, 2)), rng.uniform(230, 270, 300)])                             # k = 10
    seed(0, np.c_[rng.uniform(-20, 20, 300), rng.uniform(-520, -480, 300), rng.uniform(-2, 2, 300)])      # j = 0
    # a ground lattice and poles around the jump targets, so that the gate of :448 passes and the solve runs there
    for cx, cy in ((400.0, 0.0), (400.0, 400.0), (-400.0, 400.0)):
        g = np.arange(-70.0, 70.0, 0.8)
        gx, gy = np.meshgrid(g + cx, g + cy)
        seed(1, np.c_[gx.ravel(), gy.ravel(), np.full(gx.size, -1.73)] + rng.uniform(-0.2, 0.2, (gx.size, 3)) * [1, 1, 0.02])
        px, py = rng.uniform(cx - 60, cx + 60, 30), rng.uniform(cy - 60, cy + 60, 30)
        seed(0, np.c_[np.repeat(px, 12), np.repeat(py, 12), np.tile(np.arange(12) * 0.4 - 1.5, 30)])
    for (kind, cube), pts in seeded.items():
        lom.map_set_cube(kind, cube, pts)
        pipe.lm.set_cube(kind, cube, pts)
    jumps = [None, (400.0, 0.0, 0.0), (400.0, 400.0, 0.0), (-400.0, 400.0, 130.0), (-400.0, -400.0, -130.0), (0.0, 0.0, 0.0)]
    expect_cen = [(10, 10, 5), (9, 10, 5), (9, 9, 5), (11, 9, 4), (11, 11, 6), (11, 11, 6)]
    solved = 0
    for k, scan in enumerate(scans):
        lom.reset(); lom.scanRegistrationIO(scan)
        lp = lom.laserOdometryIO()
        assert pipe.process(scan, do_mapping=False) == 0
        if jumps[k] is not None:
            pose = np.r_[lp["q_w_curr"][0], np.asarray(jumps[k])]      # the jump target is absolute: identical on both sides
            lom.set_lo_pose(pose[None])
            pipe.lo.set_pose(pose[:4], pose[4:])
        mp = lom.laserMappingIO()
        n_before = pipe.lm.map_points(1)
        pipe.lm.reset(); pipe.lm.input_from_lo(pipe.lo); pipe.lm.solve()
        ost = pipe.lm.state
        info = lom.lm_info()[0]
        assert tuple(ost["cen"]) == expect_cen[k], (k, ost["cen"])        # the oracle really shifted (the test's premise)
        if k == 1:   # ... and the shift towards -x dropped the seeded i = 0 plane (300 points) on its way
            assert pipe.lm.map_points(1) <= n_before + pipe.lm.cloud(1).shape[0] - 300
        assert list(info[:4]) == list(ost["cen"]) + [ost["validNum"]], (k, info, ost)
        _same_xyz(lom.cloud(V.CLOUD_CORNER_MAP), pipe.lm.cloud(2), f"scan {k} corner from map")
        _same_xyz(lom.cloud(V.CLOUD_SURF_MAP), pipe.lm.cloud(3), f"scan {k} surf from map")
        for p, t in enumerate(pipe.lm.trace()):
            g = lom.lm_trace(p)
            assert (g["n_corner"], g["n_plane"]) == (len(t["corner"]), len(t["plane"])), (k, p)
            assert np.array_equal(lom.lm_queries(p, 1), t["plane"].ravel())
            np.testing.assert_allclose(g["iterations"][: t["iterations"].shape[0], 0], t["iterations"][:, 0], rtol=1e-8, atol=1e-12)
            solved += 1
        assert np.max(np.abs(mp["t_w_curr"][0] - ost["t_w_curr"])) < POSE_TOL_M, k
        assert _quat_angle(mp["q_w_curr"][0], ost["q_w_curr"]) < POSE_TOL_RAD
        assert _compare_all_cubes(lom, pipe.lm, f"scan {k}") > 0
    assert solved >= 4          # the solve ran on shifted tables, not only the insertion
    lom.close()


def test_laser_mapping_bench_configuration(synth, oracle):
    """The configuration bench.py times (BASELINE configs[2]): full-size scans on the pre-built 1 M-point map, 2 passes x 5 LM
    iterations, several streams in one handle — poses, solver traces, query sets and every cube of the map against one
    oracle pipeline per stream, after every scan."""
    import vloam_b200 as V
    B, n_scans = 2, 4
    import bench
    cubes = synth.map_cubes(1_000_000, bench.BENCH_SEED)
    streams = [synth.ScanStream(bench.BENCH_SEED + b, n_cols=2048, yaw_rate_max=bench.BENCH_YAW_RATE_MAX, on_road=True) for b in range(B)]
    cap = 64 * 2048
    lom = V.LidarOdometryMapping(batch=B, max_points=cap, map_capacity_points=1 << 21, lm_max_iterations=5)
    pipes = [oracle.Pipeline() for _ in range(B)]
    for b in range(B):
        pipes[b].lm.set_iterations(2, 5)
        for (kind, cube), pts in cubes.items():
            lom.map_set_cube(kind, cube, pts, stream=b)
            pipes[b].lm.set_cube(kind, cube, pts)
    its = []
    for k in range(n_scans):
        buf = np.stack([st.scan(k) for st in streams])
        lom.reset(); lom.scanRegistrationIO(buf); lom.laserOdometryIO()
        mp = lom.laserMappingIO()
        assert not lom.lm_status().any()
        for b in range(B):
            assert pipes[b].process(buf[b], do_mapping=True) == 0
            ost = pipes[b].lm.state
            assert np.max(np.abs(mp["t_w_curr"][b] - ost["t_w_curr"])) < POSE_TOL_M, (k, b)
            assert _quat_angle(mp["q_w_curr"][b], ost["q_w_curr"]) < POSE_TOL_RAD
            otr = pipes[b].lm.trace()
            assert len(otr) == 2
            for p, t in enumerate(otr):
                g = lom.lm_trace(p, b)
                assert np.array_equal(lom.lm_queries(p, 0, b), t["corner"].ravel()), (k, b, p)
                assert np.array_equal(lom.lm_queries(p, 1, b), t["plane"].ravel()), (k, b, p)
                n = t["iterations"].shape[0]
                assert g["n_records"] == n and g["termination"] == t["termination"]
                np.testing.assert_allclose(g["iterations"][:n, 0], t["iterations"][:, 0], rtol=1e-8, atol=1e-12)
                np.testing.assert_array_equal(g["iterations"][:n, 5:7], t["iterations"][:, 5:7])
                assert len(t["plane"]) > 1000
                its.append(n - 1)
            assert _compare_all_cubes(lom, pipes[b].lm, f"scan {k} stream {b}", stream=b) >= 50
    print("LM iterations executed per pass (oracle == CUDA):", its)
    lom.close()


def test_laser_mapping_long_forward_sequence(synth, oracle):
    """60 consecutive scans of a forward trajectory (no replay; ~65 m of travel: the window moves to new cubes, voxels are
    inserted, slabs grow and the map is re-packed): pose within 1e-4 m / 1e-4 rad after EVERY scan, every cube of the map
    identical every 10 scans and at the end."""
    import vloam_b200 as V
    s = synth.ScanStream(34, n_cols=512)
    n_scans = 60
    lom = V.LidarOdometryMapping(batch=1, max_points=64 * 512, map_capacity_points=1 << 17)
    pipe = oracle.Pipeline()
    worst = 0.0
    cens = set()
    for k in range(n_scans):
        scan = s.scan(k)
        lom.reset(); lom.scanRegistrationIO(scan); lom.laserOdometryIO()
        mp = lom.laserMappingIO()
        assert pipe.process(scan, do_mapping=True) == 0
        ost = pipe.lm.state
        dt = float(np.max(np.abs(mp["t_w_curr"][0] - ost["t_w_curr"])))
        worst = max(worst, dt)
        assert dt < POSE_TOL_M, (k, dt)
        assert _quat_angle(mp["q_w_curr"][0], ost["q_w_curr"]) < POSE_TOL_RAD, k
        for p, t in enumerate(pipe.lm.trace()):
            g = lom.lm_trace(p)
            assert (g["n_corner"], g["n_plane"], g["termination"]) == (len(t["corner"]), len(t["plane"]), t["termination"]), (k, p)
        cens.add(tuple(ost["cen"]) + (int(_cube_of(ost["t_w_curr"], ost["cen"])),))
        if k % 10 == 9 or k == n_scans - 1:
            _compare_all_cubes(lom, pipe.lm, f"scan {k}")
        assert not lom.lm_status()[0].any()
    assert np.linalg.norm(pipe.lm.state["t_w_curr"]) > 30.0        # the trajectory really went somewhere
    assert len(cens) >= 2                                          # ... into another centre cube
    print("max |t_w_curr - oracle| over", n_scans, "scans =", worst)
    lom.close()


def test_laser_mapping_skip_frame(synth, oracle):
    """mapping_skip_frame = 2 (laser_odometry.cpp:618-628, laser_mapping.cpp:175-195, 742-756): every second frame only
    refreshes the high-frequency pose q_wmap_wodom * q_wodom_curr; the map and q_w_curr are untouched by it."""
    import vloam_b200 as V
    s = synth.ScanStream(35, n_cols=512)
    lom = V.LidarOdometryMapping(batch=1, max_points=64 * 512, map_capacity_points=1 << 17, mapping_skip_frame=2)
    pipe = oracle.Pipeline(mapping_skip_frame=2)
    skipped = 0
    for k in range(7):
        scan = s.scan(k)
        lom.reset(); lom.scanRegistrationIO(scan); lom.laserOdometryIO()
        before = lom.map_stats()[0, :, 0].copy()
        mp = lom.laserMappingIO()
        assert pipe.process(scan, do_mapping=True) == 0
        q, t, skip = pipe.lm.published_pose
        skipped += skip
        assert np.max(np.abs(mp["t_w_curr"][0] - t)) < POSE_TOL_M, (k, skip)
        assert _quat_angle(mp["q_w_curr"][0], q) < POSE_TOL_RAD, (k, skip)
        ost = pipe.lm.state
        assert np.max(np.abs(mp["t_wmap_wodom"][0] - ost["t_wmap_wodom"])) < POSE_TOL_M
        if skip:
            assert np.array_equal(lom.map_stats()[0, :, 0], before)       # a skipped frame inserts nothing
        _compare_all_cubes(lom, pipe.lm, f"scan {k}")
    assert skipped == 4          # frameCount is 1, 3, 5, 7 after the odometry of scans 0, 2, 4, 6: those frames are skipped
    lom.close()
    with pytest.raises(V.VloamError):
        V.LidarOdometryMapping(batch=1, max_points=4096, mapping_skip_frame=0)


def test_laser_mapping_capacity_overflow_is_per_kind_and_recoverable(synth, oracle):
    """A map pool too small for the surf map: that kind keeps its pre-insertion content and reports VLOAM_LM_SURF_MAP_FULL
    for the scan; the corner map still takes the scan, nothing is corrupted (every cube stays readable and indexed: the next
    scans keep solving), and the bit clears once a scan fits again."""
    import vloam_b200 as V
    s = synth.ScanStream(31, n_cols=1024)
    lom = V.LidarOdometryMapping(batch=1, max_points=64 * 1024, map_capacity_points=6000)
    full_seen = False
    for k in range(6):
        scan = s.scan(k)
        lom.reset(); lom.scanRegistrationIO(scan); lom.laserOdometryIO()
        st0 = lom.map_stats()[0].copy()
        mp = lom.laserMappingIO()
        status = lom.lm_status()[0]
        st1 = lom.map_stats()[0]
        assert np.isfinite(mp["t_w_curr"]).all()
        if status[0] & V.LM_SURF_MAP_FULL:
            full_seen = True
            assert st1[1, 0] == st0[1, 0]                      # the surf map kept its point count ...
            assert not (status[0] & V.LM_CORNER_MAP_FULL)
            assert st1[0, 0] >= st0[0, 0]                      # ... while the corner map took the scan
        for kind in (0, 1):                                     # tables stay consistent: counts add up, slabs inside the pool
            assert st1[kind, 0] <= st1[kind, 7] <= 6000
            if k == 5:      # every cube is still readable and the counts add up
                assert sum(lom.map_get_cube(kind, c).shape[0] for c in range(4851)) == st1[kind, 0]
        assert status[1] & status[0] == status[0]              # the sticky word accumulates the per-scan word
    assert full_seen, "the test's map capacity did not overflow: lower it"
    g = lom.lm_trace(1)
    assert g["n_plane"] > 100                                  # still solving against the (frozen) surf map
    lom.close()


def test_laser_mapping_large_cube_uses_flat_index(synth, oracle):
    """A cube of more than 65 535 points cannot use 16-bit z-layer offsets: its index falls back to column-only order with
    32-bit starts.  Results must not change: a dense seeded ground cube (70 k points), full parity with the oracle."""
    import vloam_b200 as V
    s = synth.ScanStream(36, n_cols=512)
    lom = V.LidarOdometryMapping(batch=1, max_points=64 * 512, map_capacity_points=1 << 18)
    pipe = oracle.Pipeline()
    rng = np.random.default_rng(3)
    g = np.arange(-24.9, 24.9, 0.186)
    gx, gy = np.meshgrid(g, g)
    surf = np.c_[gx.ravel(), gy.ravel(), np.full(gx.size, -1.73) + rng.uniform(-0.02, 0.02, gx.size), np.zeros(gx.size)].astype(np.float32)
    assert surf.shape[0] > 65535
    px, py = rng.uniform(-24, 24, 40), rng.uniform(-24, 24, 40)
    corner = np.c_[np.repeat(px, 12), np.repeat(py, 12), np.tile(np.arange(12) * 0.4 - 1.5, 40), np.zeros(480)].astype(np.float32)
    c0 = int(_cube_of(np.zeros(3)))
    for kind, pts in ((0, corner), (1, surf)):
        lom.map_set_cube(kind, c0, pts)
        pipe.lm.set_cube(kind, c0, pts)
    for k in range(3):
        scan = s.scan(k)
        lom.reset(); lom.scanRegistrationIO(scan); lom.laserOdometryIO()
        mp = lom.laserMappingIO()
        assert pipe.process(scan, do_mapping=True) == 0
        ost = pipe.lm.state
        assert np.max(np.abs(mp["t_w_curr"][0] - ost["t_w_curr"])) < POSE_TOL_M, k
        for p, t in enumerate(pipe.lm.trace()):
            g_ = lom.lm_trace(p)
            assert np.array_equal(lom.lm_queries(p, 1), t["plane"].ravel()), (k, p)
            assert np.array_equal(lom.lm_queries(p, 0), t["corner"].ravel()), (k, p)
            if k == 0:
                assert g_["n_plane"] > 200      # the first scan matches against the dense cube (still > 65 535 points: flat index)
        _compare_all_cubes(lom, pipe.lm, f"scan {k}")
    lom.close()


@pytest.mark.parametrize("batch", [1, 3])
def test_one_call_per_frame_with_cuda_graph_replay_matches_stage_calls(synth, batch):
    """vloam_lidar_process (reset + scanRegistrationIO + laserOdometryIO + laserMappingIO in one call, the frame's launch
    sequence replayed as a CUDA graph from the fourth frame on) must leave exactly the state the three stage calls leave:
    identical poses (bit for bit — same kernels, same order) and identical maps, host-buffer and device-buffer variants."""
    import torch
    import vloam_b200 as V
    streams = [synth.ScanStream(60 + b, n_cols=512) for b in range(batch)]
    cap = 64 * 512
    mk = lambda: V.LidarOdometryMapping(batch=batch, max_points=cap, map_capacity_points=1 << 17)   # noqa: E731
    ref, g_host, g_dev, nog = mk(), mk(), mk(), mk()
    n_dev = torch.full((batch,), cap, dtype=torch.int32, device="cuda")
    for k in range(8):
        buf = np.stack([s.scan(k) for s in streams])
        ref.reset(); ref.scanRegistrationIO(buf); lo_ref = ref.laserOdometryIO(); lm_ref = ref.laserMappingIO()
        g_host.process(buf, use_graph=True)
        xyz = torch.from_numpy(buf).cuda()
        torch.cuda.synchronize()
        g_dev.processDevice(xyz, n_dev, 3, cap, use_graph=True)
        nog.processDevice(xyz, n_dev, 3, cap, use_graph=False)
        for name, h in (("graph/host", g_host), ("graph/device", g_dev), ("direct/device", nog)):
            lo, lm = h.lo_pose(), h.lm_pose()
            for key in lo_ref:
                assert np.array_equal(lo[key], lo_ref[key]), (k, name, key)
            for key in lm_ref:
                assert np.array_equal(lm[key], lm_ref[key]), (k, name, key)
            assert np.array_equal(h.map_stats()[:, :, 0], ref.map_stats()[:, :, 0]), (k, name)
    # the maps themselves
    for kind in (0, 1):
        for cube in range(4851):
            a = ref.map_get_cube(kind, cube)
            if a.shape[0]:
                for h in (g_host, g_dev):
                    assert np.array_equal(a.view(np.uint32), h.map_get_cube(kind, cube).view(np.uint32)), (kind, cube)
    assert g_host.ctx.launch_count > 0
    for h in (ref, g_host, g_dev, nog):
        h.close()
