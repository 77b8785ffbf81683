"""GPU parity tests (run with -m gpu on the B200 box): CUDA path through the C-ABI vs the CPU oracle.

Bars (BASELINE.json north_star): identical edge / plane index sets; pose within 1e-4 m / 1e-4 rad after the
same iteration count.  Integer / index / float-bit work is compared bit-exactly — including `intensity`
(ring + 0.1 * relTime): the device computes azimuths with the C library's own atan2f algorithm
(csrc/fdlibm_atan2f.h, checked against glibc in tests/test_oracle_units.py), so relTime, the 2 pi unwrapping
decisions of scan_registration.cpp:236-262 and int(intensity) — the scan id laserOdometry reads — carry the
oracle's bits.  (Round 1 tolerated last-place atan2f differences and 1 % of 2 pi flips; the seeds 0-99 sweep showed
such a flip can change int(intensity).)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INTENSITY_TOL = 0.0          # bit-exact (kept as parameters of the helpers below)
INTENSITY_TOL_AVG = 0.0
POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-4


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _assert_cloud_equal(gpu, ref, name, tol=INTENSITY_TOL):
    assert gpu.shape == ref.shape, f"{name}: {gpu.shape} vs {ref.shape}"
    if gpu.shape[0] == 0:
        return
    assert np.array_equal(_bits(gpu[:, :3]), _bits(ref[:, :3])), f"{name}: xyz not bit-identical"
    assert np.array_equal(gpu[:, 3].astype(np.int32), ref[:, 3].astype(np.int32)), f"{name}: ring ids differ"
    bad = np.nonzero(_bits(gpu[:, 3]) != _bits(ref[:, 3]))[0]
    assert bad.size == 0, f"{name}: intensity differs at {bad.size} of {gpu.shape[0]} points, first {bad[:5]}: {gpu[bad[:5], 3]} vs {ref[bad[:5], 3]}"


def _check_sr(lom, ref, stream=0):
    import vloam_b200 as V
    assert lom.stream_status()[stream] == 0
    counts = lom.feature_counts()[stream]
    assert list(counts) == [ref.laserCloud.shape[0], ref.cornerPointsSharp.shape[0], ref.cornerPointsLessSharp.shape[0],
                            ref.surfPointsFlat.shape[0], ref.surfPointsLessFlat.shape[0]], counts
    _assert_cloud_equal(lom.cloud(V.CLOUD_FULL, stream), ref.laserCloud, "laserCloud")
    curv = lom.curvature(stream)
    assert np.array_equal(_bits(curv), _bits(ref.curvature)), "curvature not bit-identical"
    assert np.array_equal(lom.labels(stream).astype(np.int32), ref.label), "labels differ"
    assert np.array_equal(lom.feature_indices(V.CLOUD_SHARP, stream), ref.sharpInd)
    assert np.array_equal(lom.feature_indices(V.CLOUD_LESS_SHARP, stream), ref.lessSharpInd)
    assert np.array_equal(lom.feature_indices(V.CLOUD_FLAT, stream), ref.flatInd)
    _assert_cloud_equal(lom.cloud(V.CLOUD_SHARP, stream), ref.cornerPointsSharp, "sharp")
    _assert_cloud_equal(lom.cloud(V.CLOUD_LESS_SHARP, stream), ref.cornerPointsLessSharp, "lessSharp")
    _assert_cloud_equal(lom.cloud(V.CLOUD_FLAT, stream), ref.surfPointsFlat, "flat")
    _assert_cloud_equal(lom.cloud(V.CLOUD_LESS_FLAT, stream), ref.surfPointsLessFlat, "lessFlat", INTENSITY_TOL_AVG)


def _quat_angle(q1, q2):
    d = abs(float(np.dot(q1, q2)))
    return 2.0 * np.arccos(min(1.0, d))


def _check_lo(lom, olo, pose, stream=0, require_identical_sets=True):
    st = olo.state
    tr = olo.trace()
    for p, t in enumerate(tr):
        g = lom.lo_trace(p, stream)
        if require_identical_sets:
            assert np.array_equal(g["corner"], t["corner"]), f"pass {p}: corner correspondences differ"
            assert np.array_equal(g["plane"], t["plane"]), f"pass {p}: plane correspondences differ"
        assert g["termination"] == t["termination"], (p, g["termination"], t["termination"])
        n = t["iterations"].shape[0]
        assert g["n_records"] == n
        np.testing.assert_allclose(g["iterations"][:n, 0], t["iterations"][:, 0], rtol=1e-9, atol=1e-12)  # cost
        np.testing.assert_array_equal(g["iterations"][:n, 5:7], t["iterations"][:, 5:7])                 # valid / successful
        np.testing.assert_allclose(g["iterations"][:n, 4], t["iterations"][:, 4], rtol=1e-6)              # radius
        np.testing.assert_allclose(g["para"], t["para"], atol=1e-9)
    assert np.max(np.abs(pose["t_last_curr"][stream] - st["t_last_curr"])) < POSE_TOL_M
    assert _quat_angle(pose["q_last_curr"][stream], st["q_last_curr"]) < POSE_TOL_RAD
    assert np.max(np.abs(pose["t_w_curr"][stream] - st["t_w_curr"])) < POSE_TOL_M
    assert _quat_angle(pose["q_w_curr"][stream], st["q_w_curr"]) < POSE_TOL_RAD
    assert pose["corner_correspondence"][stream] == st["corner_correspondence"]
    assert pose["plane_correspondence"][stream] == st["plane_correspondence"]
    return float(np.max(np.abs(pose["t_last_curr"][stream] - st["t_last_curr"])))


@pytest.mark.parametrize("solver_mode", [1, 2])
def test_scan_registration_and_odometry_full_size(scans_full, oracle, solver_mode):
    """BASELINE configs[1] shape: 64 x 2048 scans, scanRegistration + laserOdometry, one stream; both layouts of the solve
    (vloam_lidar_params::solver_mode)."""
    import vloam_b200 as V
    scans, _ = scans_full
    lom = V.LidarOdometryMapping(batch=1, max_points=scans[0].shape[0], solver_mode=solver_mode)
    olo = oracle.LaserOdometry()
    worst = 0.0
    for k, scan in enumerate(scans):
        lom.reset()
        lom.scanRegistrationIO(scan)
        ref = oracle.scan_registration(scan)
        _check_sr(lom, ref)
        pose = lom.laserOdometryIO()
        olo.solve(ref)
        if k > 0:
            worst = max(worst, _check_lo(lom, olo, pose))
        else:
            assert np.allclose(pose["q_w_curr"][0], [0, 0, 0, 1]) and np.allclose(pose["t_w_curr"][0], 0)
        # the swap (laser_odometry.cpp:511-517): corner/surf "last" are now this scan's less-sharp / less-flat
        _assert_cloud_equal(lom.cloud(V.CLOUD_CORNER_LAST), ref.cornerPointsLessSharp, "cornerLast")
        _assert_cloud_equal(lom.cloud(V.CLOUD_SURF_LAST), ref.surfPointsLessFlat, "surfLast", INTENSITY_TOL_AVG)
    print("max |t_last_curr - oracle| =", worst)
    lom.close()


def test_batched_ragged_streams(synth, oracle):
    """Three independent streams of different seeds and lengths in one handle == three oracle runs."""
    import vloam_b200 as V
    streams = [synth.ScanStream(seed, n_cols=c) for seed, c in ((3, 512), (4, 384), (5, 640))]
    B = len(streams)
    cap = max(s.n_cols for s in streams) * 64
    lom = V.LidarOdometryMapping(batch=B, max_points=cap)
    olos = [oracle.LaserOdometry() for _ in range(B)]
    for k in range(3):
        buf = np.full((B, cap, 4), np.nan, np.float32)  # pcl::PointXYZ stride (4 floats)
        n = np.zeros(B, np.int32)
        scans = []
        for b, s in enumerate(streams):
            sc = s.scan(k)
            scans.append(sc)
            buf[b, : sc.shape[0], :3] = sc
            n[b] = sc.shape[0]
        lom.reset()
        lom.scanRegistrationIO(buf, n)
        pose = lom.laserOdometryIO()
        for b in range(B):
            ref = oracle.scan_registration(scans[b])
            _check_sr(lom, ref, b)
            olos[b].solve(ref)
            if k > 0:
                _check_lo(lom, olos[b], pose, b)
    lom.close()


def test_edge_cases(synth, oracle):
    """Empty scan, all-NaN scan, scan with every point inside minimum_range, tiny rings."""
    import vloam_b200 as V
    lom = V.LidarOdometryMapping(batch=1, max_points=4096)
    # all NaN
    lom.reset()
    lom.scanRegistrationIO(np.full((1000, 3), np.nan, np.float32))
    assert lom.stream_status()[0] & V.STREAM_EMPTY
    assert list(lom.feature_counts()[0]) == [0, 0, 0, 0, 0]
    assert oracle.scan_registration(np.full((1000, 3), np.nan, np.float32)).status == 1
    # zero points
    lom.reset()
    lom.scanRegistrationIO(np.zeros((16, 3), np.float32), np.zeros(1, np.int32))
    assert lom.stream_status()[0] & V.STREAM_EMPTY
    # all closer than minimum_range
    lom.reset()
    near = np.random.default_rng(0).normal(0, 1.0, (2000, 3)).astype(np.float32)
    lom.scanRegistrationIO(near)
    assert lom.stream_status()[0] & V.STREAM_EMPTY
    # short rings: only 10..40 points per ring -> most rings skipped by the `< 6` rule (scan_registration.cpp:314)
    s = synth.ScanStream(11, n_cols=40)
    sc = s.scan(0)
    lom.reset()
    lom.scanRegistrationIO(sc)
    ref = oracle.scan_registration(sc)
    _check_sr(lom, ref)
    # laser odometry on nearly empty features must not crash and must agree with the oracle
    pose0 = lom.laserOdometryIO()
    olo = oracle.LaserOdometry()
    olo.solve(ref)
    sc1 = s.scan(1)
    lom.reset()
    lom.scanRegistrationIO(sc1)
    ref1 = oracle.scan_registration(sc1)
    _check_sr(lom, ref1)
    pose1 = lom.laserOdometryIO()
    olo.solve(ref1)
    _check_lo(lom, olo, pose1)
    lom.close()


def test_call_order_errors():
    import vloam_b200 as V
    lom = V.LidarOdometryMapping(batch=1, max_points=2048)
    with pytest.raises(V.VloamError):
        lom.laserOdometryIO()  # before any scan registration
    lom.scanRegistrationIO(np.full((100, 3), np.nan, np.float32))
    lom.laserOdometryIO()
    with pytest.raises(V.VloamError):
        lom.laserOdometryIO()  # twice for the same scan
    with pytest.raises(V.VloamError):
        lom.scanRegistrationIO(np.zeros((4096, 3), np.float32))  # larger than max_points
    lom.close()


def test_pipelined_host_api_matches_blocking(synth):
    """One scan in flight (upload on the copy stream, pose of scan k-1 read while k runs) == blocking calls."""
    import vloam_b200 as V
    s = synth.ScanStream(21, n_cols=256)
    scans = [s.scan(k) for k in range(4)]
    a = V.LidarOdometryMapping(batch=1, max_points=scans[0].shape[0])
    b = V.LidarOdometryMapping(batch=1, max_points=scans[0].shape[0])
    blocking, piped = [], []
    for k, sc in enumerate(scans):
        a.reset(); a.scanRegistrationIO(sc); blocking.append(a.laserOdometryIO())
        b.reset(); b.scanRegistrationIO(sc); b.laserOdometryIO(fetch=False)
        if k > 0:
            piped.append(b.lo_pose(prev=True))
    piped.append(b.lo_pose())
    for p, q in zip(blocking, piped):
        for key in p:
            assert np.array_equal(p[key], q[key]), key
    a.close(); b.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_one_host_buffer_per_stream_upload(synth, pinned):
    """vloam_scan_registration_ptrs: pinned buffers go to the driver as one batched copy (cudaMemcpyBatchAsync), pageable ones
    as one cudaMemcpyAsync each; ragged counts; every stream's cloud and pose equal the slab upload's."""
    import torch
    import vloam_b200 as V
    B, n_cols = 5, 256
    streams = [synth.ScanStream(300 + i, n_cols=n_cols) for i in range(B)]
    cap = 64 * n_cols
    a = V.LidarOdometryMapping(batch=B, max_points=cap)
    b = V.LidarOdometryMapping(batch=B, max_points=cap)
    n = np.array([cap, cap - 700, cap, 0, cap - 64], np.int32)        # ragged, one empty stream
    for k in range(3):
        scans = np.stack([st.scan(k) for st in streams]).astype(np.float32)
        bufs = [torch.from_numpy(scans[i].copy()) for i in range(B)]
        if pinned:
            bufs = [t.pin_memory() for t in bufs]
        ptrs = np.array([t.data_ptr() for t in bufs], np.uint64)
        a.reset(); a.scanRegistrationIO(scans, n_points=n); pa = a.laserOdometryIO()
        b.reset(); b.scanRegistrationPtrs(ptrs, n, 3, keep=bufs); pb = b.laserOdometryIO()
        for i in range(B):
            if n[i]:
                assert np.array_equal(a.cloud(V.CLOUD_FULL, stream=i), b.cloud(V.CLOUD_FULL, stream=i)), (k, i)
        assert np.array_equal(a.stream_status(), b.stream_status())
        for key in pa:
            assert np.array_equal(pa[key], pb[key]), (k, key)
    a.close(); b.close()


def test_wire_formats_feed_scan_registration(synth, oracle, tmp_path):
    """a KITTI .bin record stream (4 floats per point) and a sensor_msgs/PointCloud2 payload with
    point_step = 32 bytes go into scanRegistrationIO as they are (no pcl::fromROSMsg copy, vloam_main_node.cpp:148) and
    give the same features as the packed cloud."""
    import vloam_b200 as V
    from vloam_b200 import wire
    sc = synth.ScanStream(17, n_cols=512).scan(0)
    ref = oracle.scan_registration(sc)
    n = sc.shape[0]
    # KITTI .bin: x y z reflectance
    path = tmp_path / "000000.bin"
    np.c_[sc, np.full(n, 0.5, np.float32)].astype(np.float32).tofile(str(path))
    kitti = wire.load_kitti_bin(str(path))
    # PointCloud2: 32-byte records, x y z at offset 0, intensity at 16, padding elsewhere
    msg = np.zeros((n, 32), np.uint8)
    msg[:, 0:12] = sc.copy().view(np.uint8).reshape(n, 12)
    msg[:, 16:20] = np.full((n, 1), 0.25, np.float32).view(np.uint8)
    pc2, zero_copy = wire.pointcloud2_xyz(msg.reshape(-1), n, 32, {"x": 0, "y": 4, "z": 8})
    assert zero_copy and pc2.shape == (n, 8)
    lom = V.LidarOdometryMapping(batch=1, max_points=n)
    for cloud in (kitti, pc2, sc):
        lom.reset()
        lom.scanRegistrationIO(cloud)
        _check_sr(lom, ref)
        lom.laserOdometryIO()
    lom.close()


@pytest.mark.parametrize("scan_line", [16, 32])
def test_other_sensors_ring_formulas(synth, oracle, scan_line):
    """the VLP-16 / HDL-32 ring formulas (scan_registration.cpp:195-212), selected by the
    `scan_line` parameter like the reference's launch files.  The same synthetic sweep is classified with each formula;
    scan registration and laser odometry must agree with the oracle run with the same setting."""
    import vloam_b200 as V
    s = synth.ScanStream(23, n_cols=512)
    lom = V.LidarOdometryMapping(batch=1, max_points=64 * 512, scan_line=scan_line)
    olo = oracle.LaserOdometry()
    for k in range(2):
        sc = s.scan(k)
        lom.reset()
        lom.scanRegistrationIO(sc)
        ref = oracle.scan_registration(sc, n_scans=scan_line)
        assert ref.laserCloud.shape[0] > 1000
        assert int(ref.laserCloud[:, 3].max()) <= scan_line - 1
        _check_sr(lom, ref)
        pose = lom.laserOdometryIO()
        olo.solve(ref)
        if k > 0:
            _check_lo(lom, olo, pose)
    lom.close()


@pytest.mark.parametrize("seed0", [0, 25, 50, 75])
def test_seed_sweep_scan_registration_and_odometry(synth, oracle, seed0):
    """seeds 0..99 x 2 consecutive scans (64 x 1024) — curvature, labels, every feature index set and
    every cloud bit-exact, laser-odometry correspondences identical and poses within 1e-4 on the second scan.  Five streams
    per handle, so the sweep also covers the batched layout."""
    import vloam_b200 as V
    B, cols = 5, 1024
    cap = 64 * cols
    lom = V.LidarOdometryMapping(batch=B, max_points=cap)
    for base in range(seed0, seed0 + 25, B):
        streams = [synth.ScanStream(base + b, n_cols=cols) for b in range(B)]
        olos = [oracle.LaserOdometry() for _ in range(B)]
        for k in range(2):
            buf = np.stack([st.scan(k) for st in streams])
            lom.reset()
            lom.scanRegistrationIO(buf)
            refs = [oracle.scan_registration(buf[b]) for b in range(B)]
            for b in range(B):
                _check_sr(lom, refs[b], b)
            pose = lom.laserOdometryIO()
            for b in range(B):
                olos[b].solve(refs[b])
                if k > 0:
                    _check_lo(lom, olos[b], pose, b)
        lom.init()      # the next group of seeds starts from a fresh handle, like its fresh oracles
    lom.close()


def test_device_resident_count_beyond_capacity_is_clamped_and_reported(synth):
    """vloam_scan_registration_device cannot validate counts that live in device memory: a count above the handle
    capacity must be clamped inside the kernels and reported per stream, not overrun the buffers."""
    import torch
    import vloam_b200 as V
    cols = 256
    cap = 64 * cols
    sc = synth.ScanStream(12, n_cols=cols).scan(0)
    big = np.concatenate([sc, sc, sc])                             # 3 x capacity points in the caller's buffer
    lom = V.LidarOdometryMapping(batch=2, max_points=cap)
    xyz = torch.from_numpy(np.stack([big, big])).cuda()
    n = torch.tensor([cap, 3 * cap], dtype=torch.int32).cuda()    # stream 1 claims three times the capacity
    torch.cuda.synchronize()
    lom.reset()
    lom.scanRegistrationDevice(xyz, n, 3, 3 * cap)
    st = lom.stream_status()
    assert st[0] == 0 and st[1] == V.STREAM_CAPACITY, st
    c = lom.feature_counts()
    assert list(c[0]) == list(c[1])                                # the clamped stream saw exactly the first `cap` points
    lom.laserOdometryIO()
    lom.close()
    # the VO handle rounds its capacity differently from the lidar handle: a larger count is clamped, not overrun
    vo = V.VisualOdometry(batch=1, max_points=cap, max_matches=64)
    vo.setUpPointCloud(*synth.kitti_like_calibration())
    vo.reset()
    vo.processPointCloudDevice(xyz[1:], n[1:], 3, 3 * cap)
    ref = V.VisualOdometry(batch=1, max_points=cap, max_matches=64)
    ref.setUpPointCloud(*synth.kitti_like_calibration())
    ref.reset()
    ref.processPointCloud(sc)
    for a, b in zip(vo.buckets(0), ref.buckets(0)):
        assert np.array_equal(a, b)
    vo.close(); ref.close()


def test_motion_distortion_path(synth, oracle):
    """laser_odometry.h:90 DISTORTION == true (a compile-time constant of the reference, false as shipped):
    every feature is placed inside the sweep by s = frac(intensity) / 0.1 — TransformToStart through
    Identity.slerp(s, q_last_curr) (laser_odometry.cpp:149-167) for the association, and the same interpolation inside
    LidarEdgeFactor / LidarPlaneFactor (lidarFactor.hpp:28-35, 78-83) for the solve, whose Jacobians go through the slerp.
    Correspondences identical, cost trace and poses equal to the oracle with its switch on."""
    import vloam_b200 as V
    s = synth.ScanStream(27, n_cols=1024)
    lom = V.LidarOdometryMapping(batch=1, max_points=64 * 1024, distortion=1)
    olo = oracle.LaserOdometry()
    olo.set_distortion(True)
    plain = oracle.LaserOdometry()
    moved = 0.0
    for k in range(4):
        sc = s.scan(k)
        lom.reset()
        lom.scanRegistrationIO(sc)
        ref = oracle.scan_registration(sc)
        pose = lom.laserOdometryIO()
        olo.solve(ref)
        plain.solve(ref)
        if k > 0:
            _check_lo(lom, olo, pose)
            moved = max(moved, float(np.max(np.abs(olo.state["t_last_curr"] - plain.state["t_last_curr"]))))
    assert moved > 1e-3          # the switch really changes the estimate (the test is not vacuous)
    lom.close()
