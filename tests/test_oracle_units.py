"""CPU tests (-m "not gpu"): pin the oracle's building blocks against independent restatements.

The reference ships no tests or golden vectors and its third-party arithmetic (PCL, FLANN,
Ceres, Eigen) cannot be run here, so the oracle is validated piece by piece against numpy / scipy:
voxel grid vs a numpy restatement, kNN vs brute force and scipy.cKDTree, small dense linear algebra vs numpy,
auto-diff (dual number) Jacobians of the literal functors vs finite differences and vs the analytic Jacobians the
CUDA kernels use, the Levenberg-Marquardt loop vs scipy.optimize.least_squares, and scan-to-scan odometry vs the
generator's ground-truth motion.
"""
import os

import numpy as np
import pytest


def rand_quat(rng, angle=0.3):
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    return np.r_[np.sin(angle / 2) * ax, np.cos(angle / 2)]


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


# ------------------------------------------------------------------------------------------------ voxel grid
def numpy_voxel_grid(pts, leaf):
    """Independent restatement of pcl::VoxelGrid (float32 arithmetic, stable order inside a voxel)."""
    f32 = np.float32
    inv = f32(1.0) / f32(leaf)
    xyz = pts[:, :3]
    mn, mx = xyz.min(0), xyz.max(0)
    minb = np.floor(mn * inv).astype(np.int64)
    maxb = np.floor(mx * inv).astype(np.int64)
    div = maxb - minb + 1
    ijk = (np.floor(xyz * inv) - minb.astype(f32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(key, kind="stable")
    out = []
    ks = key[order]
    start = 0
    for end in list(np.nonzero(np.diff(ks))[0] + 1) + [len(ks)]:
        acc = np.zeros(4, f32)
        for i in order[start:end]:
            acc = (acc + pts[i]).astype(f32)
        out.append(acc / f32(end - start))
        start = end
    return np.array(out, f32)


@pytest.mark.parametrize("leaf", [0.2, 0.4, 0.8])
def test_voxel_grid_matches_numpy(oracle, leaf):
    rng = np.random.default_rng(1)
    pts = np.c_[rng.uniform(-20, 20, (3000, 2)), rng.uniform(-2, 3, 3000), rng.uniform(0, 50, 3000)].astype(np.float32)
    got = oracle.voxel_grid(pts, leaf)
    ref = numpy_voxel_grid(pts, leaf)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_voxel_grid_literal_std_sort_within_ulps(oracle):
    """PCL sorts on the key only (unstable): the summation order differs, the centroids agree to a few ulp."""
    rng = np.random.default_rng(2)
    pts = np.c_[rng.uniform(-10, 10, (5000, 3)), rng.uniform(0, 60, 5000)].astype(np.float32)
    a = oracle.voxel_grid(pts, 0.8, literal_unstable=False)
    b = oracle.voxel_grid(pts, 0.8, literal_unstable=True)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(a))) < 1e-6


def test_voxel_grid_edge_cases(oracle):
    assert oracle.voxel_grid(np.zeros((0, 4), np.float32), 0.2).shape == (0, 4)
    one = np.array([[1, 2, 3, 4]], np.float32)
    assert np.array_equal(oracle.voxel_grid(one, 0.2), one)
    # index overflow guard: PCL returns the input unfiltered
    far = np.array([[0, 0, 0, 1], [1e5, 1e5, 1e5, 2]], np.float32)
    assert np.array_equal(oracle.voxel_grid(far, 0.01), far)


# ------------------------------------------------------------------------------------------------ kNN
def test_kdtree_matches_brute_and_scipy(oracle):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(3)
    tgt = np.c_[rng.uniform(-30, 30, (4000, 3)), np.zeros(4000)].astype(np.float32)
    qry = np.c_[rng.uniform(-32, 32, (500, 3)), np.zeros(500)].astype(np.float32)
    for k in (1, 5):
        i_tree, d_tree = oracle.knn(tgt, qry, k)
        i_brute, d_brute = oracle.knn(tgt, qry, k, brute=True)
        assert np.array_equal(i_tree, i_brute)
        assert np.array_equal(d_tree.view(np.uint32), d_brute.view(np.uint32))
        d_sp, i_sp = cKDTree(tgt[:, :3].astype(np.float64)).query(qry[:, :3].astype(np.float64), k=k)
        assert np.array_equal(i_tree.reshape(i_sp.shape), i_sp)
        np.testing.assert_allclose(np.sqrt(d_tree.reshape(d_sp.shape)), d_sp, rtol=1e-5)


def test_knn_fewer_points_than_k(oracle):
    tgt = np.array([[0, 0, 0, 0], [1, 0, 0, 0]], np.float32)
    idx, d = oracle.knn(tgt, np.array([[0.1, 0, 0, 0]], np.float32), 5)
    assert list(idx[0]) == [0, 1, -1, -1, -1]


# ------------------------------------------------------------------------------------------------ small linear algebra
def test_sym_eig3_and_lstsq(oracle):
    rng = np.random.default_rng(4)
    for _ in range(50):
        m = rng.normal(size=(5, 3))
        A = m.T @ m
        ev, evec = oracle.sym_eig3(A)
        w, v = np.linalg.eigh(A)
        np.testing.assert_allclose(ev, w, rtol=1e-12, atol=1e-12)
        for k in range(3):
            assert abs(abs(evec[k] @ v[:, k]) - 1) < 1e-9
        P = rng.normal(size=(5, 3)) + np.array([20.0, -5.0, 1.0])
        x = oracle.colpiv_qr_solve_5x3(P, -np.ones(5))
        np.testing.assert_allclose(x, np.linalg.lstsq(P, -np.ones(5), rcond=None)[0], rtol=1e-9, atol=1e-12)


# ------------------------------------------------------------------------------------------------ factors
def analytic_edge(p, a, b, x):
    """The closed form the CUDA kernel uses (lo_kernels.cu edge_block)."""
    q, t = x[:4], x[4:]
    Rp = quat_to_R(q) @ p
    lp = Rp + t
    d = a - b
    den = np.linalg.norm(d)
    r = np.cross(lp - a, lp - b) / den
    e = d / den
    A = -np.array([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]])
    G = -2 * np.array([[0, -Rp[2], Rp[1]], [Rp[2], 0, -Rp[0]], [-Rp[1], Rp[0], 0]])
    return r, np.c_[A @ G, A]


def analytic_plane(p, n, d0, x):
    q, t = x[:4], x[4:]
    Rp = quat_to_R(q) @ p
    r = n @ (Rp + t) + d0
    return np.array([r]), np.r_[-2 * np.cross(n, Rp), n][None, :]


def local_jacobian(J7, q):
    """global (n x 7) * EigenQuaternionParameterization::ComputeJacobian (4 x 3), translation block unchanged."""
    x, y, z, w = q
    P = np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]])
    return np.c_[J7[:, :4] @ P, J7[:, 4:]]


def test_factor_autodiff_vs_finite_difference_and_analytic(oracle):
    rng = np.random.default_rng(5)
    for _ in range(20):
        q = rand_quat(rng, rng.uniform(0.0, 0.5))
        t = rng.normal(size=3)
        x = np.r_[q, t]
        p, a, b, c = (rng.uniform(-20, 20, 3) for _ in range(4))
        # edge
        r, J = oracle.factor_eval(0, np.r_[p, a, b], x)
        num = np.zeros((3, 7))
        for k in range(7):
            dx = np.zeros(7); dx[k] = 1e-6
            num[:, k] = (oracle.factor_eval(0, np.r_[p, a, b], x + dx)[0] - oracle.factor_eval(0, np.r_[p, a, b], x - dx)[0]) / 2e-6
        np.testing.assert_allclose(J, num, rtol=1e-5, atol=1e-5)
        ra, Ja = analytic_edge(p, a, b, x)
        np.testing.assert_allclose(r, ra, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(local_jacobian(J, q), Ja, rtol=1e-10, atol=1e-10)
        # plane (LidarPlaneFactor): normal from the three last points
        r, J = oracle.factor_eval(1, np.r_[p, a, b, c], x)
        n = np.cross(a - b, a - c); n /= np.linalg.norm(n)
        ra, Ja = analytic_plane(p, n, -n @ a, x)
        np.testing.assert_allclose(r, ra, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(local_jacobian(J, q), Ja, rtol=1e-10, atol=1e-10)
        # plane-norm (LidarPlaneNormFactor)
        r, J = oracle.factor_eval(2, np.r_[p, n, 0.7], x)
        ra, Ja = analytic_plane(p, n, 0.7, x)
        np.testing.assert_allclose(r, ra, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(local_jacobian(J, q), Ja, rtol=1e-10, atol=1e-10)


def test_vo_factor_autodiff_vs_finite_difference(oracle):
    rng = np.random.default_rng(6)
    for kind, nobs in ((0, 5), (1, 4)):
        for aa_scale in (0.0, 1e-9, 0.2):
            x = np.r_[rng.normal(size=3) * aa_scale, rng.normal(size=3)]
            obs = rng.uniform(-1, 1, nobs)
            if kind == 0:
                obs[2] = 8.0
            r, J = oracle.vo_factor_eval(kind, obs, x)
            num = np.zeros_like(J)
            for k in range(6):
                dx = np.zeros(6); dx[k] = 1e-6
                num[:, k] = (oracle.vo_factor_eval(kind, obs, x + dx)[0] - oracle.vo_factor_eval(kind, obs, x - dx)[0]) / 2e-6
            np.testing.assert_allclose(J, num, rtol=2e-5, atol=2e-5)


# ------------------------------------------------------------------------------------------------ scan registration properties
def test_scan_registration_properties(oracle, scans_small):
    scans, _ = scans_small
    r = oracle.scan_registration(scans[0])
    n = r.laserCloud.shape[0]
    assert r.status == 0 and n > 10000
    rings = r.laserCloud[:, 3].astype(int)
    assert rings.max() <= 50                          # rings above 50 dropped
    # xyz of the cloud are exactly the input points, ring-major and order-preserving inside a ring
    src = scans[0][np.isfinite(scans[0][:, 0])]
    assert set(map(bytes, r.laserCloud[:100, :3])) <= set(map(bytes, src))
    # labels: at most 2 sharp / 20 less sharp / 4 flat per sector
    for ring in range(51):
        s, e = r.scanStartInd[ring], r.scanEndInd[ring]
        if e - s < 6:
            continue
        for j in range(6):
            sp, ep = s + (e - s) * j // 6, s + (e - s) * (j + 1) // 6 - 1
            lab = r.label[sp:ep + 1]
            assert (lab == 2).sum() <= 2 and (lab >= 1).sum() <= 20 and (lab == -1).sum() <= 4
    assert np.array_equal(r.label[r.sharpInd], np.full(len(r.sharpInd), 2))
    assert np.all(r.curvature[r.sharpInd] > 0.1) and np.all(r.curvature[r.flatInd] < 0.1)
    assert len(set(r.lessSharpInd)) == len(r.lessSharpInd)
    # the less-flat cloud is ring-major
    assert np.all(np.diff(np.floor(r.surfPointsLessFlat[:, 3] + 0.5).astype(int)) >= -1)
    assert r.ringLessFlatCount.sum() == r.surfPointsLessFlat.shape[0]


def test_scan_registration_strides_and_nan_prefix(oracle, scans_small):
    scans, _ = scans_small
    sc = scans[1].copy()
    a = oracle.scan_registration(sc)
    padded = np.c_[sc, np.full(len(sc), 123.0, np.float32)]      # pcl::PointXYZ stride
    b = oracle.scan_registration(padded)
    assert np.array_equal(a.laserCloud, b.laserCloud) and np.array_equal(a.flatInd, b.flatInd)
    assert oracle.scan_registration(np.full((50, 3), np.nan, np.float32)).status == 1


# ------------------------------------------------------------------------------------------------ LM solver vs scipy
def test_lm_solver_matches_scipy_on_odometry_problem(oracle, scans_small):
    """Converged pose of the oracle's Ceres-style LM == scipy least_squares(huber) on the same residuals."""
    from scipy.optimize import least_squares
    scans, stream = scans_small
    lo = oracle.LaserOdometry()
    r0, r1 = oracle.scan_registration(scans[0]), oracle.scan_registration(scans[1])
    lo.solve(r0)
    lo.set_iterations(1, 60)        # one association pass, LM run to convergence
    lo.solve(r1)
    tr = lo.trace()[0]
    assert tr["termination"] in (1, 2, 3)
    x_or = tr["para"]
    corner, plane = tr["corner"], tr["plane"]
    CL, SL = r0.cornerPointsLessSharp, r0.surfPointsLessFlat
    P, F = r1.cornerPointsSharp, r1.surfPointsFlat

    def residuals(x6):
        ang = np.linalg.norm(x6[:3])
        q = np.r_[np.sin(ang / 2) * x6[:3] / ang, np.cos(ang / 2)] if ang > 0 else np.array([0, 0, 0, 1.0])
        R, t = quat_to_R(q), x6[3:]
        out = []
        for i, a, b in corner:
            lp = R @ P[i, :3].astype(np.float64) + t
            A, B = CL[a, :3].astype(np.float64), CL[b, :3].astype(np.float64)
            out.append(np.linalg.norm(np.cross(lp - A, lp - B)) / np.linalg.norm(A - B))  # Huber acts on the block norm
        for i, j, l, m in plane:
            lp = R @ F[i, :3].astype(np.float64) + t
            J, L, M = (SL[k, :3].astype(np.float64) for k in (j, l, m))
            n = np.cross(J - L, J - M); n /= np.linalg.norm(n)
            out.append((lp - J) @ n)
        return np.array(out)

    q = x_or[:4]
    ang = 2 * np.arccos(np.clip(q[3], -1, 1))
    x0 = np.r_[q[:3] / max(np.sin(ang / 2), 1e-12) * ang, x_or[4:]]
    sol = least_squares(residuals, np.zeros(6), loss="huber", f_scale=0.1, xtol=1e-12, ftol=1e-12, gtol=1e-12)
    # Ceres stops on function_tolerance = 1e-6 (relative cost change), i.e. slightly before the exact minimum:
    # compare the robust cost (same definition in both: 1/2 sum rho) and the pose with matching slack.
    assert tr["iterations"][-1, 0] <= sol.cost * (1 + 2e-5)
    np.testing.assert_allclose(sol.x[3:], x0[3:], atol=3e-3)
    np.testing.assert_allclose(sol.x[:3], x0[:3], atol=3e-4)
    assert abs(residuals(x0).size - (len(corner) + len(plane))) == 0


def test_laser_odometry_tracks_ground_truth(oracle, scans_small):
    scans, stream = scans_small
    lo = oracle.LaserOdometry()
    for k, sc in enumerate(scans):
        lo.solve(oracle.scan_registration(sc))
        if k >= 2:   # after the first solve the motion prior is warm
            R, t = stream.relative_pose(k)
            st = lo.state
            assert np.linalg.norm(st["t_last_curr"] - t) < 0.12
            assert np.linalg.norm(quat_to_R(st["q_last_curr"]) - R) < 0.02
    assert lo.state["frameCount"] == len(scans)


def test_laser_mapping_runs_and_refines(oracle, scans_small):
    scans, stream = scans_small
    pipe = oracle.Pipeline()
    for sc in scans:
        assert pipe.process(sc, do_mapping=True) == 0
    st = pipe.lm.state
    R, t = stream.pose(len(scans) - 1)
    assert np.linalg.norm(st["t_w_curr"] - t) < 0.3
    assert pipe.lm.map_points(0) > 100 and pipe.lm.map_points(1) > 1000
    tr = pipe.lm.trace()
    assert len(tr) == 2 and tr[0]["iterations"][0, 0] >= tr[-1]["iterations"][-1, 0]   # cost does not increase


def test_oracle_grid_shift_equals_array_shift(oracle):
    """laser_mapping.cpp:218-402 moves 21 x 21 x 11 cloud pointers one plane per while-iteration and clears the plane that
    wraps around.  Independent model: a (k, j, i) integer array shifted with slice assignment.  Every cube holds one
    point whose intensity is its original index, so the oracle's array after the shift can be read back as that integer array."""
    W, H, D = 21, 21, 11
    rng = np.random.default_rng(11)
    cases = [(400.0, 0.0, 0.0), (-400.0, 0.0, 0.0), (0.0, 460.0, 0.0), (0.0, -460.0, 0.0), (0.0, 0.0, 130.0), (0.0, 0.0, -180.0),
             (600.0, -520.0, 140.0), (24.9, -24.9, 0.0), (140.0, 140.0, 0.0)]
    for tx, ty, tz in cases:
        lm = oracle.LaserMapping()
        ids = np.arange(W * H * D).reshape(D, H, W)           # ids[k, j, i] = i + 21 j + 441 k
        occupied = rng.random(ids.shape) < 0.5
        for c in ids[occupied]:
            i, j, k = c % W, (c // W) % H, c // (W * H)
            p = np.array([[(i - 10) * 50.0, (j - 10) * 50.0, (k - 5) * 50.0, float(c)]], np.float32)   # the cube's own centre
            lm.set_cube(1, int(c), p)
        lm.input_clouds(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), [0, 0, 0, 1.0], [tx, ty, tz])
        lm.solve()
        # model
        model = np.where(occupied, ids, -1)
        cen = [10, 10, 5]
        def coord(v, c):
            r = int((v + 25.0) / 50.0) + c
            return r - 1 if v + 25.0 < 0 else r
        cc = [coord(tx, 10), coord(ty, 10), coord(tz, 5)]
        for axis, n in ((2, W), (1, H), (0, D)):              # numpy axis of i, j, k
            a = 2 - axis
            while cc[a] < 3:
                model = np.roll(model, 1, axis=axis)
                sl = [slice(None)] * 3; sl[axis] = 0
                model[tuple(sl)] = -1
                cc[a] += 1; cen[a] += 1
            while cc[a] >= n - 3:
                model = np.roll(model, -1, axis=axis)
                sl = [slice(None)] * 3; sl[axis] = n - 1
                model[tuple(sl)] = -1
                cc[a] -= 1; cen[a] -= 1
        assert list(lm.state["cen"]) == cen, (tx, ty, tz)
        got = np.full(ids.shape, -1)
        for c in range(W * H * D):
            n = lm.cube_count(1, c)
            assert n <= 1
            if n:
                got[c // (W * H), (c // W) % H, c % W] = int(lm.cube(1, c)[0, 3])
        assert np.array_equal(got, model), (tx, ty, tz)


def test_oracle_skip_frame_publishes_high_frequency_pose(oracle, scans_small):
    """mapping_skip_frame = 2: LaserOdometry::output flags frames with frameCount % 2 != 0 (laser_odometry.cpp:618-628); for
    those LaserMapping::input only forms q_wmap_wodom * q_wodom_curr (laser_mapping.cpp:186-190) and solveMapping is not
    run (lidar_odometry_mapping.cpp:134-135): the map must not change and the published pose must be that composition."""
    scans, _ = scans_small
    pipe = oracle.Pipeline(mapping_skip_frame=2)

    def qmul(a, b):
        ax, ay, az, aw = a; bx, by, bz, bw = b
        return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                         aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])

    def qrot(q, v):
        u = q[:3]
        return v + 2.0 * np.cross(u, np.cross(u, v) + q[3] * v)
    flags = []
    for k, sc in enumerate(scans):
        before = (pipe.lm.map_points(0), pipe.lm.map_points(1))
        st_before = pipe.lm.state
        assert pipe.process(sc, do_mapping=True) == 0
        q, t, skip = pipe.lm.published_pose
        flags.append(skip)
        lo, st = pipe.lo.state, pipe.lm.state
        if skip:
            assert (pipe.lm.map_points(0), pipe.lm.map_points(1)) == before
            np.testing.assert_array_equal(st["q_w_curr"], st_before["q_w_curr"])      # q_w_curr itself is not touched
            np.testing.assert_allclose(q, qmul(st["q_wmap_wodom"], lo["q_w_curr"]), atol=1e-14)
            np.testing.assert_allclose(t, qrot(st["q_wmap_wodom"], lo["t_w_curr"]) + st["t_wmap_wodom"], atol=1e-12)
        else:
            np.testing.assert_array_equal(np.r_[q, t], np.r_[st["q_w_curr"], st["t_w_curr"]])
    assert flags == [True, False, True, False]


def test_device_atan2f_is_the_c_librarys_atan2f(tmp_path, synth):
    """The CUDA scan registration computes azimuths with csrc/fdlibm_atan2f.h (the fdlibm algorithm glibc <= 2.40 ships) so
    that relTime, intensity and the 2 pi unwrapping decisions of scan_registration.cpp:166-262 carry the bits of the
    platform the reference runs on.  Host build of the same header vs the C library's atan2f: tens of millions of random
    arguments, the special values, and the (y, x) of a synthetic scan."""
    import ctypes as C
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "fdlibm_check.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(root, "tests", "native", "fdlibm_check.cpp"), "-o", so])
    L = C.CDLL(so)
    L.fd_mismatches.restype = C.c_long
    L.fd_mismatches.argtypes = [C.c_long, C.c_uint, C.c_float]
    L.fd_same.argtypes = [C.c_float, C.c_float]
    for seed, scale in ((1, 80.0), (2, 1.0), (3, 1e-3), (4, 1e6)):
        assert L.fd_mismatches(8_000_000, seed, scale) == 0, (seed, scale)
    inf, nan = float("inf"), float("nan")
    for y in (0.0, -0.0, 1.0, -1.0, inf, -inf, nan, 1e-40, 3e38, 0.4375, 0.6875, 1.1875, 2.4375, 2.0 ** 26, 2.0 ** -30):
        for x in (0.0, -0.0, 1.0, -1.0, inf, -inf, nan, 1e-40, -1e-40, 3e38, -3e38, 2.0 ** -70, 2.0 ** 70):
            assert L.fd_same(y, x), (y, x)
    sc = synth.ScanStream(3, n_cols=512).scan(0)
    sc = sc[np.isfinite(sc[:, 0])]
    L.fd_atan2f.restype = C.c_float
    L.fd_atan2f.argtypes = [C.c_float, C.c_float]
    for x, y in sc[::5, :2]:
        assert L.fd_same(float(y), float(x)), (y, x)
