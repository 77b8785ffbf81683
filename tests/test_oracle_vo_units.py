"""CPU cross-checks of the visual-odometry part of the oracle (oracle/visual_odometry.hpp) against literal numpy / Python
restatements of the reference lines (point_cloud_util.cpp:148-174 projectPointCloud, :205-260 downsamplePointCloud with its
count-before-increment running "mean" (SURVEY Q6), :302-407 queryDepth), plus the VO -> LO prior conversion
(vloam_tf.cpp:59-63) against a matrix restatement, and the solve against ground-truth motion."""
import numpy as np
import pytest


def _project_numpy(xyz, cam_T_velo, rect0_T_cam, P_rect0):
    X = np.c_[xyz.astype(np.float32), np.ones(len(xyz), np.float32)]                          # visual_odometry.cpp:163-170
    # Eigen evaluates the chain left to right in float: ((X * T') * R') * P'
    p = ((X @ cam_T_velo.T.astype(np.float32)).astype(np.float32) @ rect0_T_cam.T.astype(np.float32)).astype(np.float32) \
        @ P_rect0.T.astype(np.float32)
    p = p.astype(np.float32)
    front = p[p[:, 2] > np.float32(0.1)]
    out = front.copy()
    out[:, 0] = front[:, 0] * (np.float32(1.0) / front[:, 2])
    out[:, 1] = front[:, 1] * (np.float32(1.0) / front[:, 2])
    return out


def _downsample_python(p2d, W=249, H=75, g=5):
    bx = np.zeros((W, H), np.float32); by = np.zeros((W, H), np.float32); bd = np.zeros((W, H), np.float32)
    bc = np.zeros((W, H), np.int32)
    f = np.float32
    for x, y, d in p2d:
        ix, iy = int(f(x) / f(g)), int(f(y) / f(g))                      # static_cast<int>: truncation toward zero
        if 0 <= ix < W and 0 <= iy < H:
            if bc[ix, iy] == 0:
                bx[ix, iy], by[ix, iy], bd[ix, iy] = x, y, d
            else:                                                        # divisor = hits BEFORE this one
                c = f(bc[ix, iy])
                bx[ix, iy] = f(bx[ix, iy] + f(f(x - bx[ix, iy]) / c))
                by[ix, iy] = f(by[ix, iy] + f(f(y - by[ix, iy]) / c))
                bd[ix, iy] = f(bd[ix, iy] + f(f(d - bd[ix, iy]) / c))
            bc[ix, iy] += 1
    return bx, by, bd, bc


def _query_depth_python(bx, by, bd, bc, x, y, radius=2, g=5):
    f = np.float32
    ix, iy = int(f(x) / f(g)), int(f(y) / f(g))
    nb = []
    for i in range(ix - radius, ix + radius + 1):
        for j in range(iy - radius, iy + radius + 1):
            if 0 <= i < bx.shape[0] and 0 <= j < bx.shape[1] and bc[i, j] > 0:
                dist = f(np.sqrt(f(f(f(x) - bx[i, j]) ** 2 + f(f(y) - by[i, j]) ** 2)))
                nb.append((dist, bd[i, j]))
    if len(nb) < 10:
        return -1.0
    nb.sort(key=lambda t: t[0])
    (d0, z0), (d1, z1), (d2, z2) = nb[:3]
    return float(f(f(z0 * d1 * d2 + z1 * d0 * d2 + z2 * d0 * d1) / f(f(0.0001) + d1 * d2 + d0 * d2 + d0 * d1)))


@pytest.fixture(scope="module")
def vo_setup(synth, oracle):
    calib = synth.kitti_like_calibration()
    s = synth.ScanStream(11, n_cols=1024)
    return calib, s


def test_projection_and_buckets_match_literal_restatement(vo_setup, oracle):
    calib, s = vo_setup
    vo = oracle.VisualOdometry(*calib)
    scan = s.scan(0)
    scan = scan[np.isfinite(scan).all(1)]
    vo.reset()
    vo.process_cloud(scan)
    slot = vo.slot
    got = vo.projected(slot)
    ref = _project_numpy(scan, *[np.asarray(c, np.float32) for c in calib])
    assert got.shape == ref.shape
    # float matrix chains: the oracle follows Eigen's evaluation order; numpy's BLAS may fuse differently -> few ulp
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-3)
    bx, by, bd, bc = vo.buckets(slot)
    rx, ry, rd, rc = _downsample_python(got)                # feed the oracle's own projection: the fold must then be bit-exact
    assert np.array_equal(bc, rc)
    for a, b in ((bx, rx), (by, ry), (bd, rd)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert bc.max() >= 3 and (bc > 0).sum() > 500            # the order-dependent branch is exercised


def test_query_depth_matches_literal_restatement(vo_setup, oracle):
    calib, s = vo_setup
    vo = oracle.VisualOdometry(*calib)
    scan = s.scan(1)
    scan = scan[np.isfinite(scan).all(1)]
    vo.reset(); vo.process_cloud(scan)
    slot = vo.slot
    bx, by, bd, bc = vo.buckets(slot)
    rng = np.random.default_rng(0)
    hits = 0
    for _ in range(400):
        x, y = float(int(rng.uniform(0, 1241))), float(int(rng.uniform(100, 375)))     # integer pixels (Q7)
        z = vo.query_depth(slot, x, y)
        zr = _query_depth_python(bx, by, bd, bc, x, y)
        if zr < 0:
            assert z == -1.0
        else:
            hits += 1
            assert abs(z - zr) <= 2e-6 * max(1.0, abs(zr)), (x, y, z, zr)
    assert hits > 100


def test_vo_to_lo_prior_matches_matrix_chain(oracle):
    """velo_last_VOT_velo_curr = velo_T_cam0 * cam0_curr_T_cam0_last^-1 * velo_T_cam0^-1 (vloam_tf.cpp:59-63)."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(4)
    velo_T_cam0 = np.eye(4)
    velo_T_cam0[:3, :3] = R.from_rotvec(rng.normal(0, 0.8, 3)).as_matrix()
    velo_T_cam0[:3, 3] = rng.normal(0, 0.5, 3)
    for _ in range(20):
        aa, t = rng.normal(0, 0.05, 3), rng.normal(0, 1.0, 3)
        T = np.eye(4); T[:3, :3] = R.from_rotvec(aa).as_matrix(); T[:3, 3] = t
        ref = velo_T_cam0 @ np.linalg.inv(T) @ np.linalg.inv(velo_T_cam0)
        got = oracle.vo_to_lo_prior(aa, t, velo_T_cam0)
        q, tt = got[:4], got[4:]
        np.testing.assert_allclose(R.from_quat(q).as_matrix(), ref[:3, :3], atol=1e-12)
        np.testing.assert_allclose(tt, ref[:3, 3], atol=1e-12)


def test_vo_solve_recovers_generator_motion(vo_setup, synth, oracle):
    calib, s = vo_setup
    vo = oracle.VisualOdometry(*calib)
    errs = []
    for k in range(3):
        scan = s.scan(k)
        vo.reset(); vo.process_cloud(scan[np.isfinite(scan).all(1)])
        if k == 0:
            continue
        pu, cu, (Rc, tc) = synth.make_matches(s, k, n_matches=600, outlier_frac=0.0, pixel_sigma=0.1)
        r = vo.solve(pu, cu)
        assert r["counter32"] > 100
        # the estimate is the camera motion the generator applied (about 1 m forward along z per frame)
        errs.append(abs(np.linalg.norm(r["t_0to1"]) - np.linalg.norm(tc)))
    assert errs and max(errs) < 0.3, errs
