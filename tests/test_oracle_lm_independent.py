"""A second, independent implementation (numpy, dense QR via lstsq on the augmented system, written from the published
description of Ceres 2.0's TrustRegionMinimizer + LevenbergMarquardtStrategy + Corrector) run on the very problem the oracle's
C++ restatement (oracle/ceres_lm.hpp, normal equations) solved: the per-iteration records — cost, candidate cost, model cost
change, relative decrease, trust-region radius, accept / reject — and the final parameters must agree.  This does not pin the
oracle to Ceres (nothing in this image can) but it removes transcription slips as a source of error: the
GPU solver (gn_solver.cuh) is compared with the same records in tests/test_gpu_lidar.py."""
import numpy as np


def _plus(x, d):
    nd = np.linalg.norm(d[:3])
    q = x[:4].copy()
    if nd > 0:
        s = np.sin(nd) / nd
        dq = np.r_[s * d[:3], np.cos(nd)]
        ax, ay, az, aw = x[:4]
        qx, qy, qz, qw = dq
        q = np.array([qw * ax + qx * aw + qy * az - qz * ay, qw * ay + qy * aw + qz * ax - qx * az,
                      qw * az + qz * aw + qx * ay - qy * ax, qw * aw - qx * ax - qy * ay - qz * az])
    return np.r_[q, x[4:] + d[3:]]


def _evaluate(oracle, blocks, x, a=0.1):
    """blocks: list of (kind, pts).  Returns cost, loss-corrected residual vector and local Jacobian (n x 6)."""
    xq, yq, zq, wq = x[:4]
    P = np.array([[wq, zq, -yq], [-zq, wq, xq], [yq, -xq, wq], [-xq, -yq, -zq]])
    rs, Js, cost = [], [], 0.0
    for kind, pts in blocks:
        r, J = oracle.factor_eval(kind, pts, x)
        Jl = np.c_[J[:, :4] @ P, J[:, 4:]]
        s = float(r @ r)
        if s > a * a:                                  # HuberLoss: rho = 2 a sqrt(s) - a^2, rho' = a / sqrt(s), rho'' < 0
            cost += 0.5 * (2 * a * np.sqrt(s) - a * a)
            w = np.sqrt(a / np.sqrt(s))                # Corrector with rho'' <= 0: scale by sqrt(rho')
        else:
            cost += 0.5 * s
            w = 1.0
        rs.append(w * r); Js.append(w * Jl)
    return cost, np.concatenate(rs), np.vstack(Js)


def _numpy_ceres_lm(oracle, blocks, x0, max_iterations=4):
    x = x0.copy()
    cost, r, J = _evaluate(oracle, blocks, x)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    radius, decrease = 1e4, 2.0
    records = [(cost, 0.0, 0.0, 0.0, radius, 0, 0)]
    diag, reuse, it, invalid_run, term = None, False, 0, 0, 0
    while True:
        if it >= max_iterations:
            term = 0; break
        if np.max(np.abs(J.T @ r)) <= 1e-10:
            term = 1; break
        if radius <= 1e-32:
            term = 2; break
        it += 1
        Js = J * scale
        if not reuse:
            diag = np.clip((Js * Js).sum(0), 1e-6, 1e32)
        D = np.sqrt(diag / radius)
        y = np.linalg.lstsq(np.vstack([Js, np.diag(D)]), np.r_[r, np.zeros(6)], rcond=None)[0]
        reuse = True
        step = -y
        Jstep = Js @ step
        mcc = -float(Jstep @ (r + 0.5 * Jstep))
        if not (np.all(np.isfinite(y)) and mcc > 0):
            invalid_run += 1
            radius /= decrease; decrease *= 2
            records.append((cost, 0.0, mcc, 0.0, radius, 0, 0))
            if invalid_run >= 5:
                term = 4; break
            continue
        invalid_run = 0
        cand = _plus(x, step * scale)
        cand_cost, rc, Jc = _evaluate(oracle, blocks, cand)
        if np.linalg.norm(cand - x) <= 1e-8 * (np.linalg.norm(x) + 1e-8):
            records.append((cost, cand_cost, mcc, 0.0, radius, 1, 0)); term = 2; break
        if abs(cost - cand_cost) <= 1e-6 * cost:
            records.append((cost, cand_cost, mcc, 0.0, radius, 1, 0)); term = 3; break
        rho = (cost - cand_cost) / mcc
        if rho > 1e-3:
            x, cost, r, J = cand, cand_cost, rc, Jc
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease, reuse = 2.0, False
            records.append((cost, cand_cost, mcc, rho, radius, 1, 1))
        else:
            radius /= decrease; decrease *= 2; reuse = True
            records.append((cost, cand_cost, mcc, rho, radius, 1, 0))
    return x, np.array(records), term


def test_independent_lm_reproduces_the_oracles_iteration_records(oracle, scans_small):
    scans, _ = scans_small
    lo = oracle.LaserOdometry()
    srs = [oracle.scan_registration(s) for s in scans[:3]]
    lo.solve(srs[0])
    lo.solve(srs[1])                                   # warms the motion prior
    st = lo.state
    x_start = np.r_[st["q_last_curr"], st["t_last_curr"]]
    lo.solve(srs[2])
    CL, SL = srs[1].cornerPointsLessSharp, srs[1].surfPointsLessFlat
    P, F = srs[2].cornerPointsSharp, srs[2].surfPointsFlat
    x = x_start
    for p, tr in enumerate(lo.trace()):
        blocks = []
        for i, a, b in tr["corner"]:
            blocks.append((0, np.r_[P[i, :3], CL[a, :3], CL[b, :3]].astype(np.float64)))
        for i, j, l, m in tr["plane"]:
            blocks.append((1, np.r_[F[i, :3], SL[j, :3], SL[l, :3], SL[m, :3]].astype(np.float64)))
        assert len(blocks) > 200
        x_np, rec, term = _numpy_ceres_lm(oracle, blocks, x, max_iterations=4)
        ref = tr["iterations"]
        assert rec.shape == ref.shape, (p, rec.shape, ref.shape)
        np.testing.assert_allclose(rec[:, 0], ref[:, 0], rtol=1e-9)                       # cost after each iteration
        np.testing.assert_allclose(rec[:, 1], ref[:, 1], rtol=1e-9, atol=1e-15)           # candidate cost
        np.testing.assert_allclose(rec[:, 2], ref[:, 2], rtol=1e-6, atol=1e-15)           # model cost change
        np.testing.assert_allclose(rec[:, 3], ref[:, 3], rtol=1e-6, atol=1e-12)           # relative decrease
        np.testing.assert_allclose(rec[:, 4], ref[:, 4], rtol=1e-6)                       # radius
        assert np.array_equal(rec[:, 5:7], ref[:, 5:7])                                   # valid / successful
        assert term == tr["termination"]
        np.testing.assert_allclose(x_np, tr["para"], atol=1e-9)
        x = tr["para"].copy()                          # the next pass starts where this one ended (detach_VO_LO = true)


def _numpy_ceres_lm_euclid(eval_fn, x0, max_iterations=100):
    """Same algorithm for plain R^6 parameters (visual odometry: angle-axis + t, no manifold)."""
    x = x0.copy()
    cost, r, J = eval_fn(x)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    radius, decrease = 1e4, 2.0
    records = [(cost, 0.0, 0.0, 0.0, radius, 0, 0)]
    diag, reuse, it, invalid_run, term = None, False, 0, 0, 0
    while True:
        if it >= max_iterations:
            term = 0; break
        if np.max(np.abs(J.T @ r)) <= 1e-10:
            term = 1; break
        if radius <= 1e-32:
            term = 2; break
        it += 1
        Js = J * scale
        if not reuse:
            diag = np.clip((Js * Js).sum(0), 1e-6, 1e32)
        y = np.linalg.lstsq(np.vstack([Js, np.diag(np.sqrt(diag / radius))]), np.r_[r, np.zeros(6)], rcond=None)[0]
        reuse = True
        Jstep = Js @ (-y)
        mcc = -float(Jstep @ (r + 0.5 * Jstep))
        if not (np.all(np.isfinite(y)) and mcc > 0):
            invalid_run += 1
            radius /= decrease; decrease *= 2
            records.append((cost, 0.0, mcc, 0.0, radius, 0, 0))
            if invalid_run >= 5:
                term = 4; break
            continue
        invalid_run = 0
        cand = x - y * scale
        cand_cost, rc, Jc = eval_fn(cand)
        if np.linalg.norm(cand - x) <= 1e-8 * (np.linalg.norm(x) + 1e-8):
            records.append((cost, cand_cost, mcc, 0.0, radius, 1, 0)); term = 2; break
        if abs(cost - cand_cost) <= 1e-6 * cost:
            records.append((cost, cand_cost, mcc, 0.0, radius, 1, 0)); term = 3; break
        rho = (cost - cand_cost) / mcc
        if rho > 1e-3:
            x, cost, r, J = cand, cand_cost, rc, Jc
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease, reuse = 2.0, False
            records.append((cost, cand_cost, mcc, rho, radius, 1, 1))
        else:
            radius /= decrease; decrease *= 2; reuse = True
            records.append((cost, cand_cost, mcc, rho, radius, 1, 0))
    return x, np.array(records), term


def test_independent_lm_reproduces_the_visual_odometry_solve(oracle, synth):
    calib = synth.kitti_like_calibration()
    s = synth.ScanStream(11, n_cols=1024)
    vo = oracle.VisualOdometry(*calib)
    for k in range(2):
        scan = s.scan(k)
        vo.reset(); vo.process_cloud(scan[np.isfinite(scan).all(1)])
    pu, cu, _ = synth.make_matches(s, 1, n_matches=500)
    res = vo.solve(pu, cu)
    types, obs = vo.residuals(pu.shape[0])
    blocks = [(int(t) - 1, o) for t, o in zip(types, obs) if t > 0]          # type 1 = CostFunctor32, 2 = CostFunctor22
    assert len(blocks) == res["counter32"] + res["counter22"] and res["counter32"] > 100

    def eval_fn(x, a=0.1):
        rs, Js, cost = [], [], 0.0
        for kind, o in blocks:
            r, J = oracle.vo_factor_eval(kind, o, x)
            sq = float(r @ r)
            if sq > a * a:
                cost += 0.5 * (2 * a * np.sqrt(sq) - a * a); w = np.sqrt(a / np.sqrt(sq))
            else:
                cost += 0.5 * sq; w = 1.0
            rs.append(w * r); Js.append(w * J)
        return cost, np.concatenate(rs), np.vstack(Js)

    x_np, rec, term = _numpy_ceres_lm_euclid(eval_fn, np.zeros(6), max_iterations=100)
    ref = vo.trace()
    assert rec.shape[0] == ref.shape[0], (rec.shape, ref.shape)
    np.testing.assert_allclose(rec[:, 0], ref[:, 0], rtol=1e-8)
    assert np.array_equal(rec[:, 5:7], ref[:, 5:7])
    np.testing.assert_allclose(rec[:, 4], ref[:, 4], rtol=1e-5)
    assert term == res["termination"]
    np.testing.assert_allclose(x_np[:3], res["angles_0to1"], atol=1e-8)
    np.testing.assert_allclose(x_np[3:], res["t_0to1"], atol=1e-8)
