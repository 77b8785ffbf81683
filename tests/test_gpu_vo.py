"""GPU parity tests for the visual-odometry part vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("rect33", [1.0, 0.0])   # 0.0 = the ROS path's rect0_T_cam(3,3)
def test_vo_depth_buckets_query_and_solve(synth, oracle, rect33):
    import vloam_b200 as V
    stream = synth.ScanStream(51, n_cols=1024)
    cam_T_velo, rect0_T_cam, P = synth.kitti_like_calibration()
    rect0_T_cam = rect0_T_cam.copy(); rect0_T_cam[3, 3] = rect33
    vo = V.VisualOdometry(batch=1, max_points=64 * 1024, max_matches=1024)
    vo.setUpPointCloud(cam_T_velo, rect0_T_cam, P)
    ovo = oracle.VisualOdometry(cam_T_velo, rect0_T_cam, P)
    rng = np.random.default_rng(5)
    for k in range(3):
        cloud = stream.scan(k)
        vo.reset(); ovo.reset()
        vo.processPointCloud(cloud); ovo.process_cloud(cloud)
        slot = ovo.slot
        for g, o, name in zip(vo.buckets(0), ovo.buckets(slot), ("x", "y", "depth", "count")):
            if name == "count":
                assert np.array_equal(g, o)
            else:
                assert np.array_equal(_bits(g), _bits(o)), f"frame {k}: bucket {name}"
        xy = np.c_[rng.uniform(0, 1242, 400), rng.uniform(0, 375, 400)].astype(np.float32)
        gd = vo.queryDepth(xy, slot=0)
        od = np.array([ovo.query_depth(slot, x, y) for x, y in xy], np.float32)
        assert np.array_equal(_bits(gd), _bits(od)), f"frame {k}: queryDepth"
        assert (gd > 0).sum() > 50
        if k == 0:
            continue
        prev_uv, curr_uv, (Rc, tc) = synth.make_matches(stream, k, n_matches=800)
        gs = vo.solveNlsAll(prev_uv, curr_uv)
        os_ = ovo.solve(prev_uv, curr_uv)
        assert (gs["counter32"][0], gs["counter22"][0]) == (os_["counter32"], os_["counter22"])
        assert gs["counter32"][0] > 100
        gt_, go_ = vo.residuals()
        ot_, oo_ = ovo.residuals(prev_uv.shape[0])
        m_ = prev_uv.shape[0]
        assert np.array_equal(gt_[:m_], ot_), "residual types differ"
        bad = np.nonzero(np.abs(go_[:m_] - oo_).max(axis=1) > 1e-12)[0]
        assert bad.size == 0, (bad[:5], go_[bad[:3]], oo_[bad[:3]])
        tg, to = vo.trace(), ovo.trace()
        np.testing.assert_allclose(tg["iterations"][0, 0], to[0, 0], rtol=1e-10)      # initial cost
        assert tg["n_records"] == to.shape[0]
        nk = min(8, to.shape[0])                                                      # the library keeps the first 8 records
        np.testing.assert_allclose(tg["iterations"][:nk, 0], to[:nk, 0], rtol=1e-8)   # cost after every kept iteration
        np.testing.assert_array_equal(tg["iterations"][:nk, 5:7], to[:nk, 5:7])       # step valid / successful
        np.testing.assert_allclose(tg["iterations"][:nk, 4], to[:nk, 4], rtol=1e-6)   # trust-region radius
        np.testing.assert_allclose(gs["angles_0to1"][0], os_["angles_0to1"], atol=1e-6)
        np.testing.assert_allclose(gs["t_0to1"][0], os_["t_0to1"], atol=1e-5)
        # and the estimate is a sensible camera motion (about 1 m forward along z)
        assert abs(np.linalg.norm(gs["t_0to1"][0]) - np.linalg.norm(tc)) < 0.3
        # the LO prior as initial value (reset_VO_to_identity == false)
        init = np.r_[0.0, 0.0, 0.0, tc]
        g2 = vo.solveNlsAll(prev_uv, curr_uv, init=init[None])
        o2 = ovo.solve(prev_uv, curr_uv, init_aa=init[:3], init_t=init[3:])
        np.testing.assert_allclose(g2["t_0to1"][0], o2["t_0to1"], atol=1e-5)
    vo.close()


def test_vo_batched_and_edge_cases(synth, oracle):
    import vloam_b200 as V
    cam_T_velo, rect0_T_cam, P = synth.kitti_like_calibration()
    streams = [synth.ScanStream(61 + i, n_cols=512) for i in range(2)]
    vo = V.VisualOdometry(batch=2, max_points=64 * 512, max_matches=256)
    vo.setUpPointCloud(cam_T_velo, rect0_T_cam, P)
    with pytest.raises(V.VloamError):
        vo.processPointCloud(np.zeros((2, 16, 3), np.float32))   # before reset()
    ovos = [oracle.VisualOdometry(cam_T_velo, rect0_T_cam, P) for _ in streams]
    for k in range(2):
        vo.reset()
        clouds = np.stack([s.scan(k) for s in streams])
        vo.processPointCloud(clouds)
        for b, o in enumerate(ovos):
            o.reset(); o.process_cloud(clouds[b])
            assert np.array_equal(vo.buckets(0, b)[3], o.buckets(o.slot)[3])
    m = [synth.make_matches(s, 1, n_matches=200) for s in streams]
    prev = np.stack([x[0][:200] for x in m]); curr = np.stack([x[1][:200] for x in m])
    gs = vo.solveNlsAll(prev, curr)
    for b, o in enumerate(ovos):
        r = o.solve(prev[b], curr[b])
        np.testing.assert_allclose(gs["t_0to1"][b], r["t_0to1"], atol=1e-5)
    # no matches at all: the solve returns the initial value
    g0 = vo.solveNlsAll(np.zeros((2, 0, 2), np.float32), np.zeros((2, 0, 2), np.float32))
    assert np.all(g0["t_0to1"] == 0) and np.all(g0["counter32"] == 0)
    vo.close()


def test_vo_prior_feeds_laser_odometry_on_device(synth, oracle):
    """The coupled mode of vloam_main (detach_VO_LO = false, vloam_main_node.cpp:150-167): VO result -> VloamTF::VO2VeloAndBase
    -> LaserOdometry prior, with the whole chain resident on the device, against the same chain through the oracle."""
    import torch
    import vloam_b200 as V
    n_cols = 512
    stream = synth.ScanStream(77, n_cols=n_cols)
    cam_T_velo, rect0_T_cam, P = synth.kitti_like_calibration()
    velo_T_cam0 = np.linalg.inv(cam_T_velo.astype(np.float64))
    ctx = V.Context()
    vo = V.VisualOdometry(ctx, batch=1, max_points=64 * n_cols, max_matches=1024)
    vo.setUpPointCloud(cam_T_velo, rect0_T_cam, P)
    lom = V.LidarOdometryMapping(ctx, batch=1, max_points=64 * n_cols, detach_VO_LO=0)
    ovo = oracle.VisualOdometry(cam_T_velo, rect0_T_cam, P)
    olo = oracle.LaserOdometry(detach_VO_LO=False)
    dev = torch.device("cuda", 0)
    prior_dev = torch.zeros((1, 7), dtype=torch.float64, device=dev)
    prior_dev[0, 3] = 1.0
    for k in range(4):
        cloud = stream.scan(k)
        cloud_dev = torch.from_numpy(np.ascontiguousarray(cloud[None])).to(dev)
        n_dev = torch.tensor([cloud.shape[0]], dtype=torch.int32, device=dev)
        vo.reset(); ovo.reset()
        vo.processPointCloudDevice(cloud_dev, n_dev, 3, cloud.shape[0]); ovo.process_cloud(cloud)
        o_prior = np.r_[0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
        if k > 0:
            prev_uv, curr_uv, _ = synth.make_matches(stream, k, n_matches=600)
            m = prev_uv.shape[0]
            pu = np.zeros((1, 1024, 2), np.float32); cu = np.zeros((1, 1024, 2), np.float32)
            pu[0, :m] = prev_uv; cu[0, :m] = curr_uv
            vo.solveNlsAllDevice(torch.from_numpy(pu).to(dev), torch.from_numpy(cu).to(dev),
                                 torch.tensor([m], dtype=torch.int32, device=dev))
            vo.exportLOPrior(velo_T_cam0, prior_dev)
            os_ = ovo.solve(prev_uv, curr_uv)
            o_prior = oracle.vo_to_lo_prior(os_["angles_0to1"], os_["t_0to1"], velo_T_cam0)
            ctx.synchronize()
            g_prior = prior_dev.cpu().numpy()[0]
            np.testing.assert_allclose(g_prior[4:], o_prior[4:], atol=1e-5)
            np.testing.assert_allclose(g_prior[:4], o_prior[:4], atol=1e-6)
            # given the same VO result the conversion itself agrees to rounding
            gr = vo.result()
            o_same = oracle.vo_to_lo_prior(gr["angles_0to1"][0], gr["t_0to1"][0], velo_T_cam0)
            np.testing.assert_allclose(g_prior, o_same, atol=1e-14)
        lom.reset()
        lom.scanRegistrationDevice(cloud_dev, n_dev, 3, cloud.shape[0])
        lom.laserOdometryIO(prior=prior_dev, fetch=False)
        pose = lom.lo_pose()
        ref = oracle.scan_registration(cloud)
        olo.solve(ref, prior_q=o_prior[:4], prior_t=o_prior[4:])
        st = olo.state
        if k > 0:
            # the two arms start LO from VO estimates that differ by the VO solve tolerance (1e-6 rad / 1e-5 m above)
            np.testing.assert_allclose(pose["t_last_curr"][0], st["t_last_curr"], atol=1e-4)
            np.testing.assert_allclose(np.abs(pose["q_last_curr"][0]), np.abs(st["q_last_curr"]), atol=1e-4)
            assert pose["corner_correspondence"][0] > 50 and pose["plane_correspondence"][0] > 200
    vo.close(); lom.close(); ctx.close()
