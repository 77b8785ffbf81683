"""Descriptor matching of the visual-odometry front end — the one row whose parity is PINNED by
the reference's own dependency: tests/golden/vo_frontend_cv2.npz was produced by OpenCV (cv2) with the calls of
image_util.cpp:13-26, 176-204, 228-283 (tests/golden/make_golden_vo_frontend.py)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vo_frontend_cv2.npz")
PAIRS = (0, 1, 2)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_numpy_restatement_matches_opencv(golden):
    from oracle import vo_frontend as F
    for p in PAIRS:
        d0, d1 = golden[f"pair{p}_desc0"], golden[f"pair{p}_desc1"]
        idx, dist = F.knn2(d0, d1)
        assert np.array_equal(idx, golden[f"pair{p}_knn_idx"][:, 1:]), f"pair {p}: 2-NN train indices (ties included)"
        assert np.array_equal(golden[f"pair{p}_knn_idx"][:, 0], np.arange(d0.shape[0]))
        assert np.array_equal(dist.astype(np.float32), golden[f"pair{p}_knn_dist"]), f"pair {p}: distances"
        assert np.array_equal(F.match_descriptors(d0, d1), golden[f"pair{p}_matches"]), f"pair {p}: ratio-tested matches"
    assert (golden["pair2_knn_dist"][:, 0] == golden["pair2_knn_dist"][:, 1]).mean() > 0.9      # the tie fixture really ties


def test_opencv_still_reproduces_the_golden_vectors(golden):
    """If cv2 is importable, the committed vectors are what it produces today (guards against a stale fixture)."""
    cv2 = pytest.importorskip("cv2")
    for p in PAIRS:
        d0, d1 = golden[f"pair{p}_desc0"], golden[f"pair{p}_desc1"]
        knn = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False).knnMatch(d0, d1, 2)
        got = np.array([[m[0].queryIdx, m[0].trainIdx, m[1].trainIdx] for m in knn], np.int32)
        assert np.array_equal(got, golden[f"pair{p}_knn_idx"])
        good = np.array([[m[0].queryIdx, m[0].trainIdx, int(m[0].distance)] for m in knn if m[0].distance < 0.8 * m[1].distance], np.int32).reshape(-1, 3)
        assert np.array_equal(good, golden[f"pair{p}_matches"])


DETECT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vo_detect_cv2.npz")
IMAGES = ("kitti", "small", "shapes")


@pytest.fixture(scope="module")
def detect_golden():
    return np.load(DETECT)


def _ulp_distance(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def test_numpy_detection_matches_opencv(detect_golden):
    """Shi-Tomasi detection (image_util.cpp:11-37) against cv2's own outputs: the response map bit for bit on > 99.99 % of the
    pixels and within one unit in the last place elsewhere (module docstring of oracle/vo_frontend.py), the selected corners —
    coordinates AND order — exactly; and the selection stage alone exactly, fed with cv2's response map."""
    from oracle import vo_frontend as F
    g = detect_golden
    for name in IMAGES:
        img = g[f"{name}_image"]
        resp = F.min_eigen_response(img)
        ref = g[f"{name}_response"] if name == "small" else g[f"{name}_response_rows"]
        got = resp if name == "small" else resp[::8]
        assert got.shape == ref.shape
        ulp = _ulp_distance(got, ref)
        assert (ulp == 0).mean() > 0.9999, (name, (ulp == 0).mean())
        big = np.abs(ref) > 1e-3 * float(g[f"{name}_response_max"])      # (tiny responses are differences of nearly equal sums)
        assert ulp[big].max() <= 1, (name, ulp[big].max())
        assert np.float32(resp.max()) == g[f"{name}_response_max"]
        corners = F.select_corners(resp)
        assert np.array_equal(corners, g[f"{name}_corners"]), name
    assert np.array_equal(F.select_corners(g["small_response"]), g["small_corners"])
    assert len(g["shapes_corners"]) < 1024          # the quality threshold and the spacing decide here, not the cap


def test_corner_selection_against_a_literal_greedy_pass():
    """select_corners (grid of cvRound(min_distance) cells, as OpenCV) against the definition it implements: walk the local maxima
    by (response descending, address descending) and keep one when no kept corner is closer than min_distance — on response
    maps with many exactly equal values (the tie-break matters) and for spacings on both sides of the cell size."""
    from oracle import vo_frontend as F
    rng = np.random.default_rng(11)
    for min_distance, quant in ((7.5, 12), (3.0, 5), (12.4, 0)):
        eig = rng.random((90, 140)).astype(np.float32)
        if quant:
            eig = (np.floor(eig * quant) / quant).astype(np.float32)          # few distinct levels: plateaus and ties
        got = F.select_corners(eig, max_corners=300, quality=0.2, min_distance=min_distance)
        H, W = eig.shape
        thr = np.float32(np.float64(eig.max()) * 0.2)
        e = np.where(eig > thr, eig, np.float32(0))
        cand = []
        for y in range(1, H - 1):
            for x in range(1, W - 1):
                v = e[y, x]
                if v != 0 and v == e[y - 1:y + 2, x - 1:x + 2].max():
                    cand.append((-float(v), -(y * W + x), x, y))
        cand.sort()
        cell = int(np.rint(min_distance))
        kept = []
        for _, _, x, y in cand:
            # OpenCV only looks into the 3 x 3 cells around the candidate: for min_distance <= cell that is every corner in range
            near = [(px, py) for px, py in kept if abs(px // cell - x // cell) <= 1 and abs(py // cell - y // cell) <= 1]
            if all((x - px) ** 2 + (y - py) ** 2 >= min_distance * min_distance for px, py in near):
                kept.append((x, y))
                if len(kept) == 300:
                    break
        assert np.array_equal(got, np.array(kept, np.float32).reshape(-1, 2)), min_distance
        assert len(kept) > 20


def test_response_restatement_against_opencv_across_image_sizes():
    """If cv2 is importable: the restated operation order (fused Sobel taps, the scalar tail of width mod 32 columns, double box
    sums) holds for narrow, odd and large images, not only for the committed ones."""
    cv2 = pytest.importorskip("cv2")
    from oracle import vo_frontend as F
    rng = np.random.default_rng(9)
    for H, W in ((40, 20), (33, 50), (64, 64), (100, 95), (37, 129), (480, 752)):
        img = cv2.GaussianBlur((rng.random((H, W)) * 255).astype(np.uint8), (0, 0), 1.5)
        ulp = _ulp_distance(F.min_eigen_response(img), cv2.cornerMinEigenVal(img, 5, ksize=3))
        assert (ulp == 0).mean() > 0.9999 and ulp.max() <= 4, ((H, W), (ulp == 0).mean(), ulp.max())


def test_opencv_still_reproduces_the_detection_vectors(detect_golden):
    cv2 = pytest.importorskip("cv2")
    g = detect_golden
    for name in IMAGES:
        c = cv2.goodFeaturesToTrack(g[f"{name}_image"], 1024, 0.03, 7.5, None, blockSize=5, useHarrisDetector=False, k=0.04).reshape(-1, 2)
        assert np.array_equal(c, g[f"{name}_corners"]), name


@pytest.mark.gpu
def test_cuda_matcher_hits_the_opencv_vectors(golden):
    """vo_bf_match (TMA-staged train descriptors, popcount 2-NN, ratio test, ordered compaction) == cv2, bit for bit: the raw
    2-NN table with its tie-breaks, the accepted matches, and the matched pixel pairs left on the device for solveNlsAll.
    All three pairs as one batch of three streams."""
    import vloam_b200 as V
    vo = V.VisualOdometry(batch=3, max_points=1024, max_matches=1024)
    d0 = [golden[f"pair{p}_desc0"] for p in PAIRS]
    d1 = [golden[f"pair{p}_desc1"] for p in PAIRS]
    k0 = [golden[f"pair{p}_kp0"] for p in PAIRS]
    k1 = [golden[f"pair{p}_kp1"] for p in PAIRS]
    res = vo.matchDescriptors(d0, d1, k0, k1, ratio=0.8)
    qa, ta, na = vo.match_buffers()
    for p in PAIRS:
        g_idx, g_dist = golden[f"pair{p}_knn_idx"], golden[f"pair{p}_knn_dist"]
        assert np.array_equal(res[p]["knn"][:, :2], g_idx[:, 1:]), f"pair {p}: 2-NN indices"
        assert np.array_equal(res[p]["knn"][:, 2:].astype(np.float32), g_dist), f"pair {p}: 2-NN distances"
        assert np.array_equal(res[p]["matches"], golden[f"pair{p}_matches"]), f"pair {p}: matches"
    # the device-resident pixel pairs feed solveNlsAll without a host round trip
    nm = np.array([len(golden[f"pair{p}_matches"]) for p in PAIRS])
    assert nm[0] > 100
    assert qa and ta and na
    uq, ut = vo.match_uv()
    for buf, kps, col in ((uq, k0, 0), (ut, k1, 1)):
        for p in PAIRS:
            m = golden[f"pair{p}_matches"]
            assert np.array_equal(buf[p, : len(m)], kps[p][m[:, col]]), f"pair {p}: matched pixels ({'query' if col == 0 else 'train'})"
    # degenerate sizes: an empty query set, a train set with a single descriptor (no second neighbour -> no match)
    one = vo.matchDescriptors([d0[0][:0], d0[1], d0[2][:5]], [d1[0], d1[1][:1], d1[2][:2]])
    assert one[0]["matches"].shape == (0, 3) and one[1]["matches"].shape == (0, 3)
    assert np.array_equal(one[1]["knn"][:, 1], np.full(len(d0[1]), -1))
    from oracle import vo_frontend as F
    assert np.array_equal(one[2]["matches"], F.match_descriptors(d0[2][:5], d1[2][:2]))
    vo.close()


@pytest.mark.gpu
def test_cuda_detector_hits_the_opencv_vectors(detect_golden):
    """vo_min_eigen / vo_corner_candidates / vo_select_corners == the numpy restatement bit for bit (response map included), and
    == cv2.goodFeaturesToTrack on the committed images: the same corners in the same order."""
    import vloam_b200 as V
    from oracle import vo_frontend as F
    g = detect_golden
    for name in IMAGES:
        img = g[f"{name}_image"]
        vo = V.VisualOdometry(batch=2, max_points=1024, max_matches=1024)
        flipped = np.ascontiguousarray(img[::-1, ::-1])                # second stream: another image of the same size
        got = vo.detKeypoints(np.stack([img, flipped]))
        for b, im in enumerate((img, flipped)):
            resp = F.min_eigen_response(im)
            assert np.array_equal(vo.corner_response(b).view(np.uint32), resp.view(np.uint32)), (name, b, "response map")
            assert np.array_equal(got[b], F.select_corners(resp)), (name, b, "corners")
        assert np.array_equal(got[0], g[f"{name}_corners"]), (name, "cv2 corners")
        # other parameters: fewer corners, wider spacing (the greedy pass decides more), a flat image (no corner at all)
        few = vo.detKeypoints(np.stack([img, np.full_like(img, 77)]), max_corners=100, quality_level=0.1, min_distance=20.0)
        assert np.array_equal(few[0], F.good_features_to_track(img, 100, 0.1, 20.0))
        assert few[1].shape == (0, 2)
        if name == "kitti":
            # ~13 000 candidates: only the strongest ~4 100 are ranked at first.  Spacing 12: the 1024th corner sits inside that
            # prefix; spacing 20: the frame holds fewer than 1024 such corners, the prefix runs out and the whole list is ranked
            for md in (12.0, 20.0):
                wide = vo.detKeypoints(pair2 := np.stack([img, flipped]), max_corners=1024, quality_level=0.03, min_distance=md)
                for b in range(2):
                    assert np.array_equal(wide[b], F.good_features_to_track(pair2[b], 1024, 0.03, md)), (md, b)
            assert len(wide[0]) < 1024
        vo.close()


# ------------------------------------------------------------------------------------------------ ORB description, processImage
ORB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vo_orb_cv2.npz")
VIEWS = ("kitti", "small", "shapes", "kitti_w1087", "kitti_next")


def _views(g):
    """The images of the ORB fixture, rebuilt from the committed detection images (tests/golden/make_golden_vo_orb.py views())."""
    k = g["kitti_image"]
    nxt = np.roll(k, (-2, 5), axis=(0, 1)).astype(np.int32)
    yy, xx = np.mgrid[0:k.shape[0], 0:k.shape[1]]
    nxt = np.clip(nxt + ((xx * 7 + yy * 13) % 5) - 2, 0, 255).astype(np.uint8)
    return {"kitti": k, "small": g["small_image"], "shapes": g["shapes_image"], "kitti_w1087": np.ascontiguousarray(k[:300, 40:1127]),
            "kitti_next": nxt}


@pytest.fixture(scope="module")
def orb_golden():
    return np.load(ORB)


def test_numpy_orb_description_matches_opencv(detect_golden, orb_golden):
    """ImageUtil::descKeypoints (image_util.cpp:162-212) against cv2.ORB's own outputs: which key points survive ORB's border
    filter (indices, order) and every descriptor byte — on the detector's corners and on key points chosen to sit on the filter's
    band, at exact halves (cvRound) and outside the image; the detection + description + matching chain of processImage too."""
    from oracle import vo_frontend as F
    g = orb_golden
    feats = {}
    for name, img in _views(detect_golden).items():
        corners = F.good_features_to_track(img)
        assert np.array_equal(corners, g[f"{name}_corners"]), (name, "corners")
        kept, desc = F.orb_describe(img, corners)
        assert np.array_equal(kept, g[f"{name}_kept_index"]) and np.array_equal(desc, g[f"{name}_desc"]), name
        feats[name] = desc
        kept, desc = F.orb_describe(img, g[f"{name}_extra"])
        assert np.array_equal(kept, g[f"{name}_extra_kept_index"]) and np.array_equal(desc, g[f"{name}_extra_desc"]), (name, "extra")
        assert 0 < len(kept) < len(g[f"{name}_extra"])
    assert np.array_equal(F.match_descriptors(feats["kitti"], feats["kitti_next"]), g["chain_matches"])
    assert len(g["chain_matches"]) > 500
    # degenerate inputs: no key point, every key point filtered, an image too small to hold any
    img = detect_golden["small_image"]
    for kp in (np.zeros((0, 2), np.float32), np.array([[3.0, 3.0], [630.0, 100.0]], np.float32)):
        kept, desc = F.orb_describe(img, kp)
        assert kept.shape == (0,) and desc.shape == (0, 32)
    assert F.orb_describe(img[:62, :300], np.array([[31.0, 31.0]], np.float32))[0].shape == (0,)
    assert F.orb_describe(img[:63, :300], np.array([[31.0, 31.0]], np.float32))[0].shape == (1,)


def test_orb_pattern_is_steering_invariant_at_minus_one_degree():
    """The key points carry cv::KeyPoint's default angle -1: the rotated pattern rounds back to the table (so the table is what
    the probing script recovers, and what the kernel indexes without rotating)."""
    from oracle import vo_frontend as F
    from oracle.orb_pattern import ORB_PATTERN
    x, y = F.orb_steered_pattern(-1.0)
    p = np.array(ORB_PATTERN, np.int32).reshape(-1, 2)
    assert np.array_equal(x, p[:, 0]) and np.array_equal(y, p[:, 1])
    assert np.abs(p).max() == 13                      # the reach vo_orb_describe's 33 x 33 patch is sized for
    x5, y5 = F.orb_steered_pattern(5.0)
    assert not (np.array_equal(x5, p[:, 0]) and np.array_equal(y5, p[:, 1]))


def test_orb_blur_and_description_against_opencv_across_image_sizes():
    """The float rounding sequence of the blur (fused vector body, unfused scalar tails; oracle/vo_frontend.py gaussian_blur_7x7)
    against cv2.sepFilter2D with cv2's own taps, and the whole description against cv2.ORB, on random images whose widths put
    sampled pixels into the row pass's scalar tail."""
    cv2 = pytest.importorskip("cv2")
    from oracle import vo_frontend as F
    taps = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    assert np.array_equal(taps, F._gauss7_taps())
    rng = np.random.default_rng(5)
    orb = cv2.ORB_create()
    for t in range(10):
        h, w = int(rng.integers(70, 420)), (32 * int(rng.integers(3, 30)) + 31 if t < 4 else int(rng.integers(70, 1300)))
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        if t % 2:
            img = cv2.GaussianBlur(img, (0, 0), 2.0)             # smooth content: many near-ties in the 256 comparisons
        assert np.array_equal(F.gaussian_blur_7x7(img), cv2.sepFilter2D(img, cv2.CV_8U, taps, taps, borderType=cv2.BORDER_REFLECT_101)), (h, w)
        kp = np.stack([rng.random(1500) * (w + 4) - 2, rng.random(1500) * (h + 4) - 2], 1).astype(np.float32)
        k2, d = orb.compute(img, [cv2.KeyPoint(float(x), float(y), 5.0) for x, y in kp])
        kept, desc = F.orb_describe(img, kp)
        assert np.array_equal(np.array([k.pt for k in k2], np.float32).reshape(-1, 2), kp[kept]), (h, w)
        assert np.array_equal(d, desc), (h, w)


def test_opencv_still_reproduces_the_orb_vectors(detect_golden, orb_golden):
    cv2 = pytest.importorskip("cv2")
    g = orb_golden
    for name, img in _views(detect_golden).items():
        for key in ("corners", "extra"):
            pts = g[f"{name}_{key}"]
            kept, d = cv2.ORB_create().compute(img, [cv2.KeyPoint(float(x), float(y), 5.0) for x, y in pts])
            idx = g[f"{name}_kept_index" if key == "corners" else f"{name}_extra_kept_index"]
            assert np.array_equal(np.array([k.pt for k in kept], np.float32).reshape(-1, 2), pts[idx]), (name, key)
            assert np.array_equal(d, g[f"{name}_desc" if key == "corners" else f"{name}_extra_desc"]), (name, key)


@pytest.mark.gpu
def test_cuda_orb_description_hits_the_opencv_vectors(detect_golden, orb_golden):
    """vo_orb_describe == cv2.ORB on the committed vectors (surviving key points, their order, every descriptor byte), on the
    detector's corners taken from the device and on host key points; second stream: the flipped image against the restatement."""
    import vloam_b200 as V
    from oracle import vo_frontend as F
    g = orb_golden
    for name, img in _views(detect_golden).items():
        vo = V.VisualOdometry(batch=2, max_points=1024, max_matches=1024)
        flipped = np.ascontiguousarray(img[::-1, ::-1])
        pair = np.stack([img, flipped])
        corners = vo.detKeypoints(pair)
        assert np.array_equal(corners[0], g[f"{name}_corners"]), name
        res = vo.descKeypoints()                                     # corners and images still on the device
        assert np.array_equal(res[0]["index"], g[f"{name}_kept_index"]), name
        assert np.array_equal(res[0]["descriptors"], g[f"{name}_desc"]), name
        assert np.array_equal(res[0]["keypoints"], corners[0][g[f"{name}_kept_index"]]), name
        kept, desc = F.orb_describe(flipped, corners[1])
        assert np.array_equal(res[1]["index"], kept) and np.array_equal(res[1]["descriptors"], desc), (name, "flipped")
        # host key points and host images
        ex = g[f"{name}_extra"]
        res = vo.descKeypoints([ex, ex[::-1]], pair)
        assert np.array_equal(res[0]["index"], g[f"{name}_extra_kept_index"]) and np.array_equal(res[0]["descriptors"], g[f"{name}_extra_desc"]), (name, "extra")
        kept, desc = F.orb_describe(flipped, ex[::-1])
        assert np.array_equal(res[1]["index"], kept) and np.array_equal(res[1]["descriptors"], desc), (name, "extra, flipped")
        # degenerate: no key point / every key point filtered
        res = vo.descKeypoints([np.zeros((0, 2), np.float32), np.array([[3.0, 3.0], [5.0, 100.0]], np.float32)], pair)
        assert res[0]["descriptors"].shape == (0, 32) and res[1]["descriptors"].shape == (0, 32)
        vo.close()


@pytest.mark.gpu
def test_cuda_process_image_chain_hits_the_opencv_vectors(detect_golden, orb_golden):
    """VisualOdometry::processImage over two frames, everything on the device after the image upload: the features of both frames,
    the ratio-tested matches and the matched pixel pairs handed to solveNlsAll equal cv2's."""
    import vloam_b200 as V
    from oracle import vo_frontend as F
    g = orb_golden
    v = _views(detect_golden)
    a, b = v["kitti"], v["kitti_next"]
    vo = V.VisualOdometry(batch=2, max_points=1024, max_matches=1024)
    frames = (np.stack([a, b[::-1].copy()]), np.stack([b, a[::-1].copy()]))       # stream 1: other images, the other way round
    vo.reset()
    r0 = vo.processImage(frames[0])
    assert r0["n_keypoints"][0] == len(g["kitti_desc"]) and np.array_equal(r0["n_matches"], [0, 0])
    vo.reset()
    r1 = vo.processImage(frames[1])
    assert r1["n_keypoints"][0] == len(g["kitti_next_desc"]) and r1["n_matches"][0] == len(g["chain_matches"])
    cur, prev = vo.frame_features(0), vo.frame_features(1)
    assert np.array_equal(prev[0]["descriptors"], g["kitti_desc"]) and np.array_equal(cur[0]["descriptors"], g["kitti_next_desc"])
    assert np.array_equal(prev[0]["keypoints"], g["kitti_corners"][g["kitti_kept_index"]])
    m = vo.matches()
    assert np.array_equal(m[0], g["chain_matches"])
    uq, ut = vo.match_uv()
    assert np.array_equal(uq[0, :len(m[0])], prev[0]["keypoints"][m[0][:, 0]]) and np.array_equal(ut[0, :len(m[0])], cur[0]["keypoints"][m[0][:, 1]])
    # stream 1 against the restatement
    feats = []
    for img in (frames[0][1], frames[1][1]):
        c = F.good_features_to_track(img)
        kept, desc = F.orb_describe(img, c)
        feats.append((c[kept], desc))
    assert np.array_equal(prev[1]["keypoints"], feats[0][0]) and np.array_equal(prev[1]["descriptors"], feats[0][1])
    assert np.array_equal(cur[1]["keypoints"], feats[1][0]) and np.array_equal(cur[1]["descriptors"], feats[1][1])
    assert np.array_equal(m[1], F.match_descriptors(feats[0][1], feats[1][1]))
    # a third frame: the slots swap again (query = frame 2's features)
    vo.reset()
    r2 = vo.processImage(frames[0])
    assert np.array_equal(vo.frame_features(1)[0]["descriptors"], g["kitti_next_desc"])
    assert np.array_equal(vo.matches()[0], F.match_descriptors(g["kitti_next_desc"], g["kitti_desc"]))
    assert r2["n_matches"][0] == len(vo.matches()[0])
    vo.close()
    # the same two frames from device memory, enqueued without any read-back (the call does not synchronise): same results
    import torch
    vd = V.VisualOdometry(batch=2, max_points=1024, max_matches=1024)
    for f in frames:
        vd.reset()
        assert vd.processImage(torch.from_numpy(f).cuda(), fetch=False) is None
    assert vd.detect_status() == 0
    assert np.array_equal(vd.matches()[0], g["chain_matches"]) and np.array_equal(vd.matches()[1], m[1])
    assert np.array_equal(vd.frame_features(0)[0]["descriptors"], g["kitti_next_desc"])
    vd.close()
    # ... and from pinned host memory: the frames are double-buffered and uploaded on the handle's copy stream while the previous
    # frame's kernels run; four frames enqueued back to back reuse both buffers
    vp = V.VisualOdometry(batch=2, max_points=1024, max_matches=1024)
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    for k in range(4):
        vp.reset()
        vp.processImage(pinned[k % 2], fetch=False)
    assert vp.detect_status() == 0
    assert np.array_equal(vp.matches()[0], g["chain_matches"]) and np.array_equal(vp.matches()[1], m[1])
    assert np.array_equal(vp.frame_features(0)[0]["descriptors"], g["kitti_next_desc"]) and np.array_equal(vp.frame_features(1)[0]["descriptors"], g["kitti_desc"])
    vp.close()


@pytest.mark.gpu
def test_cuda_front_end_against_restatement_across_image_sizes():
    """Detection, description and matching on images whose sizes exercise every tile shape of the kernels — smaller than one
    min-eigen tile (all CTAs mirror their accesses), one pixel short of / past the staged-interior condition, partial last tiles,
    widths that are and are not multiples of 32 (fused / unfused tails of the Sobel and blur passes), too small for ORB to keep any
    key point — against the numpy restatement: response map bit for bit, corners, surviving key points, descriptors, matches."""
    import vloam_b200 as V
    from oracle import vo_frontend as F
    rng = np.random.default_rng(20261017)
    for (h, w) in ((17, 23), (22, 70), (23, 71), (40, 69), (63, 63), (64, 100), (97, 131), (129, 257), (150, 352), (211, 1055)):
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        img = np.clip(np.round(0.25 * (img.astype(np.float32) + np.roll(img, 1, 0) + np.roll(img, 1, 1) + np.roll(img, (1, 1), (0, 1)))), 0, 255).astype(np.uint8)
        nxt = np.ascontiguousarray(np.roll(img, (1, -2), axis=(0, 1)))
        vo = V.VisualOdometry(batch=2, max_points=1024, max_matches=1024)
        feats = []
        for k, frame in enumerate((np.stack([img, nxt]), np.stack([nxt, img]))):
            vo.reset()
            r = vo.processImage(frame)
            cur = vo.frame_features(0)
            for b in range(2):
                resp = F.min_eigen_response(frame[b])
                assert np.array_equal(vo.corner_response(b).view(np.uint32), resp.view(np.uint32)), (h, w, k, b, "response map")
                corners = F.select_corners(resp)
                kept, desc = F.orb_describe(frame[b], corners)
                assert np.array_equal(cur[b]["keypoints"], corners[kept]) and np.array_equal(cur[b]["descriptors"], desc), (h, w, k, b)
                assert r["n_keypoints"][b] == len(kept)
                if h < 63 or w < 63:
                    assert len(kept) == 0
            feats.append(cur)
        m = vo.matches()
        for b in range(2):
            assert np.array_equal(m[b], F.match_descriptors(feats[0][b]["descriptors"], feats[1][b]["descriptors"])), (h, w, b)
        vo.close()
