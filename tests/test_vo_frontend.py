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
        vo.close()
