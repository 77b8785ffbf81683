"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of two stages of the reference's visual odometry front end
(/root/reference/src/visual_odometry/src/image_util.cpp in the configuration visual_odometry.cpp:34-37 selects):
  * key-point detection, image_util.cpp:5-37: cv::goodFeaturesToTrack(img, 1024, 0.03, 7.5, Mat(), 5, false, 0.04), i.e. the
    Shi-Tomasi minimum-eigenvalue response (cv::cornerMinEigenVal) + threshold / 3x3 local maxima / sort / greedy spacing;
  * descriptor matching, image_util.cpp:214-296: cv::BFMatcher(NORM_HAMMING), knnMatch k = 2, ratio test 0.8.

PARITY PINNED: tests/golden/vo_frontend_cv2.npz and tests/golden/vo_detect_cv2.npz hold the outputs of OpenCV itself (the
reference's dependency, cv2 4.13) for the same calls; tests/test_vo_frontend.py checks this restatement against them: matches
and corner lists bit for bit, the response map bit for bit on > 99.99 % of the pixels (the rest differ in the last place:
OpenCV's box filter keeps running double-precision sums per thread stripe, whose rounding depends on the stripe layout).
"""
import numpy as np

_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_matrix(d0, d1):
    """(n0, n1) Hamming distances between the rows of two uint8 descriptor matrices (cv::NORM_HAMMING)."""
    return _POP[np.bitwise_xor(d0[:, None, :], d1[None, :, :])].sum(axis=2)


def knn2(d0, d1):
    """cv::BFMatcher::knnMatch(d0, d1, 2): per query row the two smallest (distance, train index) pairs — OpenCV's
    batchDistance admits a candidate only when strictly closer than the current k-th and keeps it behind equal
    distances, which is that order.  Returns idx (n0, 2) int32 and dist (n0, 2) int32 (-1 where the train set is too small)."""
    D = hamming_matrix(d0, d1)
    n0, n1 = D.shape
    idx = np.full((n0, 2), -1, np.int32)
    dist = np.full((n0, 2), -1, np.int32)
    if n1:
        order = np.argsort(D, axis=1, kind="stable")[:, :2]          # stable: equal distances keep ascending train index
        k = order.shape[1]
        idx[:, :k] = order
        dist[:, :k] = np.take_along_axis(D, order, axis=1)
    return idx, dist


def match_descriptors(d0, d1, ratio=0.8):
    """image_util.cpp:263-281: (queryIdx, trainIdx, distance) of the matches that pass `m0.distance < ratio * m1.distance`
    (a float compared with a double product), in query order."""
    idx, dist = knn2(d0, d1)
    out = []
    for q in range(idx.shape[0]):
        if idx[q, 1] >= 0 and float(np.float32(dist[q, 0])) < ratio * float(np.float32(dist[q, 1])):
            out.append((q, int(idx[q, 0]), int(dist[q, 0])))
    return np.array(out, np.int32).reshape(-1, 3)


# ------------------------------------------------------------------------------------------------ key-point detection
def _fma(a, b, c):
    """float32 fused multiply-add (the product of two float32 is exact in float64; one rounding to float64 before the final
    rounding to float32 — adequate for the magnitudes here and checked against OpenCV's output)."""
    return (np.float64(a) * np.float64(b) + np.float64(c)).astype(np.float32)


def _reflect101(i, n):
    """cv::BORDER_REFLECT_101 index map for i in [-n + 1, 2n - 2]."""
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def min_eigen_response(img, block=5):
    """cv::cornerMinEigenVal(img, block, 3) for an 8-bit image (imgproc corner.cpp: cornerEigenValsVecs + calcMinEigenVal), with
    the operation order of OpenCV 4.13's vectorised code path (what cv2 runs on AVX2 / AVX-512 hosts):
      scale = 1 / (2^(3-1) * block * 255); Sobel kernels with the scale folded into the smoothing taps (s, 2s, s);
      Dx = fma(s, r[y-1] + r[y+1], 2s * r[y])          with r = I[x+1] - I[x-1]   (exact)
      Dy = row[y+1] - row[y-1]                         with row = fma(s, I[x+1], fma(2s, I[x], s * I[x-1])) — except in the last
           W mod 32 columns, which OpenCV's scalar tail loop computes without fused operations;
      (Dx^2, Dx Dy, Dy^2) box-summed over block x block in double precision (rows, then columns), rounded to float;
      response = (a + c) - sqrt((a - c)^2 + b^2) with a = Sxx / 2, b = Sxy, c = Syy / 2, plain float operations."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    sc = 1.0 / (4 * block * 255.0)
    s1 = np.float32(sc)
    s2 = np.float32(2) * s1
    p = np.pad(img.astype(np.float32), 1, mode="reflect")
    r = p[:, 2:] - p[:, :-2]
    Dx = _fma(s1, r[:-2] + r[2:], (s2 * r[1:-1]).astype(np.float32))
    c0, c1, c2 = p[:, :-2], p[:, 1:-1], p[:, 2:]
    rows = _fma(s1, c2, _fma(s2, c1, (s1 * c0).astype(np.float32)))
    tail = W - W % 32
    if tail < W:
        rows[:, tail:] = (((s1 * c0[:, tail:]).astype(np.float32) + (s2 * c1[:, tail:]).astype(np.float32)).astype(np.float32)
                          + (s1 * c2[:, tail:]).astype(np.float32)).astype(np.float32)
    Dy = (rows[2:] - rows[:-2]).astype(np.float32)
    cov = np.stack([Dx * Dx, Dx * Dy, Dy * Dy], -1).astype(np.float32)
    h = block // 2
    pc = np.pad(cov.astype(np.float64), ((h, h), (h, h), (0, 0)), mode="reflect")
    rs = np.zeros((H + 2 * h, W, 3))
    for dx in range(block):
        rs += pc[:, dx:dx + W]
    acc = np.zeros((H, W, 3))
    for dy in range(block):
        acc += rs[dy:dy + H]
    box = acc.astype(np.float32)
    a = box[..., 0] * np.float32(0.5)
    b = box[..., 1]
    c = box[..., 2] * np.float32(0.5)
    t = (a - c).astype(np.float32)
    return ((a + c).astype(np.float32) - np.sqrt(((t * t).astype(np.float32) + (b * b).astype(np.float32)).astype(np.float32))).astype(np.float32)


def select_corners(eig, max_corners=1024, quality=0.03, min_distance=7.5):
    """The rest of cv::goodFeaturesToTrack (imgproc featureselect.cpp) on a response map: threshold at quality * max (to zero),
    3 x 3 local maxima (value equal to its dilation) away from the one-pixel border, std::sort with greaterThanPtr — value
    descending, equal values by descending address —, then the greedy pass that keeps a corner when no kept corner lies
    closer than min_distance, stopped at max_corners.  Returns (n, 2) float32 (x, y)."""
    eig = np.ascontiguousarray(eig, np.float32)
    H, W = eig.shape
    thr = np.float32(np.float64(eig.max()) * quality)
    e = np.where(eig > thr, eig, np.float32(0))
    pe = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = np.max(np.stack([pe[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)]), 0)
    m = (e != 0) & (e == dil)
    m[0, :] = m[-1, :] = False
    m[:, 0] = m[:, -1] = False
    ys, xs = np.nonzero(m)
    vals = e[ys, xs]
    ofs = ys * W + xs
    order = np.lexsort((-ofs, -vals.astype(np.float64)))
    cell = int(np.rint(min_distance))                      # cvRound: half to even
    md2 = min_distance * min_distance
    gw, gh = (W + cell - 1) // cell, (H + cell - 1) // cell
    grid = {}
    out = []
    for k in order:
        y, x = int(ys[k]), int(xs[k])
        xc, yc = x // cell, y // cell
        good = True
        for yy in range(max(0, yc - 1), min(gh - 1, yc + 1) + 1):
            for xx in range(max(0, xc - 1), min(gw - 1, xc + 1) + 1):
                for (px, py) in grid.get((yy, xx), ()):
                    if (x - px) * (x - px) + (y - py) * (y - py) < md2:
                        good = False
                        break
                if not good:
                    break
            if not good:
                break
        if good:
            grid.setdefault((yc, xc), []).append((x, y))
            out.append((x, y))
            if len(out) == max_corners:
                break
    return np.array(out, np.float32).reshape(-1, 2)


def good_features_to_track(img, max_corners=1024, quality=0.03, min_distance=7.5, block=5):
    """ImageUtil::detKeypoints with DetectorType::ShiTomasi (image_util.cpp:11-37): corner coordinates (x, y) in pick order."""
    return select_corners(min_eigen_response(img, block), max_corners, quality, min_distance)


# ------------------------------------------------------------------ ORB description (image_util.cpp:162-212)
ORB_EDGE = 31                                             # cv::ORB::create(): edgeThreshold 31, patchSize 31, WTA_K 2, first level 0
def _gauss7_taps():
    """cv::getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) in double, normalised in double, stored as float."""
    x = np.arange(7, dtype=np.float64) - 3.0
    k = np.exp(-(x * x) / 8.0)
    return (k / k.sum()).astype(np.float32)


def _mad(a, b, c):
    return (a * np.float32(b) + c).astype(np.float32)


def gaussian_blur_7x7(img):
    """The blur inside cv::ORB::compute: GaussianBlur(level, level, Size(7, 7), 2, 2, BORDER_REFLECT_101) on a SUB-MATRIX of the
    pyramid buffer.  For a sub-matrix OpenCV skips its 8-bit fixed-point Gaussian and calls sepFilter2D with the float taps:
    rows first (uchar -> float, taps in order 0..6), then columns (centre tap, then the three symmetric pairs
    tap_k * (row[+k] + row[-k])), one cvRound to 8 bits.  The rounding sequence below is that of cv2 4.13's AVX2 build, found
    by experiment and pinned by tests/test_vo_frontend.py against cv2.sepFilter2D / cv2.ORB on many image widths: the
    vectorised loops fuse multiply-adds, the scalar tails do not — rows: columns x < 32 * (w // 32) fused, the rest unfused;
    columns: x < 4 * (w // 4) fused, the last w % 4 unfused."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    k = _gauss7_taps()
    p = np.pad(img, 3, mode="reflect").astype(np.float32)            # BORDER_REFLECT_101
    wv = 32 * (w // 32)
    rows = np.empty((h + 6, w), np.float32)
    for lo, hi, op in ((0, wv, _fma), (wv, w, _mad)):
        if hi > lo:
            acc = (p[:, lo:hi] * k[0]).astype(np.float32)
            for t in range(1, 7):
                acc = op(p[:, lo + t:hi + t], k[t], acc)
            rows[:, lo:hi] = acc
    out = np.empty((h, w), np.float32)
    wc = 4 * (w // 4)
    for lo, hi, op in ((0, wc, _fma), (wc, w, _mad)):
        if hi > lo:
            acc = (rows[3:3 + h, lo:hi] * k[3]).astype(np.float32)
            for t in (1, 2, 3):
                acc = op((rows[3 + t:3 + t + h, lo:hi] + rows[3 - t:3 - t + h, lo:hi]).astype(np.float32), k[3 + t], acc)
            out[:, lo:hi] = acc
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def orb_steered_pattern(angle_deg=-1.0):
    """The 512 sample offsets of computeOrbDescriptors for a key point of angle `angle_deg` (cv::KeyPoint's default -1: the
    reference builds its key points from corners, image_util.cpp:29-35, so ORB never estimates an orientation): the learned
    pattern rotated in float and rounded (cvRound).  For -1 degree every offset rounds back to the table entry."""
    try:
        from .orb_pattern import ORB_PATTERN
    except ImportError:           # run as a plain module from inside oracle/
        from orb_pattern import ORB_PATTERN
    pat = np.array(ORB_PATTERN, np.float32).reshape(-1, 2)            # (512, 2) x, y
    ang = np.float32(angle_deg) * np.float32(np.pi / 180.0)
    a, b = np.float32(np.cos(np.float64(ang))), np.float32(np.sin(np.float64(ang)))
    x = pat[:, 0] * a - pat[:, 1] * b
    y = pat[:, 0] * b + pat[:, 1] * a
    return np.rint(x).astype(np.int32), np.rint(y).astype(np.int32)


def orb_describe(img, keypoints_xy):
    """ImageUtil::descKeypoints with DescriptorType::ORB (image_util.cpp:162-212): cv::ORB::create()->compute(img, keypoints,
    descriptors) on key points of octave 0 and angle -1.  Returns (kept, descriptors): the indices of the key points ORB keeps
    (KeyPointsFilter::runByImageBorder with the edge threshold, on the rounded position: 31 <= cvRound(x) < cols - 31 and
    31 <= cvRound(y) < rows - 31; the reference's key-point vector is rewritten to these, in order) and their 32-byte descriptors: bit i of byte j is
    blurred(c + a) < blurred(c + b) for pair 8 j + i, c = (cvRound(x), cvRound(y))."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    kp = np.asarray(keypoints_xy, np.float32).reshape(-1, 2)
    rx, ry = np.rint(kp[:, 0]).astype(np.int64), np.rint(kp[:, 1]).astype(np.int64)      # Rect_<int>::contains(Point(pt)): cvRound
    keep = (rx >= ORB_EDGE) & (rx < w - ORB_EDGE) & (ry >= ORB_EDGE) & (ry < h - ORB_EDGE)
    kept = np.nonzero(keep)[0].astype(np.int32)
    if len(kept) == 0 or w <= 2 * ORB_EDGE or h <= 2 * ORB_EDGE:
        return np.zeros(0, np.int32), np.zeros((0, 32), np.uint8)
    blur = gaussian_blur_7x7(img)
    cx, cy = rx[kept], ry[kept]
    ox, oy = orb_steered_pattern()
    v = blur[cy[:, None] + oy[None, :], cx[:, None] + ox[None, :]]     # (n, 512)
    bits = (v[:, 0::2] < v[:, 1::2]).astype(np.uint8)                 # (n, 256)
    return kept, np.packbits(bits, axis=1, bitorder="little")
