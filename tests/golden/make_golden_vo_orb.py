"""Generates tests/golden/vo_orb_cv2.npz with OpenCV (cv2): the ORB description of the reference's front end and the whole
VisualOdometry::processImage chain, with the reference's calls (/root/reference/src/visual_odometry/src/image_util.cpp):
cv::goodFeaturesToTrack(img, 1024, 0.03, 7.5, Mat(), 5, false, 0.04) (:13-26), key points built from the corners (:29-35, angle
-1, octave 0), cv::ORB::create()->compute (:178, :204) and, for the frame pairs, cv::BFMatcher(NORM_HAMMING).knnMatch + the 0.8
ratio test (:228-283) in the order of visual_odometry.cpp:105-119 (query = previous frame, train = current frame).

The images are those of tests/golden/vo_detect_cv2.npz and pieces of them (`views()`): no new image is stored; the widths
cover OpenCV's vector / scalar-tail split of the blur's row pass (a multiple of 32, 32 k + 25, 32 k + 31).  `extra` key points
(non-integer, on and outside the border band ORB filters on, exact halves) exercise the key-point filter and cvRound.

Run from the repo root:  python tests/golden/make_golden_vo_orb.py      (needs cv2; version recorded in the file)
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def views(g):
    """name -> image: the committed detection images and deterministic pieces of them."""
    k = g["kitti_image"]
    return {"kitti": k, "small": g["small_image"], "shapes": g["shapes_image"],
            "kitti_w1087": np.ascontiguousarray(k[:300, 40:1127]),          # 32 * 33 + 31 columns: the widest scalar tail
            "kitti_next": next_frame(k)}


def next_frame(img):
    """The 'following' frame of the chain fixture: the image moved by (dx, dy) = (5, -2) pixels with a deterministic brightness
    ripple (numpy only, so the test rebuilds it bit for bit without cv2)."""
    out = np.roll(img, (-2, 5), axis=(0, 1)).astype(np.int32)
    yy, xx = np.mgrid[0:img.shape[0], 0:img.shape[1]]
    out += ((xx * 7 + yy * 13) % 5) - 2
    return np.clip(out, 0, 255).astype(np.uint8)


def extra_keypoints(h, w, seed):
    rng = np.random.default_rng(seed)
    a = np.stack([rng.random(400) * (w + 6) - 3, rng.random(400) * (h + 6) - 3], 1)                  # anywhere, also outside
    band = np.stack([rng.choice([30.4, 30.5, 30.6, 31.0, 31.5, w - 32.0, w - 31.5, w - 31.4, w - 31.0], 200), rng.random(200) * h], 1)
    band2 = np.stack([rng.random(200) * w, rng.choice([30.5, 31.0, 31.49, h - 32.5, h - 31.5, h - 31.0], 200)], 1)
    halves = np.stack([rng.integers(31, w - 31, 200) + 0.5, rng.integers(31, h - 31, 200) + 0.5], 1)   # cvRound: half to even
    return np.concatenate([a, band, band2, halves]).astype(np.float32)[:1024]


def describe(img, pts):
    kps = [cv2.KeyPoint(float(x), float(y), 5.0) for x, y in pts]                                     # image_util.cpp:29-35
    kept, desc = cv2.ORB_create().compute(img, kps)                                                   # :178, :204
    kept_pts = np.array([k.pt for k in kept], np.float32).reshape(-1, 2)
    if desc is None:
        desc = np.zeros((0, 32), np.uint8)
    # position of every survivor in the input list (ORB keeps the order)
    idx, j = [], 0
    for p in kept_pts:
        while not np.array_equal(pts[j], p):
            j += 1
        idx.append(j)
        j += 1
    return kept_pts, np.array(idx, np.int32), desc


def detect(img):
    c = cv2.goodFeaturesToTrack(img, 1024, 0.03, 7.5, None, blockSize=5, useHarrisDetector=False, k=0.04)   # :13-26
    return np.zeros((0, 2), np.float32) if c is None else c.reshape(-1, 2)


def build():
    g = np.load(os.path.join(HERE, "vo_detect_cv2.npz"))
    out = {"cv2_version": np.array(cv2.__version__)}
    feats = {}
    for n, (name, img) in enumerate(views(g).items()):
        corners = detect(img)
        kept, idx, desc = describe(img, corners)
        feats[name] = (kept, desc)
        out[f"{name}_corners"], out[f"{name}_kept_index"], out[f"{name}_desc"] = corners, idx, desc
        ex = extra_keypoints(img.shape[0], img.shape[1], 100 + n)
        ekept, eidx, edesc = describe(img, ex)
        out[f"{name}_extra"], out[f"{name}_extra_kept_index"], out[f"{name}_extra_desc"] = ex, eidx, edesc
        print(f"{name}: {img.shape}, {len(corners)} corners -> {len(kept)} described; extra {len(ex)} -> {len(ekept)}")
    # the processImage chain over two frames (visual_odometry.cpp:105-119): matches of the previous frame's descriptors in the current one's
    (k0, d0), (k1, d1) = feats["kitti"], feats["kitti_next"]
    knn = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False).knnMatch(d0, d1, 2)
    good = [m[0] for m in knn if m[0].distance < 0.8 * m[1].distance]
    out["chain_matches"] = np.array([[m.queryIdx, m.trainIdx, int(m.distance)] for m in good], np.int32).reshape(-1, 3)
    print("chain:", len(k0), "x", len(k1), "key points,", len(good), "matches")
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "vo_orb_cv2.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
