"""Generates tests/golden/vo_frontend_cv2.npz with OpenCV (cv2), the library the reference's ImageUtil calls.

Unlike the LiDAR fixtures (produced by the oracle restatement because the reference cannot be built here), these vectors
come from the reference's own dependency: the exact calls of /root/reference/src/visual_odometry/src/image_util.cpp —
cv::goodFeaturesToTrack(img, 1024, 0.03, 7.5, Mat(), 5, false, 0.04) (:13-26), cv::ORB::create()->compute (:176-178, 204) and
cv::BFMatcher(NORM_HAMMING).knnMatch(d0, d1, 2) + the 0.8 ratio test (:228-283) — on synthetic KITTI-sized image pairs.
They PIN the descriptor-matching row: tests/test_vo_frontend.py checks the numpy restatement
and the CUDA kernel against them bit for bit.

Run from the repo root:  python tests/golden/make_golden_vo_frontend.py      (needs cv2; version recorded in the file)
"""
import os

import cv2
import numpy as np

H, W = 376, 1241


def synthetic_image(rng):
    img = np.full((H, W), 90, np.uint8)
    for _ in range(260):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        c = int(rng.integers(20, 235))
        if rng.random() < 0.5:
            cv2.rectangle(img, (x, y), (x + int(rng.integers(6, 70)), y + int(rng.integers(6, 50))), c, -1)
        else:
            cv2.circle(img, (x, y), int(rng.integers(3, 25)), c, -1)
    img = cv2.GaussianBlur(img, (5, 5), 1.0)
    return img


def next_frame(img, rng, shift):
    M = np.float32([[1.0, 0.004, shift[0]], [-0.004, 1.0, shift[1]]])
    out = cv2.warpAffine(img, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    noise = rng.normal(0.0, 2.0, out.shape)
    return np.clip(out.astype(np.float64) + noise, 0, 255).astype(np.uint8)


def detect_describe(img):
    corners = cv2.goodFeaturesToTrack(img, 1024, 0.03, 7.5, None, blockSize=5, useHarrisDetector=False, k=0.04)   # image_util.cpp:13-26
    kps = [cv2.KeyPoint(float(x), float(y), 5.0) for x, y in corners.reshape(-1, 2)]                                  # :29-35
    kps, desc = cv2.ORB_create().compute(img, kps)                                                                   # :176-178, :204
    pts = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    return pts, desc


def build():
    out = {"cv2_version": np.array(cv2.__version__)}
    rng = np.random.default_rng(20260117)
    for p, shift in enumerate([(3.5, -1.25), (-6.0, 2.0), (0.75, 0.5)]):
        a = synthetic_image(rng)
        b = next_frame(a, rng, shift)
        kp0, d0 = detect_describe(a)
        kp1, d1 = detect_describe(b)
        if p == 2:      # a small train set with duplicated rows: ties for both neighbours, decided by the train index
            d1 = np.concatenate([d1[:40], d1[:40][::-1], d1[:40]])
            kp1 = np.concatenate([kp1[:40], kp1[:40][::-1], kp1[:40]])
        knn = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False).knnMatch(d0, d1, 2)                                    # :228-263
        full = np.array([[m[0].queryIdx, m[0].trainIdx, m[1].trainIdx] for m in knn], np.int32)
        dist = np.array([[m[0].distance, m[1].distance] for m in knn], np.float32)
        good = [m[0] for m in knn if m[0].distance < 0.8 * m[1].distance]                                             # :275-281
        out[f"pair{p}_desc0"], out[f"pair{p}_desc1"] = d0, d1
        out[f"pair{p}_kp0"], out[f"pair{p}_kp1"] = kp0, kp1
        out[f"pair{p}_knn_idx"], out[f"pair{p}_knn_dist"] = full, dist
        out[f"pair{p}_matches"] = np.array([[m.queryIdx, m.trainIdx, int(m.distance)] for m in good], np.int32).reshape(-1, 3)
        print(f"pair {p}: {len(kp0)} x {len(kp1)} keypoints, {len(good)} matches after the ratio test")
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vo_frontend_cv2.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
