"""Generates tests/golden/vo_detect_cv2.npz with OpenCV (cv2): the key-point detection of the reference's front end,
cv::goodFeaturesToTrack(img, 1024, 0.03, 7.5, Mat(), 5, false, 0.04) (/root/reference/src/visual_odometry/src/image_util.cpp:13-26),
and the response map it thresholds (cv::cornerMinEigenVal(img, 5, 3)), on synthetic images: one KITTI-sized (376 x 1241: the
width is not a multiple of 32, so OpenCV's scalar tail columns are covered), one 200 x 640 whose full response map is stored,
one with flat regions (few corners, equal responses).

Run from the repo root:  python tests/golden/make_golden_vo_detect.py      (needs cv2; version recorded in the file)
"""
import os

import cv2
import numpy as np


def textured(rng, h, w, sigma):
    img = (rng.random((h, w)) * 255).astype(np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), sigma)
    return cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)


def shapes(rng, h, w):
    img = np.full((h, w), 90, np.uint8)
    for _ in range(120):
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        c = int(rng.integers(20, 235))
        cv2.rectangle(img, (x, y), (x + int(rng.integers(6, 70)), y + int(rng.integers(6, 50))), c, -1)
    return img


def build():
    rng = np.random.default_rng(20260217)
    images = {"kitti": textured(rng, 376, 1241, 2.0), "small": textured(rng, 200, 640, 1.5), "shapes": shapes(rng, 240, 800)}
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, img in images.items():
        eig = cv2.cornerMinEigenVal(img, 5, ksize=3)
        corners = cv2.goodFeaturesToTrack(img, 1024, 0.03, 7.5, None, blockSize=5, useHarrisDetector=False, k=0.04)
        corners = np.zeros((0, 2), np.float32) if corners is None else corners.reshape(-1, 2)
        out[f"{name}_image"] = img
        out[f"{name}_corners"] = corners
        out[f"{name}_response_max"] = np.array(eig.max(), np.float32)
        if name == "small":
            out[f"{name}_response"] = eig
        else:      # every 8th row of the response map keeps the fixture small
            out[f"{name}_response_rows"] = eig[::8].copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vo_detect_cv2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.endswith("corners")})


if __name__ == "__main__":
    build()
