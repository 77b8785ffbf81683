"""Key-point detection (vloam_vo_detect_corners) on a batch of KITTI-sized images: wall time per call through the host API
(image upload + three kernels + corner read-back) and per image, against cv2.goodFeaturesToTrack on the host.
usage (GPU box): python scripts/gpu_detect_timing.py [batch]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vloam_b200 as V  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vo_detect_cv2.npz"))
base = g["kitti_image"]
rng = np.random.default_rng(5)
imgs = np.stack([np.roll(base, int(rng.integers(0, 300)), axis=1) for _ in range(B)])
vo = V.VisualOdometry(batch=B, max_points=1024, max_matches=1024)
vo.detKeypoints(imgs)
vo.ctx.enable_timing(True)
reps = 10
t0 = time.perf_counter()
for _ in range(reps):
    out = vo.detKeypoints(imgs)
t1 = time.perf_counter()
kt = vo.ctx.kernel_timings()
res = {"batch": B, "image": list(base.shape), "ms_per_call": (t1 - t0) / reps * 1e3, "ms_per_image": (t1 - t0) / reps * 1e3 / B,
       "kernel_ms_per_call": {k: v[0] / reps for k, v in kt.items()}, "corners_stream0": int(len(out[0]))}
try:
    import cv2
    cv2.setNumThreads(1)
    t0 = time.perf_counter()
    for b in range(min(B, 8)):
        cv2.goodFeaturesToTrack(imgs[b], 1024, 0.03, 7.5, None, blockSize=5, useHarrisDetector=False, k=0.04)
    res["cv2_ms_per_image_1_thread"] = (time.perf_counter() - t0) / min(B, 8) * 1e3
except ImportError:
    pass
print(json.dumps(res))
vo.close()
