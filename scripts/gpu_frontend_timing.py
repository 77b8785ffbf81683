"""VisualOdometry::processImage on the device (vloam_vo_process_image: Shi-Tomasi detection, ORB description, BF matching) on a
batch of KITTI-sized images: wall time per call through the host API (image upload included) and per image, per-kernel CUDA-event
times, against the same three OpenCV calls on one host thread (the reference's processImage, visual_odometry.cpp:92-130).
usage (GPU box): python scripts/gpu_frontend_timing.py [batch]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vloam_b200 as V  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = np.load(os.path.join(ROOT, "tests", "golden", "vo_detect_cv2.npz"))
base = g["kitti_image"]
rng = np.random.default_rng(5)
shifts = [int(rng.integers(0, 300)) for _ in range(B)]
frames = [np.stack([np.roll(base, (0, s + 4 * k), axis=(0, 1)) for s in shifts]) for k in range(2)]      # frame k: 4 px further
vo = V.VisualOdometry(batch=B, max_points=1024, max_matches=1024)
for k in range(4):
    vo.reset()
    r = vo.processImage(frames[k % 2])
vo.ctx.enable_timing(True)
reps = 10
t0 = time.perf_counter()
for k in range(reps):
    vo.reset()
    r = vo.processImage(frames[k % 2])
t1 = time.perf_counter()
kt = vo.ctx.kernel_timings()
res = {"batch": B, "image": list(base.shape), "ms_per_call": (t1 - t0) / reps * 1e3, "ms_per_image": (t1 - t0) / reps * 1e3 / B,
       "kernel_ms_per_call": {k: v[0] / reps for k, v in kt.items()}, "kernel_launches_per_call": {k: v[1] / reps for k, v in kt.items()},
       "keypoints_stream0": int(r["n_keypoints"][0]), "matches_stream0": int(r["n_matches"][0]),
       "h2d_bytes_per_call": int(frames[0].nbytes)}
try:      # the same calls enqueued back to back from pinned memory, nothing read back: frames upload while the previous call computes
    import torch
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    n_pipe = 40
    for k in range(4):
        vo.reset(); vo.processImage(pinned[k % 2], fetch=False)
    vo.ctx.synchronize()
    t0 = time.perf_counter()
    for k in range(n_pipe):
        vo.reset(); vo.processImage(pinned[k % 2], fetch=False)
    vo.ctx.synchronize()
    res["pipelined_pinned_ms_per_call"] = (time.perf_counter() - t0) / n_pipe * 1e3
    res["pipelined_pinned_frames_per_s"] = B * n_pipe / (time.perf_counter() - t0)
    res["h2d_gbs_pipelined"] = frames[0].nbytes / (res["pipelined_pinned_ms_per_call"] * 1e-3) / 1e9
except ImportError:
    pass
try:
    import cv2
    cv2.setNumThreads(1)
    orb, bf = cv2.ORB_create(), cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)
    parts = {"detect": 0.0, "describe": 0.0, "match": 0.0}
    n = min(B, 6)
    for b in range(n):
        prev = None
        for k in range(2):
            img = frames[k][b]
            t = time.perf_counter()
            c = cv2.goodFeaturesToTrack(img, 1024, 0.03, 7.5, None, blockSize=5, useHarrisDetector=False, k=0.04).reshape(-1, 2)
            t2 = time.perf_counter()
            kps, d = orb.compute(img, [cv2.KeyPoint(float(x), float(y), 5.0) for x, y in c])
            t3 = time.perf_counter()
            if prev is not None:
                knn = bf.knnMatch(prev, d, 2)
                good = [m[0] for m in knn if m[0].distance < 0.8 * m[1].distance]
            t4 = time.perf_counter()
            parts["detect"] += t2 - t; parts["describe"] += t3 - t2; parts["match"] += (t4 - t3) * 2      # (matching runs on every second frame here)
            prev = d
    res["cv2_ms_per_image_1_thread"] = {k: v / (2 * n) * 1e3 for k, v in parts.items()}
    res["cv2_ms_per_image_1_thread"]["total"] = sum(res["cv2_ms_per_image_1_thread"].values())
except ImportError:
    pass
print(json.dumps(res))
vo.close()
