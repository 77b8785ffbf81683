"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of the descriptor matching of the reference's visual odometry
front end (/root/reference/src/visual_odometry/src/image_util.cpp:214-296 in the configuration visual_odometry.cpp:34-37
selects: cv::BFMatcher(NORM_HAMMING), knnMatch k = 2, ratio test 0.8).

PARITY PINNED: tests/golden/vo_frontend_cv2.npz holds the outputs of OpenCV itself (the reference's dependency, cv2 4.13)
for the same calls; tests/test_vo_frontend.py checks this restatement against them bit for bit.
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
