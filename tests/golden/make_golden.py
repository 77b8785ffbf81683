"""Generates tests/golden/lidar_seed7_c256.npz from the CPU oracle (oracle/).

The reference ships no golden vectors and cannot be run here (parity unpinned), so the fixtures are
produced by the oracle restatement; they pin the oracle against regressions and give the CUDA path a committed target.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vloam_b200  # noqa: E402,F401
from oracle import pyoracle as O  # noqa: E402
from vloam_b200 import synth  # noqa: E402

SEED, COLS, NSCANS = 7, 256, 4


def build():
    stream = synth.ScanStream(SEED, n_cols=COLS)
    pipe = O.Pipeline()
    out = {}
    for k in range(NSCANS):
        scan = stream.scan(k)
        out[f"scan{k}_sha1"] = np.frombuffer(hashlib.sha1(scan.tobytes()).digest(), np.uint8)
        sr = O.scan_registration(scan)
        out[f"scan{k}_curvature"] = sr.curvature
        out[f"scan{k}_label"] = sr.label.astype(np.int8)
        out[f"scan{k}_sharpInd"] = sr.sharpInd
        out[f"scan{k}_lessSharpInd"] = sr.lessSharpInd
        out[f"scan{k}_flatInd"] = sr.flatInd
        out[f"scan{k}_lessFlat"] = sr.surfPointsLessFlat
        assert pipe.process(scan, do_mapping=True) == 0
        lo, lm = pipe.lo.state, pipe.lm.state
        out[f"scan{k}_lo_pose"] = np.r_[lo["q_last_curr"], lo["t_last_curr"], lo["q_w_curr"], lo["t_w_curr"]]
        out[f"scan{k}_lo_corr"] = np.array([lo["corner_correspondence"], lo["plane_correspondence"]])
        out[f"scan{k}_lm_pose"] = np.r_[lm["q_w_curr"], lm["t_w_curr"], lm["q_wmap_wodom"], lm["t_wmap_wodom"]]
        out[f"scan{k}_map_points"] = np.array([pipe.lm.map_points(0), pipe.lm.map_points(1)])
        # the map after this scan: every occupied cube (kind, cube index, points) with the sha1 of its x, y, z bits in order
        rows, digests = [], []
        for kind in (0, 1):
            for cube in range(4851):
                n = pipe.lm.cube_count(kind, cube)
                if n:
                    rows.append((kind, cube, n))
                    digests.append(np.frombuffer(hashlib.sha1(np.ascontiguousarray(pipe.lm.cube(kind, cube)[:, :3]).tobytes()).digest(), np.uint8))
        out[f"scan{k}_cubes"] = np.array(rows, np.int32)
        out[f"scan{k}_cube_sha1"] = np.stack(digests)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lidar_seed7_c256.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
