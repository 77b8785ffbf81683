"""Data formats on either side of the hot path — Python mirror of adapter/wire_formats.hpp.

    load_kitti_bin        KITTI velodyne .bin -> (n, 4) float32       reference point_cloud_util.cpp:118-146
    pointcloud2_xyz       sensor_msgs/PointCloud2 payload -> a strided (n, stride) float32 VIEW (no copy when x, y, z are
                          consecutive float32 fields) that scanRegistrationIO accepts    reference vloam_main_node.cpp:148
    Cam0StartFrameWriter  KITTI-format pose dump                       reference vloam_tf.cpp:77-153
"""
from __future__ import annotations

import numpy as np

KITTI_BIN_MAX_FLOATS = 1_000_000     # the reference's fixed read buffer (point_cloud_util.cpp:120)


def load_kitti_bin(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.float32, count=KITTI_BIN_MAX_FLOATS)
    n = raw.size // 4
    return raw[: 4 * n].reshape(n, 4)


def pointcloud2_xyz(data, n_points: int, point_step: int, offsets: dict, is_bigendian: bool = False):
    """data: bytes / uint8 array of the message; offsets: {'x': .., 'y': .., 'z': ..} byte offsets of float32 fields.
    Returns (array, zero_copy): array[:, :3] are x, y, z; rows are point_step / 4 floats apart when zero_copy."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
    ox, oy, oz = offsets["x"], offsets["y"], offsets["z"]
    if is_bigendian:
        raise ValueError("big-endian PointCloud2 is not supported")
    if oy == ox + 4 and oz == ox + 8 and point_step % 4 == 0 and ox % 4 == 0:
        flat = buf[: n_points * point_step].view(np.float32).reshape(n_points, point_step // 4)
        return flat[:, ox // 4:], True
    rows = buf[: n_points * point_step].reshape(n_points, point_step)
    out = np.empty((n_points, 3), np.float32)
    for k, o in enumerate((ox, oy, oz)):
        out[:, k] = np.ascontiguousarray(rows[:, o:o + 4]).view(np.float32)[:, 0]
    return out, False


def mat4_from_qt(q, t) -> np.ndarray:
    x, y, z, w = (float(v) for v in q)
    m = np.eye(4)
    m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    m[:3, 3] = t
    return m


def _rigid_inverse(m: np.ndarray) -> np.ndarray:
    r = np.eye(4)
    r[:3, :3] = m[:3, :3].T
    r[:3, 3] = -r[:3, :3] @ m[:3, 3]
    return r


class Cam0StartFrameWriter:
    """VloamTF::{VO,LO,MO}2Cam0StartFrame: poses relative to the first dumped frame, in the cam0 frame, one line of the
    float-cast 3 x 4 matrix per frame printed with "%f"."""

    def __init__(self, base_T_cam0: np.ndarray):
        self.base_T_cam0 = np.asarray(base_T_cam0, np.float64)
        self.cam0_T_base = _rigid_inverse(self.base_T_cam0)
        self.start_T_init = np.eye(4)

    def write(self, fp, count: int, world_T_base_last: np.ndarray) -> str:
        if count < 0:
            return ""
        init_T_last = self.cam0_T_base @ np.asarray(world_T_base_last, np.float64) @ self.base_T_cam0
        if count == 0:
            self.start_T_init = _rigid_inverse(init_T_last)
        m = (self.start_T_init @ init_T_last)[:3, :4].astype(np.float32)
        line = " ".join("%f" % float(v) for v in m.reshape(-1)) + "\n"
        if fp is not None:
            fp.write(line)
        return line
