"""Synthetic HDL-64 scan generator (SURVEY.md §8d "Synthetic input").

Test / bench infrastructure only: produces deterministic 64x2048 range-image
scans of a seeded city scene (ground plane + axis-aligned boxes + vertical
poles) seen from a moving sensor, emitted ring-major in firing order the way
the reference's scanRegistration expects them
(/root/reference/src/lidar_odometry_mapping/src/scan_registration.cpp:166-176,
213-226: start/end azimuth logic and the HDL-64 elevation -> ring table).

Ring r has elevation  2 - r/3 - 0.05 deg          (r = 0..32)
                     -8.83 - (r-32)/2 - 0.05 deg  (r = 33..63)
i.e. the bin centres of the reference's ring formula shifted 0.05 deg away
from the bin edges so the ring id is unambiguous in float arithmetic.
Rings 51..63 are generated (the input really is 64x2048 = 131072 points) and
are dropped by the reference's own `scanID > 50` rule.
"""
from __future__ import annotations

import ctypes as _C
import os as _os
import subprocess as _sp

import numpy as np

N_RINGS = 64
N_COLS = 2048
GROUND_Z = -1.73
R_MIN = 2.0
R_MAX = 80.0


def ring_elevations_deg() -> np.ndarray:
    r = np.arange(N_RINGS, dtype=np.float64)
    upper = 2.0 - r / 3.0 - 0.05
    lower = -8.83 - (r - 32.0) / 2.0 - 0.05
    return np.where(r <= 32, upper, lower)


def _ray_dirs(n_cols: int = N_COLS) -> np.ndarray:
    """Unit ray directions in the sensor frame, shape (64, n_cols, 3).

    Firing azimuth decreases with the column index (the Velodyne spins
    clockwise), so the reference's ori = -atan2(y, x) increases from about
    -pi to +pi over one ring.
    """
    el = np.deg2rad(ring_elevations_deg())[:, None]
    c = np.arange(n_cols, dtype=np.float64)[None, :]
    az = np.pi - (c + 0.5) * (2.0 * np.pi / n_cols)
    ce = np.cos(el)
    d = np.stack([ce * np.cos(az), ce * np.sin(az), np.sin(el) * np.ones_like(az)], axis=-1)
    return d


_HERE = _os.path.dirname(_os.path.abspath(__file__))
RAYCAST_SRC = _os.path.join(_HERE, "csrc_host", "synth_raycast.c")
RAYCAST_LIB = _os.path.join(_HERE, "lib", "libsynth_raycast.so")
_raycast = [False, None]     # [probed, library or None]


def build_raycast() -> str:
    """gcc -O2 -ffp-contract=off (no FMA: the numpy code it restates has none) -> lib/libsynth_raycast.so."""
    _os.makedirs(_os.path.dirname(RAYCAST_LIB), exist_ok=True)
    if not _os.path.exists(RAYCAST_LIB) or _os.path.getmtime(RAYCAST_LIB) < _os.path.getmtime(RAYCAST_SRC):
        _sp.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", RAYCAST_SRC, "-o", RAYCAST_LIB, "-lm"])
    return RAYCAST_LIB


def _raycast_lib():
    if not _raycast[0]:
        _raycast[0] = True
        if _os.environ.get("VLOAM_SYNTH_NUMPY") != "1":
            try:
                L = _C.CDLL(build_raycast())
                dp = _C.POINTER(_C.c_double)
                L.synth_raycast.restype = None
                L.synth_raycast.argtypes = [dp, dp, _C.c_long, dp, dp, _C.c_int, dp, dp, dp, _C.c_int, _C.c_double, dp]
                _raycast[1] = L
            except Exception:
                _raycast[1] = None
    return _raycast[1]


class Scene:
    """Seeded static world: ground plane, boxes (buildings), poles."""

    def __init__(self, seed: int, n_boxes: int = 40, n_poles: int = 60, extent: float = 120.0):
        rng = np.random.Generator(np.random.Philox(key=[seed, 0xC17E]))
        # boxes: centre (x, y), half sizes, height; keep a corridor |y| < 4 m free (the road)
        cx = rng.uniform(-extent, extent, n_boxes)
        side = rng.choice([-1.0, 1.0], n_boxes)
        hy = rng.uniform(2.5, 15.0, n_boxes)
        hx = rng.uniform(2.5, 15.0, n_boxes)
        cy = side * (rng.uniform(6.0, 40.0, n_boxes) + hy)
        h = rng.uniform(3.0, 25.0, n_boxes)
        self.box_min = np.stack([cx - hx, cy - hy, np.full(n_boxes, GROUND_Z)], axis=1)
        self.box_max = np.stack([cx + hx, cy + hy, GROUND_Z + h], axis=1)
        px = rng.uniform(-extent, extent, n_poles)
        py = rng.choice([-1.0, 1.0], n_poles) * rng.uniform(4.5, 30.0, n_poles)
        self.pole_xy = np.stack([px, py], axis=1)
        self.pole_r = np.full(n_poles, 0.15)
        self.pole_top = GROUND_Z + rng.uniform(3.0, 8.0, n_poles)

    def raycast(self, origin: np.ndarray, dirs: np.ndarray) -> np.ndarray:
        """Range along each ray (inf for no hit). dirs: (M, 3) unit, world frame.  Uses the C restatement
        (csrc_host/synth_raycast.c, same IEEE operations in the same order: identical bits) when it is built."""
        L = _raycast_lib()
        if L is not None:
            o = np.ascontiguousarray(origin, np.float64)
            d = np.ascontiguousarray(dirs, np.float64)
            best = np.empty(d.shape[0], np.float64)
            dp = _C.POINTER(_C.c_double)
            arrs = [np.ascontiguousarray(a, np.float64) for a in (self.box_min, self.box_max, self.pole_xy, self.pole_r, self.pole_top)]
            L.synth_raycast(o.ctypes.data_as(dp), d.ctypes.data_as(dp), d.shape[0], arrs[0].ctypes.data_as(dp), arrs[1].ctypes.data_as(dp),
                            self.box_min.shape[0], arrs[2].ctypes.data_as(dp), arrs[3].ctypes.data_as(dp), arrs[4].ctypes.data_as(dp),
                            self.pole_xy.shape[0], GROUND_Z, best.ctypes.data_as(dp))
            return best
        return self.raycast_numpy(origin, dirs)

    def raycast_numpy(self, origin: np.ndarray, dirs: np.ndarray) -> np.ndarray:
        """The reference implementation of raycast (numpy); kept as the fallback and as the checker of the C version."""
        o = origin.astype(np.float64)
        d = dirs
        M = d.shape[0]
        best = np.full(M, np.inf)
        # ground
        with np.errstate(divide="ignore", invalid="ignore"):
            tg = (GROUND_Z - o[2]) / d[:, 2]
        tg = np.where((d[:, 2] < 0) & (tg > 0), tg, np.inf)
        best = np.minimum(best, tg)
        # boxes (slab test), chunked over boxes to bound memory
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
        for b in range(self.box_min.shape[0]):
            t0 = (self.box_min[b] - o) * inv
            t1 = (self.box_max[b] - o) * inv
            tmin = np.minimum(t0, t1).max(axis=1)
            tmax = np.maximum(t0, t1).min(axis=1)
            hit = (tmax >= tmin) & (tmax > 0)
            t = np.where(tmin > 0, tmin, tmax)
            best = np.where(hit & (t < best), t, best)
        # poles (vertical cylinders)
        dxy = d[:, :2]
        a = (dxy * dxy).sum(axis=1)
        for p in range(self.pole_xy.shape[0]):
            oc = o[:2] - self.pole_xy[p]
            bq = dxy[:, 0] * oc[0] + dxy[:, 1] * oc[1]          # (no BLAS: its FMA use differs between CPUs)
            cq = (oc[0] * oc[0] + oc[1] * oc[1]) - self.pole_r[p] * self.pole_r[p]
            disc = bq * bq - a * cq
            ok = disc > 0
            sq = np.sqrt(np.where(ok, disc, 0.0))
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (-bq - sq) / a
            z = o[2] + t * d[:, 2]
            hit = ok & (t > 0) & (z <= self.pole_top[p]) & (z >= GROUND_Z)
            best = np.where(hit & (t < best), t, best)
        return best


def _mm(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a @ b for (n, 3) x (3, m) with plain multiplies and adds in a fixed order.  BLAS kernels fuse multiply-adds
    differently from CPU to CPU; the generator must give the same bits everywhere (tests/golden pins them)."""
    return (a[:, 0:1] * b[0] + a[:, 1:2] * b[1]) + a[:, 2:3] * b[2]


def _mv(a: np.ndarray, v: np.ndarray) -> np.ndarray:
    return (a[:, 0] * v[0] + a[:, 1] * v[1]) + a[:, 2] * v[2]


def _rot_z(yaw: float) -> np.ndarray:
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rot_small(roll: float, pitch: float) -> np.ndarray:
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    return _mm(ry, rx)


class ScanStream:
    """One synthetic LiDAR sequence: scene + ego-motion + per-scan noise.

    poses[k] = (R_k, t_k): sensor frame k -> world.  Scans are rendered from
    the pose at the scan's end (no intra-scan motion distortion; the reference
    runs with DISTORTION=false, laser_odometry.h:90).
    """

    def __init__(self, seed: int, n_cols: int = N_COLS, noise_sigma: float = 0.01, yaw_rate_max: float = 0.2, on_road: bool = False):
        self.seed = int(seed)
        self.n_cols = n_cols
        self.noise_sigma = noise_sigma
        self.scene = Scene(seed)
        self.dirs = _ray_dirs(n_cols).reshape(-1, 3)
        rng = np.random.Generator(np.random.Philox(key=[self.seed, 0xE60]))
        self.v = rng.uniform(5.0, 15.0)            # m/s forward
        # rad/s.  The scene keeps a road |y| < 4 m free of buildings: a long drive (the benchmark's 64 scans) uses a small
        # yaw_rate_max so that it stays on it instead of driving through the boxes (inside one, most returns fall under
        # the 5 m minimum range and the scan thins out)
        self.yaw_rate = rng.uniform(-0.2, 0.2) * (yaw_rate_max / 0.2)
        # on_road: the vehicle stays on the ground plane — pitch, roll and height jitter around zero per scan instead of
        # accumulating (over a 64-scan drive the default random walk sinks or lifts the sensor by up to a metre)
        self.on_road = bool(on_road)
        self._poses: list[tuple[np.ndarray, np.ndarray]] = []
        self._plane: list[tuple[float, np.ndarray]] = []       # on_road: heading and position on the ground plane
        self._jit = rng

    def pose(self, k: int) -> tuple[np.ndarray, np.ndarray]:
        while len(self._poses) <= k:
            i = len(self._poses)
            if i == 0:
                R, t = np.eye(3), np.zeros(3)
                self._plane.append((0.0, np.zeros(3)))
            elif self.on_road:
                psi, p = self._plane[-1]
                dt = 0.1
                jr = np.random.Generator(np.random.Philox(key=[self.seed, 0x9000 + i]))
                pitch, roll = jr.normal(0, 2e-3), jr.normal(0, 2e-3)
                step = np.array([self.v * dt, jr.normal(0, 0.01), 0.0])
                p = p + _mv(_rot_z(psi), step)
                psi = psi + self.yaw_rate * dt
                self._plane.append((psi, p))
                R = _mm(_rot_z(psi), _rot_small(pitch, roll))
                t = p + np.array([0.0, 0.0, jr.normal(0, 0.005)])
            else:
                Rp, tp = self._poses[-1]
                dt = 0.1
                jr = np.random.Generator(np.random.Philox(key=[self.seed, 0x9000 + i]))
                dR = _mm(_rot_z(self.yaw_rate * dt), _rot_small(jr.normal(0, 2e-3), jr.normal(0, 2e-3)))
                dtv = np.array([self.v * dt, jr.normal(0, 0.01), jr.normal(0, 0.005)])
                R = _mm(Rp, dR)
                t = tp + _mv(Rp, dtv)
            self._poses.append((R, t))
        return self._poses[k]

    def relative_pose(self, k: int) -> tuple[np.ndarray, np.ndarray]:
        """Ground-truth (R, t) mapping scan-k coordinates into scan-(k-1) coordinates
        (the reference's q_last_curr / t_last_curr, laser_odometry.cpp:477-478)."""
        R0, t0 = self.pose(k - 1)
        R1, t1 = self.pose(k)
        return _mm(np.ascontiguousarray(R0.T), R1), _mv(np.ascontiguousarray(R0.T), t1 - t0)

    def scan(self, k: int) -> np.ndarray:
        """float32 (64*n_cols, 3) points in the sensor frame, ring-major; NaN = no return."""
        R, t = self.pose(k)
        dw = _mm(self.dirs, np.ascontiguousarray(R.T))
        rng_ = self.scene.raycast(t, dw)
        noise = np.random.Generator(np.random.Philox(key=[self.seed, 0x5CA0 + k])).normal(
            0.0, self.noise_sigma, rng_.shape[0])
        r = rng_ + noise
        valid = np.isfinite(rng_) & (rng_ >= R_MIN) & (rng_ <= R_MAX)
        pts = self.dirs * np.where(valid, r, 0.0)[:, None]
        pts = pts.astype(np.float32)
        pts[~valid] = np.nan
        return pts


def make_scans(seed: int, n_scans: int, n_cols: int = N_COLS) -> tuple[list[np.ndarray], ScanStream]:
    s = ScanStream(seed, n_cols=n_cols)
    return [s.scan(k) for k in range(n_scans)], s


# ------------------------------------------------------------------------------------------------ pre-built map (BASELINE configs[2])
def map_cubes(n_points: int, seed: int, x_range=(-100.0, 250.0), y_half=124.0):
    """A synthetic pre-built map for configs[2]: n_points surf points (one per 0.8 m voxel —
    the map's own resolution — jittered inside the voxel, on the ground plane and on stacked horizontal layers, so the
    map keeps its size under the reference's per-scan re-filter) plus 10 % as many corner points (vertical poles).  The map
    is a corridor along +x, the direction the synthetic vehicles drive: it fills the 5 x 5 x 3 cube window around the origin
    and reaches 125 m beyond it, so that the cubes entering the window when a vehicle crosses a cube border hold map points
    (they have to be indexed on entry).  Returns {(kind, cube_index): (n, 4) float32}."""
    rng = np.random.default_rng(seed)
    kx = np.arange(int(np.floor(x_range[0] / 0.8)), int(np.ceil(x_range[1] / 0.8)))
    ky = np.arange(-int(y_half / 0.8), int(y_half / 0.8))
    per_layer = kx.size * ky.size
    layers = max(1, int(np.ceil(n_points / per_layer)))
    pts = []
    for l in range(layers):
        gx, gy = np.meshgrid((kx + 0.5) * 0.8, (ky + 0.5) * 0.8)
        z = -1.73 + 7.0 * l
        p = np.c_[gx.ravel(), gy.ravel(), np.full(gx.size, z)] + rng.uniform(-0.3, 0.3, (gx.size, 3)) * [1, 1, 0.02]
        pts.append(p)
    surf = np.concatenate(pts)[:n_points].astype(np.float32)
    ncor = max(1000, n_points // 10)
    cx, cy = rng.uniform(x_range[0] + 4.0, x_range[1] - 4.0, ncor // 20), rng.uniform(-y_half + 4.0, y_half - 4.0, ncor // 20)
    corner = np.c_[np.repeat(cx, 20), np.repeat(cy, 20), np.tile(np.arange(20) * 0.4 - 1.7, ncor // 20)].astype(np.float32)
    out = {}
    for kind, cloud in ((0, corner), (1, surf)):
        ci = (np.floor((cloud[:, 0] + 25.0) / 50.0).astype(int) + 10) + 21 * (np.floor((cloud[:, 1] + 25.0) / 50.0).astype(int) + 10) \
            + 441 * (np.floor((cloud[:, 2] + 25.0) / 50.0).astype(int) + 5)
        for c in np.unique(ci):
            sel = cloud[ci == c]
            out[(kind, int(c))] = np.c_[sel, np.zeros(len(sel), np.float32)].astype(np.float32)
    return out


# ------------------------------------------------------------------------------------------------ visual odometry inputs
def kitti_like_calibration():
    """cam_T_velo (4x4), rect0_T_cam (4x4), P_rect0 (3x4): KITTI-like camera 0."""
    cam_T_velo = np.array([[0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, -0.08], [1.0, 0.0, 0.0, -0.27], [0.0, 0.0, 0.0, 1.0]], np.float32)
    rect0_T_cam = np.eye(4, dtype=np.float32)
    P_rect0 = np.array([[718.856, 0.0, 607.1928, 0.0], [0.0, 718.856, 185.2157, 0.0], [0.0, 0.0, 1.0, 0.0]], np.float32)
    return cam_T_velo, rect0_T_cam, P_rect0


def make_matches(stream: "ScanStream", k: int, n_matches: int = 800, pixel_sigma: float = 0.5, outlier_frac: float = 0.1,
                 seed: int = 0):
    """Matched keypoint pixels between camera frames k-1 and k: scene points seen by the LiDAR in frame k-1, projected
    into both images with Gaussian pixel noise and a fraction of gross outliers.  Returns (prev_uv, curr_uv) float32 (m, 2)
    and the ground-truth (R, t) taking frame k-1 camera coordinates to frame k camera coordinates."""
    rng = np.random.Generator(np.random.Philox(key=[stream.seed, 0x7A7C + 131 * k + seed]))
    cam_T_velo, _, P = kitti_like_calibration()
    Tcv = cam_T_velo.astype(np.float64)
    pts = stream.scan(k - 1).astype(np.float64)
    pts = pts[np.isfinite(pts[:, 0])]
    cam = pts @ Tcv[:3, :3].T + Tcv[:3, 3]
    K = P[:, :3].astype(np.float64)
    uvw = cam @ K.T
    uv0 = uvw[:, :2] / uvw[:, 2:3]
    ok = (cam[:, 2] > 2.0) & (uv0[:, 0] > 5) & (uv0[:, 0] < 1236) & (uv0[:, 1] > 5) & (uv0[:, 1] < 370)
    idx = rng.choice(np.nonzero(ok)[0], size=min(n_matches, int(ok.sum())), replace=False)
    # velo k-1 -> velo k is the inverse of relative_pose(k); conjugate into the camera frame
    R_lc, t_lc = stream.relative_pose(k)
    Rv, tv = R_lc.T, -R_lc.T @ t_lc
    Rc = Tcv[:3, :3] @ Rv @ Tcv[:3, :3].T
    tc = Tcv[:3, :3] @ tv + Tcv[:3, 3] - Rc @ Tcv[:3, 3]
    cam1 = cam[idx] @ Rc.T + tc
    uvw1 = cam1 @ K.T
    uv1 = uvw1[:, :2] / uvw1[:, 2:3]
    prev = uv0[idx] + rng.normal(0, pixel_sigma, (len(idx), 2))
    curr = uv1 + rng.normal(0, pixel_sigma, (len(idx), 2))
    bad = rng.random(len(idx)) < outlier_frac
    curr[bad] += rng.uniform(-60, 60, (int(bad.sum()), 2))
    return prev.astype(np.float32), curr.astype(np.float32), (Rc, tc)
