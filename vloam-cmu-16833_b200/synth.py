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
        """Range along each ray (inf for no hit). dirs: (M, 3) unit, world frame."""
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
            bq = dxy @ oc
            cq = oc @ oc - self.pole_r[p] ** 2
            disc = bq * bq - a * cq
            ok = disc > 0
            sq = np.sqrt(np.where(ok, disc, 0.0))
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (-bq - sq) / a
            z = o[2] + t * d[:, 2]
            hit = ok & (t > 0) & (z <= self.pole_top[p]) & (z >= GROUND_Z)
            best = np.where(hit & (t < best), t, best)
        return best


def _rot_z(yaw: float) -> np.ndarray:
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rot_small(roll: float, pitch: float) -> np.ndarray:
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    return ry @ rx


class ScanStream:
    """One synthetic LiDAR sequence: scene + ego-motion + per-scan noise.

    poses[k] = (R_k, t_k): sensor frame k -> world.  Scans are rendered from
    the pose at the scan's end (no intra-scan motion distortion; the reference
    runs with DISTORTION=false, laser_odometry.h:90).
    """

    def __init__(self, seed: int, n_cols: int = N_COLS, noise_sigma: float = 0.01):
        self.seed = int(seed)
        self.n_cols = n_cols
        self.noise_sigma = noise_sigma
        self.scene = Scene(seed)
        self.dirs = _ray_dirs(n_cols).reshape(-1, 3)
        rng = np.random.Generator(np.random.Philox(key=[self.seed, 0xE60]))
        self.v = rng.uniform(5.0, 15.0)            # m/s forward
        self.yaw_rate = rng.uniform(-0.2, 0.2)     # rad/s
        self._poses: list[tuple[np.ndarray, np.ndarray]] = []
        self._jit = rng

    def pose(self, k: int) -> tuple[np.ndarray, np.ndarray]:
        while len(self._poses) <= k:
            i = len(self._poses)
            if i == 0:
                R, t = np.eye(3), np.zeros(3)
            else:
                Rp, tp = self._poses[-1]
                dt = 0.1
                jr = np.random.Generator(np.random.Philox(key=[self.seed, 0x9000 + i]))
                dR = _rot_z(self.yaw_rate * dt) @ _rot_small(jr.normal(0, 2e-3), jr.normal(0, 2e-3))
                dtv = np.array([self.v * dt, jr.normal(0, 0.01), jr.normal(0, 0.005)])
                R = Rp @ dR
                t = tp + Rp @ dtv
            self._poses.append((R, t))
        return self._poses[k]

    def relative_pose(self, k: int) -> tuple[np.ndarray, np.ndarray]:
        """Ground-truth (R, t) mapping scan-k coordinates into scan-(k-1) coordinates
        (the reference's q_last_curr / t_last_curr, laser_odometry.cpp:477-478)."""
        R0, t0 = self.pose(k - 1)
        R1, t1 = self.pose(k)
        return R0.T @ R1, R0.T @ (t1 - t0)

    def scan(self, k: int) -> np.ndarray:
        """float32 (64*n_cols, 3) points in the sensor frame, ring-major; NaN = no return."""
        R, t = self.pose(k)
        dw = self.dirs @ R.T
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
