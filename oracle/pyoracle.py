"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle.so.

Only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() may import this module (the checker, never the thing
shipped).  PARITY UNPINNED: see oracle/types.hpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

c_fp = C.POINTER(C.c_float)
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the C++ restatement (g++ only).  Safe to call repeatedly."""
    # make is incremental: a no-op when the library is newer than every source, a rebuild when a header changed
    subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        sig = {
            "orc_sr_run": (vp, [c_fp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]),
            "orc_sr_free": (None, [vp]),
            "orc_sr_status": (C.c_int, [vp]),
            "orc_sr_count": (C.c_int, [vp, C.c_int]),
            "orc_sr_copy_cloud": (None, [vp, C.c_int, c_fp]),
            "orc_sr_array_len": (C.c_int, [vp, C.c_int]),
            "orc_sr_copy_array": (None, [vp, C.c_int, vp]),
            "orc_voxel_grid": (C.c_int, [c_fp, C.c_int, C.c_float, C.c_int, c_fp]),
            "orc_knn": (None, [c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, c_ip, c_fp]),
            "orc_sym_eig3": (None, [c_dp, c_dp, c_dp]),
            "orc_colpiv_qr_solve_5x3": (None, [c_dp, c_dp, c_dp]),
            "orc_factor_eval": (C.c_int, [C.c_int, c_dp, c_dp, c_dp, c_dp]),
            "orc_lo_create": (vp, [C.c_int, C.c_int]),
            "orc_lo_free": (None, [vp]),
            "orc_lo_set_iterations": (None, [vp, C.c_int, C.c_int]),
            "orc_lo_solve_sr": (None, [vp, vp, c_dp, c_dp]),
            "orc_lo_solve_clouds": (None, [vp] + [c_fp, C.c_int] * 5 + [c_dp, c_dp]),
            "orc_lo_get_state": (None, [vp, c_dp]),
            "orc_lo_set_motion": (None, [vp, c_dp, c_dp]),
            "orc_lo_set_pose": (None, [vp, c_dp, c_dp]),
            "orc_lo_set_skip": (None, [vp, C.c_int]),
            "orc_lo_set_distortion": (None, [vp, C.c_int]),
            "orc_lm_published_pose": (None, [vp, c_dp]),
            "orc_lo_trace_passes": (C.c_int, [vp]),
            "orc_lo_trace_sizes": (None, [vp, C.c_int, c_ip]),
            "orc_lo_trace_copy": (None, [vp, C.c_int, c_ip, c_ip, c_dp, c_dp, c_ip]),
            "orc_lo_last_count": (C.c_int, [vp, C.c_int]),
            "orc_lo_last_copy": (None, [vp, C.c_int, c_fp]),
            "orc_lm_create": (vp, [C.c_double, C.c_double]),
            "orc_lm_free": (None, [vp]),
            "orc_lm_reset": (None, [vp]),
            "orc_lm_set_iterations": (None, [vp, C.c_int, C.c_int]),
            "orc_lm_input_from_lo": (None, [vp, vp]),
            "orc_lm_input_clouds": (None, [vp, c_fp, C.c_int, c_fp, C.c_int, c_dp, c_dp]),
            "orc_lm_solve": (None, [vp]),
            "orc_lm_get_state": (None, [vp, c_dp]),
            "orc_lm_set_cube": (None, [vp, C.c_int, C.c_int, c_fp, C.c_int]),
            "orc_lm_cube_count": (C.c_int, [vp, C.c_int, C.c_int]),
            "orc_lm_cube_copy": (None, [vp, C.c_int, C.c_int, c_fp]),
            "orc_lm_map_points": (C.c_longlong, [vp, C.c_int]),
            "orc_lm_cloud_count": (C.c_int, [vp, C.c_int]),
            "orc_lm_cloud_copy": (None, [vp, C.c_int, c_fp]),
            "orc_lm_publish_registered": (None, [vp]),
            "orc_lm_trace_passes": (C.c_int, [vp]),
            "orc_lm_trace_sizes": (None, [vp, C.c_int, c_ip]),
            "orc_lm_trace_copy": (None, [vp, C.c_int, c_ip, c_ip, c_dp, c_dp, c_ip]),
            "orc_pipe_create": (vp, [C.c_int, C.c_double, C.c_double, C.c_double]),
            "orc_pipe_free": (None, [vp]),
            "orc_pipe_lo": (vp, [vp]),
            "orc_pipe_lm": (vp, [vp]),
            "orc_pipe_process": (C.c_int, [vp, c_fp, C.c_int, C.c_int, C.c_int]),
            "orc_pipe_timings": (None, [vp, c_dp, C.POINTER(C.c_longlong)]),
            "orc_vo_create": (vp, [c_fp, c_fp, c_fp, C.c_int]),
            "orc_vo_free": (None, [vp]),
            "orc_vo_reset": (None, [vp]),
            "orc_vo_slot": (C.c_int, [vp]),
            "orc_vo_set_max_iterations": (None, [vp, C.c_int]),
            "orc_vo_process_cloud": (None, [vp, c_fp, C.c_int, C.c_int]),
            "orc_vo_projected_count": (C.c_int, [vp, C.c_int]),
            "orc_vo_projected_copy": (None, [vp, C.c_int, c_fp]),
            "orc_vo_buckets_copy": (None, [vp, C.c_int, c_fp, c_fp, c_fp, c_ip]),
            "orc_vo_query_depth": (C.c_float, [vp, C.c_int, C.c_float, C.c_float]),
            "orc_vo_solve": (None, [vp, c_fp, c_fp, C.c_int, c_dp, c_dp, c_dp]),
            "orc_vo_factor_eval": (C.c_int, [C.c_int, c_dp, c_dp, c_dp, c_dp]),
            "orc_vo_trace": (C.c_int, [vp, c_dp, C.c_int]),
            "orc_vo_residuals": (C.c_int, [vp, c_ip, c_dp]),
            "orc_vo_to_lo_prior": (None, [c_dp, c_dp, c_dp, c_dp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(c_fp)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip)


SR_CLOUDS = ("laserCloud", "cornerPointsSharp", "cornerPointsLessSharp", "surfPointsFlat", "surfPointsLessFlat")
_SR_ARRAYS = (("curvature", np.float32), ("label", np.int32), ("picked", np.int32), ("scanStartInd", np.int32),
              ("scanEndInd", np.int32), ("sharpInd", np.int32), ("lessSharpInd", np.int32), ("flatInd", np.int32),
              ("ringLessFlatCount", np.int32))


class SRResult:
    """Owns a ScanRegistrationOutput; attributes are numpy copies."""

    def __init__(self, handle):
        L = lib()
        self._h = handle
        self.status = L.orc_sr_status(handle)
        for i, name in enumerate(SR_CLOUDS):
            n = L.orc_sr_count(handle, i)
            a = np.empty((n, 4), np.float32)
            L.orc_sr_copy_cloud(handle, i, _fp(a))
            setattr(self, name, a)
        for i, (name, dt) in enumerate(_SR_ARRAYS):
            n = L.orc_sr_array_len(handle, i)
            a = np.empty(n, dt)
            L.orc_sr_copy_array(handle, i, a.ctypes.data_as(C.c_void_p))
            setattr(self, name, a)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sr_free(self._h)
            self._h = None


def scan_registration(xyz, n_scans=64, minimum_range=5.0, literal_unstable=False) -> SRResult:
    xyz = _f32(xyz)
    assert xyz.ndim == 2 and xyz.shape[1] in (3, 4)
    h = lib().orc_sr_run(_fp(xyz), xyz.shape[0], xyz.shape[1], n_scans, float(minimum_range), int(literal_unstable))
    return SRResult(h)


def voxel_grid(xyzi, leaf, literal_unstable=False):
    xyzi = _f32(xyzi)
    out = np.empty_like(xyzi)
    n = lib().orc_voxel_grid(_fp(xyzi), xyzi.shape[0], float(leaf), int(literal_unstable), _fp(out))
    return out[:n].copy()


def knn(target_xyzi, query_xyzi, k, brute=False):
    t, q = _f32(target_xyzi), _f32(query_xyzi)
    idx = np.empty((q.shape[0], k), np.int32)
    d = np.empty((q.shape[0], k), np.float32)
    lib().orc_knn(_fp(t), t.shape[0], _fp(q), q.shape[0], k, int(brute), _ip(idx), _fp(d))
    return idx, d


def sym_eig3(A):
    A = np.ascontiguousarray(A, np.float64)
    ev = np.empty(3)
    evec = np.empty((3, 3))
    lib().orc_sym_eig3(_dp(A), _dp(ev), _dp(evec))
    return ev, evec


def colpiv_qr_solve_5x3(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.empty(3)
    lib().orc_colpiv_qr_solve_5x3(_dp(A), _dp(b), _dp(x))
    return x


def factor_eval(kind, pts, x):
    pts = np.ascontiguousarray(pts, np.float64).ravel()
    x = np.ascontiguousarray(x, np.float64)
    r = np.zeros(3)
    J = np.zeros(21)
    n = lib().orc_factor_eval(kind, _dp(pts), _dp(x), _dp(r), _dp(J))
    return r[:n].copy(), J[: n * 7].reshape(n, 7).copy()


def vo_factor_eval(kind, obs, x):
    obs = np.ascontiguousarray(obs, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    r = np.zeros(2)
    J = np.zeros(12)
    n = lib().orc_vo_factor_eval(kind, _dp(obs), _dp(x), _dp(r), _dp(J))
    return r[:n].copy(), J[: n * 6].reshape(n, 6).copy()


def vo_to_lo_prior(angles, t, velo_T_cam0):
    """(angles_0to1, t_0to1) -> velo_last_VOT_velo_curr as q(xyzw) t  (visual_odometry.cpp:426-430, vloam_tf.cpp:59-63)."""
    a = np.ascontiguousarray(angles, np.float64)
    tt = np.ascontiguousarray(t, np.float64)
    m = np.ascontiguousarray(velo_T_cam0, np.float64).reshape(16)
    out = np.zeros(7)
    lib().orc_vo_to_lo_prior(_dp(a), _dp(tt), _dp(m), _dp(out))
    return out


def _trace(L, h, prefix, npasses, widths):
    passes = []
    for p in range(npasses):
        sizes = np.zeros(3, np.int32)
        getattr(L, prefix + "_trace_sizes")(h, p, _ip(sizes))
        a = np.zeros(max(1, sizes[0] * widths[0]), np.int32)
        b = np.zeros(max(1, sizes[1] * widths[1]), np.int32)
        iters = np.zeros((max(1, sizes[2]), 7))
        para = np.zeros(7)
        term = C.c_int(0)
        getattr(L, prefix + "_trace_copy")(h, p, _ip(a), _ip(b), _dp(iters), _dp(para), C.byref(term))
        passes.append({
            "corner": a[: sizes[0] * widths[0]].reshape(-1, widths[0]).copy(),
            "plane": b[: sizes[1] * widths[1]].reshape(-1, widths[1]).copy(),
            "iterations": iters[: sizes[2]].copy(),
            "para": para,
            "termination": term.value,
        })
    return passes


class LaserOdometry:
    def __init__(self, detach_VO_LO=True, mapping_skip_frame=1, _borrowed=None):
        self._own = _borrowed is None
        self._h = lib().orc_lo_create(int(detach_VO_LO), mapping_skip_frame) if self._own else _borrowed

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            lib().orc_lo_free(self._h)
            self._h = None

    def set_iterations(self, passes, lm_iters):
        lib().orc_lo_set_iterations(self._h, passes, lm_iters)

    def solve(self, sr: SRResult, prior_q=None, prior_t=None):
        pq = np.ascontiguousarray(prior_q, np.float64) if prior_q is not None else None
        pt = np.ascontiguousarray(prior_t, np.float64) if prior_t is not None else None
        lib().orc_lo_solve_sr(self._h, sr._h, _dp(pq), _dp(pt))

    def solve_clouds(self, full, sharp, less_sharp, flat, less_flat, prior_q=None, prior_t=None):
        cl = [_f32(c).reshape(-1, 4) for c in (full, sharp, less_sharp, flat, less_flat)]
        args = []
        for c in cl:
            args += [_fp(c), c.shape[0]]
        pq = np.ascontiguousarray(prior_q, np.float64) if prior_q is not None else None
        pt = np.ascontiguousarray(prior_t, np.float64) if prior_t is not None else None
        lib().orc_lo_solve_clouds(self._h, *args, _dp(pq), _dp(pt))

    def set_motion(self, q, t):
        q = np.ascontiguousarray(q, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        lib().orc_lo_set_motion(self._h, _dp(q), _dp(t))

    def set_pose(self, q, t):
        """Overwrite q_w_curr / t_w_curr (the accumulated odometry pose)."""
        q = np.ascontiguousarray(q, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        lib().orc_lo_set_pose(self._h, _dp(q), _dp(t))

    def set_distortion(self, on=True):
        """laser_odometry.h:90 DISTORTION (a compile-time constant of the reference, false as shipped)."""
        lib().orc_lo_set_distortion(self._h, int(on))

    def set_mapping_skip_frame(self, n):
        lib().orc_lo_set_skip(self._h, int(n))

    @property
    def state(self):
        s = np.zeros(18)
        lib().orc_lo_get_state(self._h, _dp(s))
        return {"q_last_curr": s[0:4].copy(), "t_last_curr": s[4:7].copy(), "q_w_curr": s[7:11].copy(),
                "t_w_curr": s[11:14].copy(), "corner_correspondence": int(s[14]), "plane_correspondence": int(s[15]),
                "frameCount": int(s[16]), "systemInited": bool(s[17])}

    def trace(self):
        L = lib()
        return _trace(L, self._h, "orc_lo", L.orc_lo_trace_passes(self._h), (3, 4))

    def last_cloud(self, which):
        n = lib().orc_lo_last_count(self._h, which)
        a = np.empty((n, 4), np.float32)
        lib().orc_lo_last_copy(self._h, which, _fp(a))
        return a


class LaserMapping:
    def __init__(self, line_res=0.4, plane_res=0.8, _borrowed=None):
        self._own = _borrowed is None
        self._h = lib().orc_lm_create(line_res, plane_res) if self._own else _borrowed

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            lib().orc_lm_free(self._h)
            self._h = None

    def reset(self):
        lib().orc_lm_reset(self._h)

    def set_iterations(self, passes, lm_iters):
        lib().orc_lm_set_iterations(self._h, passes, lm_iters)

    def input_from_lo(self, lo: LaserOdometry):
        lib().orc_lm_input_from_lo(self._h, lo._h)

    def input_clouds(self, corner, surf, q_odom, t_odom):
        c, s = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
        q = np.ascontiguousarray(q_odom, np.float64)
        t = np.ascontiguousarray(t_odom, np.float64)
        lib().orc_lm_input_clouds(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0], _dp(q), _dp(t))

    def solve(self):
        lib().orc_lm_solve(self._h)

    @property
    def state(self):
        s = np.zeros(18)
        lib().orc_lm_get_state(self._h, _dp(s))
        return {"q_w_curr": s[0:4].copy(), "t_w_curr": s[4:7].copy(), "q_wmap_wodom": s[7:11].copy(),
                "t_wmap_wodom": s[11:14].copy(), "cen": s[14:17].astype(int), "validNum": int(s[17])}

    @property
    def published_pose(self):
        """(q_w xyzw, t_w, skipped): the pose LaserMapping::publish sends for the last input (high-frequency pose on a skipped frame)."""
        s = np.zeros(8)
        lib().orc_lm_published_pose(self._h, _dp(s))
        return s[0:4].copy(), s[4:7].copy(), bool(s[7])

    def set_cube(self, which, cube, xyzi):
        a = _f32(xyzi).reshape(-1, 4)
        lib().orc_lm_set_cube(self._h, which, cube, _fp(a), a.shape[0])

    def cube(self, which, cube):
        n = lib().orc_lm_cube_count(self._h, which, cube)
        a = np.empty((n, 4), np.float32)
        lib().orc_lm_cube_copy(self._h, which, cube, _fp(a))
        return a

    def cube_count(self, which, cube):
        return lib().orc_lm_cube_count(self._h, which, cube)

    def map_points(self, which):
        return int(lib().orc_lm_map_points(self._h, which))

    def cloud(self, which):
        n = lib().orc_lm_cloud_count(self._h, which)
        a = np.empty((n, 4), np.float32)
        lib().orc_lm_cloud_copy(self._h, which, _fp(a))
        return a

    def publish_registered(self):
        """LaserMapping::publish's /velodyne_cloud_registered (laser_mapping.cpp:797-805): moves laserCloudFullRes into the
        map frame in place (once per frame) and returns it."""
        lib().orc_lm_publish_registered(self._h)
        return self.cloud(4)

    def map_cloud(self):
        """/laser_cloud_map (laser_mapping.cpp:778-790): every cube's corner then surf points."""
        return self.cloud(5)

    def trace(self):
        L = lib()
        return _trace(L, self._h, "orc_lm", L.orc_lm_trace_passes(self._h), (1, 1))


class Pipeline:
    """scanRegistration -> laserOdometry [-> laserMapping] for one stream (CPU baseline)."""

    def __init__(self, n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, mapping_skip_frame=1):
        L = lib()
        self._h = L.orc_pipe_create(n_scans, float(minimum_range), line_res, plane_res)
        self.lo = LaserOdometry(_borrowed=L.orc_pipe_lo(self._h))
        self.lm = LaserMapping(_borrowed=L.orc_pipe_lm(self._h))
        self.lo.set_mapping_skip_frame(mapping_skip_frame)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_pipe_free(self._h)
            self._h = None

    def process(self, xyz, do_mapping=False):
        xyz = _f32(xyz)
        return lib().orc_pipe_process(self._h, _fp(xyz), xyz.shape[0], xyz.shape[1], int(do_mapping))

    def timings(self):
        ms = np.zeros(3)
        n = C.c_longlong(0)
        lib().orc_pipe_timings(self._h, _dp(ms), C.byref(n))
        return {"sr_ms": ms[0], "lo_ms": ms[1], "lm_ms": ms[2], "scans": n.value}


class VisualOdometry:
    def __init__(self, cam_T_velo, rect0_T_cam, P_rect0, remove_VO_outlier=100):
        a, b, c = _f32(cam_T_velo).ravel(), _f32(rect0_T_cam).ravel(), _f32(P_rect0).ravel()
        assert a.size == 16 and b.size == 16 and c.size == 12
        self._h = lib().orc_vo_create(_fp(a), _fp(b), _fp(c), remove_VO_outlier)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_vo_free(self._h)
            self._h = None

    def reset(self):
        lib().orc_vo_reset(self._h)

    @property
    def slot(self):
        return lib().orc_vo_slot(self._h)

    def set_max_iterations(self, n):
        lib().orc_vo_set_max_iterations(self._h, n)

    def process_cloud(self, xyz):
        xyz = _f32(xyz)
        lib().orc_vo_process_cloud(self._h, _fp(xyz), xyz.shape[0], xyz.shape[1])

    def projected(self, slot):
        n = lib().orc_vo_projected_count(self._h, slot)
        a = np.empty((n, 3), np.float32)
        lib().orc_vo_projected_copy(self._h, slot, _fp(a))
        return a

    def buckets(self, slot):
        nb = 249 * 75
        bx, by, bd = (np.empty(nb, np.float32) for _ in range(3))
        bc = np.empty(nb, np.int32)
        lib().orc_vo_buckets_copy(self._h, slot, _fp(bx), _fp(by), _fp(bd), _ip(bc))
        return bx.reshape(249, 75), by.reshape(249, 75), bd.reshape(249, 75), bc.reshape(249, 75)

    def query_depth(self, slot, x, y):
        return float(lib().orc_vo_query_depth(self._h, slot, float(x), float(y)))

    def residuals(self, m):
        t = np.zeros(m, np.int32)
        o = np.zeros((m, 5))
        n = lib().orc_vo_residuals(self._h, _ip(t), _dp(o))
        return t[:n], o[:n]

    def trace(self, max_records=128):
        it = np.zeros((max_records, 7))
        n = lib().orc_vo_trace(self._h, _dp(it), max_records)
        return it[: min(n, max_records)].copy()

    def solve(self, prev_uv, curr_uv, init_aa=None, init_t=None):
        p, c = _f32(prev_uv).reshape(-1, 2), _f32(curr_uv).reshape(-1, 2)
        ia = np.ascontiguousarray(init_aa, np.float64) if init_aa is not None else None
        it = np.ascontiguousarray(init_t, np.float64) if init_t is not None else None
        out = np.zeros(16)
        lib().orc_vo_solve(self._h, _fp(p), _fp(c), p.shape[0], _dp(ia), _dp(it), _dp(out))
        return {"angles_0to1": out[0:3].copy(), "t_0to1": out[3:6].copy(), "counter32": int(out[6]),
                "counter22": int(out[7]), "termination": int(out[8]), "iterations": int(out[9]), "final_cost": out[10]}
