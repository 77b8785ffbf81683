"""vloam_b200 — host-side mirror of the reference's LiDAR / visual odometry class API over the C-ABI.

The product is `lib/libvloam_b200.so` (CUDA, sm_100a; include/vloam_b200.h).  This module is the thin
Python harness around it used by tests/ and bench.py: it loads the shared library with ctypes and exposes
classes with the reference's method names

    LidarOdometryMapping.init / reset / scanRegistrationIO / laserOdometryIO / laserMappingIO
        (reference include/lidar_odometry_mapping/lidar_odometry_mapping.h:45-86)
    VisualOdometry.init / reset / setUpPointCloud / processPointCloud / solveNlsAll
        (reference include/visual_odometry/visual_odometry.h:40-121)

There is NO CPU fallback: if the library is missing or no CUDA device is present the calls raise.
The directory name contains a '-', so import it through the repo-root shim `vloam_b200.py`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvloam_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vloam_b200.h")

VLOAM_OK = 0
CLOUD_FULL, CLOUD_SHARP, CLOUD_LESS_SHARP, CLOUD_FLAT, CLOUD_LESS_FLAT, CLOUD_CORNER_LAST, CLOUD_SURF_LAST, \
    CLOUD_CORNER_STACK, CLOUD_SURF_STACK, CLOUD_CORNER_MAP, CLOUD_SURF_MAP, CLOUD_MAP, CLOUD_FULL_REGISTERED = range(13)
STREAM_EMPTY, STREAM_RING_OVERFLOW, STREAM_VOXEL_OVERFLOW, STREAM_CAPACITY = 1, 2, 4, 8
LM_WORKLIST_OVERFLOW, LM_SCRATCH_OVERFLOW, LM_CORNER_MAP_FULL, LM_SURF_MAP_FULL = 1, 2, 4, 8
MAX_SHARP, MAX_FLAT = 768, 1536

c_fp = C.POINTER(C.c_float)
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class VloamError(RuntimeError):
    pass


class LidarParams(C.Structure):
    _fields_ = [("batch", C.c_int), ("max_points", C.c_int), ("scan_line", C.c_int), ("minimum_range", C.c_double),
                ("mapping_line_resolution", C.c_double), ("mapping_plane_resolution", C.c_double),
                ("mapping_skip_frame", C.c_int), ("detach_VO_LO", C.c_int), ("lo_outer_passes", C.c_int),
                ("lo_max_iterations", C.c_int), ("lm_outer_passes", C.c_int), ("lm_max_iterations", C.c_int),
                ("map_capacity_points", C.c_int), ("debug_keep_submap", C.c_int), ("solver_mode", C.c_int), ("distortion", C.c_int)]


def build(verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run([os.path.join(_HERE, "build.sh")], capture_output=True, text=True)
    if out.returncode != 0:
        raise VloamError("build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VloamError(f"{LIB_PATH} missing: run __graft_entry__.build() (nvcc) first; there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        pp = C.POINTER(vp)
        sig = {
            "vloam_ctx_create": [C.c_int, pp], "vloam_ctx_destroy": [vp], "vloam_ctx_set_stream": [vp, vp],
            "vloam_ctx_synchronize": [vp], "vloam_ctx_enable_timing": [vp, C.c_int],
            "vloam_ctx_get_kernel_timings": [vp, c_dp, C.POINTER(C.c_longlong), C.c_int],
            "vloam_lidar_params_default": [C.POINTER(LidarParams)],
            "vloam_lidar_create": [vp, C.POINTER(LidarParams), pp], "vloam_lidar_destroy": [vp], "vloam_lidar_reset": [vp],
            "vloam_scan_registration": [vp, vp, vp, C.c_int, C.c_size_t],
            "vloam_scan_registration_device": [vp, vp, vp, C.c_int, C.c_size_t],
            "vloam_scan_registration_ptrs": [vp, vp, vp, C.c_int],
            "vloam_lidar_process": [vp, vp, vp, C.c_int, C.c_size_t, vp, C.c_int],
            "vloam_lidar_process_device": [vp, vp, vp, C.c_int, C.c_size_t, vp, C.c_int],
            "vloam_lidar_process_ptrs": [vp, vp, vp, C.c_int, vp, C.c_int],
            "vloam_get_input_device": [vp, pp, pp, c_ip, C.POINTER(C.c_size_t)], "vloam_input_consumed": [vp],
            "vloam_shard_buffer": [vp, pp, C.POINTER(C.c_size_t)], "vloam_shard_ipc_handle": [vp, C.c_char_p],
            "vloam_shard_open_ipc": [vp, C.c_int, C.c_int, C.c_char_p], "vloam_shard_enable": [vp, C.c_int, C.c_int, pp],
            "vloam_shard_disable": [vp], "vloam_shard_status": [vp, c_ip],
            "vloam_shard_nccl_unique_id": [C.c_char_p], "vloam_shard_nccl_init": [vp, C.c_int, C.c_int, C.c_char_p],
            "vloam_shard_nccl_destroy": [vp],
            "vloam_get_stream_status": [vp, c_ip], "vloam_get_feature_counts": [vp, c_ip],
            "vloam_get_cloud": [vp, C.c_int, C.c_int, c_fp, C.c_int, c_ip],
            "vloam_get_curvature": [vp, C.c_int, c_fp, C.c_int, c_ip],
            "vloam_get_labels": [vp, C.c_int, vp, C.c_int, c_ip],
            "vloam_get_feature_indices": [vp, C.c_int, C.c_int, c_ip, C.c_int, c_ip],
            "vloam_laser_odometry": [vp, vp, vp, vp], "vloam_laser_odometry_async": [vp, vp],
            "vloam_get_lo_pose": [vp, vp, vp], "vloam_get_lo_pose_prev": [vp, vp, vp], "vloam_set_lo_motion": [vp, c_dp],
            "vloam_set_lo_pose": [vp, c_dp], "vloam_get_lm_status": [vp, c_ip],
            "vloam_get_lm_queries": [vp, C.c_int, C.c_int, C.c_int, c_ip, C.c_int, c_ip],
            "vloam_get_lm_counters": [vp, C.POINTER(C.c_longlong)], "vloam_lidar_set_debug_stats": [vp, C.c_int],
            "vloam_get_lo_trace": [vp, C.c_int, C.c_int, c_ip, c_dp, c_ip, c_dp],
            "vloam_laser_mapping": [vp, vp], "vloam_get_lm_pose": [vp, c_dp],
            "vloam_map_set_cube": [vp, C.c_int, C.c_int, C.c_int, c_fp, C.c_int],
            "vloam_map_get_cube": [vp, C.c_int, C.c_int, C.c_int, c_fp, C.c_int, c_ip],
            "vloam_get_lm_info": [vp, c_ip], "vloam_get_map_stats": [vp, c_ip], "vloam_get_lm_trace": [vp, C.c_int, C.c_int, c_dp, c_ip, c_dp],
            "vloam_vo_create": [vp, C.c_int, C.c_int, C.c_int, pp], "vloam_vo_destroy": [vp],
            "vloam_vo_set_calibration": [vp, c_fp, c_fp, c_fp], "vloam_vo_reset": [vp],
            "vloam_vo_process_cloud": [vp, vp, vp, C.c_int, C.c_size_t],
            "vloam_vo_process_cloud_device": [vp, vp, vp, C.c_int, C.c_size_t],
            "vloam_vo_query_depth": [vp, C.c_int, C.c_int, c_fp, C.c_int, c_fp],
            "vloam_vo_get_buckets": [vp, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_ip],
            "vloam_vo_solve": [vp, vp, vp, vp, vp, C.c_int, C.c_int, c_dp],
            "vloam_vo_solve_device_async": [vp, vp, vp, vp, vp, C.c_int, C.c_int], "vloam_vo_get_result": [vp, c_dp],
            "vloam_vo_export_lo_prior": [vp, c_dp, vp],
            "vloam_vo_get_trace": [vp, C.c_int, c_dp, c_ip, c_dp],
            "vloam_vo_match_descriptors": [vp, vp, vp, vp, vp, vp, vp, C.c_double, vp, vp], "vloam_vo_get_knn": [vp, vp],
            "vloam_vo_get_match_buffers": [vp, pp, pp, pp], "vloam_vo_get_match_uv": [vp, vp, vp],
            "vloam_vo_detect_corners": [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp],
            "vloam_vo_get_corner_response": [vp, C.c_int, vp, C.c_size_t], "vloam_vo_get_corner_buffers": [vp, pp, pp],
            "vloam_vo_get_residuals": [vp, C.c_int, c_ip, c_dp],
            "vloam_vo_describe_orb": [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp],
            "vloam_vo_process_image": [vp, vp, C.c_int, C.c_int, vp, vp],
            "vloam_vo_get_frame_features": [vp, C.c_int, vp, vp, vp], "vloam_vo_get_matches": [vp, vp, vp],
            "vloam_vo_get_detect_status": [vp, c_ip],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        L.vloam_last_error.restype = C.c_char_p
        L.vloam_last_error.argtypes = [vp]
        L.vloam_ctx_kernel_count.restype = C.c_int
        L.vloam_ctx_kernel_count.argtypes = []
        L.vloam_ctx_kernel_name.restype = C.c_char_p
        L.vloam_ctx_kernel_name.argtypes = [C.c_int]
        L.vloam_ctx_launch_count.restype = C.c_longlong
        L.vloam_ctx_launch_count.argtypes = [vp]
        _lib = L
    return _lib


def _prefer_bundled_nccl() -> None:
    """The library binds NCCL at run time.  When the interpreter has a pip-installed NCCL (the copy PyTorch loads), name it in
    VLOAM_NCCL_LIB so both end up on the same libnccl.so.2 whichever is loaded first; a system copy loaded first would
    otherwise satisfy PyTorch's dependency by soname and may be too old for it."""
    if os.environ.get("VLOAM_NCCL_LIB"):
        return
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (spec.submodule_search_locations if spec else []):
        cand = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["VLOAM_NCCL_LIB"] = cand
            return


def shard_nccl_unique_id() -> bytes:
    """A fresh 128-byte ncclUniqueId (rank 0 creates it, the caller broadcasts it)."""
    _prefer_bundled_nccl()
    buf = C.create_string_buffer(128)
    rc = lib().vloam_shard_nccl_unique_id(buf)
    if rc != VLOAM_OK:
        raise VloamError(f"vloam_shard_nccl_unique_id failed with {rc}: no libnccl in the process?")
    return buf.raw


def exported_symbols_in_header() -> list[str]:
    """Every function name include/vloam_b200.h declares (used by the CPU-side load test)."""
    import re
    txt = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"\b(vloam_[a-z0-9_]+)\s*\(", txt)))


def _ptr(a):
    """ctypes void* of a numpy array, a torch tensor (host or device) or an int address."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Context:
    def __init__(self, device: int = 0, cuda_stream: int | None = None):
        self._h = C.c_void_p()
        rc = lib().vloam_ctx_create(device, C.byref(self._h))
        if rc != VLOAM_OK:
            raise VloamError(f"vloam_ctx_create(device={device}) failed with {rc}: no CUDA device? (there is no CPU fallback)")
        if cuda_stream is not None:
            self.set_stream(cuda_stream)

    def check(self, rc):
        if rc != VLOAM_OK:
            raise VloamError(f"vloam error {rc}: {lib().vloam_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream: int | None):
        self.check(lib().vloam_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def synchronize(self):
        self.check(lib().vloam_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().vloam_ctx_launch_count(self._h))

    def enable_timing(self, on: bool = True):
        self.check(lib().vloam_ctx_enable_timing(self._h, int(on)))

    def kernel_timings(self, reset: bool = True) -> dict:
        """{kernel name: (total ms, launches)} measured with CUDA events on the launching stream."""
        n = lib().vloam_ctx_kernel_count()
        ms = np.zeros(n)
        cnt = np.zeros(n, np.int64)
        self.check(lib().vloam_ctx_get_kernel_timings(self._h, ms.ctypes.data_as(c_dp),
                                                      cnt.ctypes.data_as(C.POINTER(C.c_longlong)), int(reset)))
        return {lib().vloam_ctx_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}

    def close(self):
        if self._h:
            lib().vloam_ctx_destroy(self._h)
            self._h = C.c_void_p()


def default_lidar_params(**kw) -> LidarParams:
    p = LidarParams()
    lib().vloam_lidar_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class LidarOdometryMapping:
    """Mirror of vloam::LidarOdometryMapping for `batch` streams driven in lock-step.

    Poses come back as dicts of (batch, k) float64 arrays; clouds as (n, 4) float32 arrays.
    """

    def __init__(self, ctx: Context | None = None, **params):
        self.ctx = ctx or Context()
        self._own_ctx = ctx is None
        self.params = default_lidar_params(**params)
        self._h = C.c_void_p()
        self.batch = self.params.batch
        self.last_pose = None
        self.last_map_pose = None
        self.init()

    # -- lidar_odometry_mapping.cpp:40-63
    def init(self, vloam_tf=None):
        if self._h:
            lib().vloam_lidar_destroy(self._h)
            self._h = C.c_void_p()
        self.ctx.check(lib().vloam_lidar_create(self.ctx._h, C.byref(self.params), C.byref(self._h)))
        self.vloam_tf = vloam_tf

    # -- lidar_odometry_mapping.cpp:65-71
    def reset(self):
        self.ctx.check(lib().vloam_lidar_reset(self._h))

    # -- lidar_odometry_mapping.cpp:73-94
    def scanRegistrationIO(self, laserCloudIn, n_points=None):
        """laserCloudIn: (batch, n, s) or (n, s) float32 host array (numpy / pinned torch), 3 <= s <= 16 floats per point
        starting with x, y, z (3 = packed, 4 = pcl::PointXYZ / KITTI .bin, more = a PointCloud2 payload); NaN = no return."""
        a = laserCloudIn
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, dtype=np.float32)
        shape = tuple(a.shape)
        if len(shape) == 2:
            shape = (1,) + shape
        assert shape[0] == self.batch and 3 <= shape[2] <= 16, shape
        if n_points is None:
            n_points = np.full(self.batch, shape[1], np.int32)
        n_points = np.ascontiguousarray(n_points, np.int32)
        self._keep = (a, n_points)  # keep host buffers alive until the async upload has been consumed
        self.ctx.check(lib().vloam_scan_registration(self._h, _ptr(a), _ptr(n_points), shape[2], shape[1]))

    def scanRegistrationPtrs(self, ptrs, n_points, stride: int, keep=None):
        """One host buffer per stream: ptrs = (batch,) uint64 array of host addresses (pinned memory uploads asynchronously);
        `keep` = whatever owns those buffers (kept alive until the next call)."""
        p = np.ascontiguousarray(ptrs, np.uint64)
        n = np.ascontiguousarray(n_points, np.int32)
        assert p.shape == (self.batch,) and n.shape == (self.batch,)
        self._keep = (p, n, keep)
        self.ctx.check(lib().vloam_scan_registration_ptrs(self._h, _ptr(p), _ptr(n), stride))

    def process(self, laserCloudIn, n_points=None, prior_dev=None, use_graph=True):
        """reset + scanRegistrationIO + laserOdometryIO + laserMappingIO in one call (asynchronous; read the poses with
        lo_pose() / lm_pose()).  use_graph: replay the frame's launch sequence as one CUDA graph."""
        a = laserCloudIn
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, dtype=np.float32)
        shape = tuple(a.shape)
        if len(shape) == 2:
            shape = (1,) + shape
        assert shape[0] == self.batch and 3 <= shape[2] <= 16, shape
        if n_points is None:
            n_points = np.full(self.batch, shape[1], np.int32)
        n_points = np.ascontiguousarray(n_points, np.int32)
        self._keep = (a, n_points)
        self.ctx.check(lib().vloam_lidar_process(self._h, _ptr(a), _ptr(n_points), shape[2], shape[1], _ptr(prior_dev), int(use_graph)))

    def processPtrs(self, ptrs, n_points, stride: int, prior_dev=None, use_graph=True, keep=None):
        """process() with one host buffer per stream (see scanRegistrationPtrs)."""
        p = np.ascontiguousarray(ptrs, np.uint64)
        n = np.ascontiguousarray(n_points, np.int32)
        self._keep = (p, n, keep)
        self.ctx.check(lib().vloam_lidar_process_ptrs(self._h, _ptr(p), _ptr(n), stride, _ptr(prior_dev), int(use_graph)))

    def processDevice(self, xyz_dev, n_points_dev, stride: int, slab_points: int, prior_dev=None, use_graph=True):
        self._keep = (xyz_dev, n_points_dev)
        self.ctx.check(lib().vloam_lidar_process_device(self._h, _ptr(xyz_dev), _ptr(n_points_dev), stride, slab_points, _ptr(prior_dev),
                                                        int(use_graph)))

    def lm_pose(self):
        pose = np.zeros((self.batch, 14))
        self.ctx.check(lib().vloam_get_lm_pose(self._h, pose.ctypes.data_as(c_dp)))
        return {"q_w_curr": pose[:, 0:4].copy(), "t_w_curr": pose[:, 4:7].copy(), "q_wmap_wodom": pose[:, 7:11].copy(),
                "t_wmap_wodom": pose[:, 11:14].copy()}

    def scanRegistrationDevice(self, xyz_dev, n_points_dev, stride: int, slab_points: int):
        """Scans already resident in HBM (torch CUDA tensors or raw device addresses)."""
        self._keep = (xyz_dev, n_points_dev)
        self.ctx.check(lib().vloam_scan_registration_device(self._h, _ptr(xyz_dev), _ptr(n_points_dev), stride, slab_points))

    # -- lidar_odometry_mapping.cpp:96-123
    def input_device(self):
        """(xyz address, n_points address, stride, slab_points) of the scan scanRegistrationIO uploaded last."""
        xyz, n = C.c_void_p(), C.c_void_p()
        stride, slab = C.c_int(), C.c_size_t()
        self.ctx.check(lib().vloam_get_input_device(self._h, C.byref(xyz), C.byref(n), C.byref(stride), C.byref(slab)))
        return xyz.value, n.value, stride.value, slab.value

    def input_consumed(self):
        self.ctx.check(lib().vloam_input_consumed(self._h))

    # -- point-sharded solve (multi-GPU): see include/vloam_b200.h "point-sharded solve"
    def shard_buffer(self) -> int:
        p, n = C.c_void_p(), C.c_size_t()
        self.ctx.check(lib().vloam_shard_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value

    def shard_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.ctx.check(lib().vloam_shard_ipc_handle(self._h, buf))
        return buf.raw

    def shard_open_ipc(self, rank: int, world: int, handles: bytes):
        assert len(handles) == 64 * world
        self.ctx.check(lib().vloam_shard_open_ipc(self._h, rank, world, handles))

    def shard_enable(self, rank: int, world: int, peer_ptrs):
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        self.ctx.check(lib().vloam_shard_enable(self._h, rank, world, arr))

    def shard_nccl_init(self, rank: int, world: int, unique_id: bytes):
        """Point-sharded streams with the exchange done by ncclAllReduce (collective over the group)."""
        assert len(unique_id) == 128
        _prefer_bundled_nccl()
        self.ctx.check(lib().vloam_shard_nccl_init(self._h, rank, world, unique_id))

    def shard_nccl_destroy(self):
        self.ctx.check(lib().vloam_shard_nccl_destroy(self._h))

    def shard_disable(self):
        self.ctx.check(lib().vloam_shard_disable(self._h))

    def shard_status(self) -> int:
        e = C.c_int(0)
        self.ctx.check(lib().vloam_shard_status(self._h, C.byref(e)))
        return e.value

    def laserOdometryIO(self, prior=None, fetch=True):
        if not fetch:
            self.ctx.check(lib().vloam_laser_odometry_async(self._h, _ptr(prior)))
            return None
        pose = np.zeros((self.batch, 14))
        corr = np.zeros((self.batch, 2), np.int32)
        pr = np.ascontiguousarray(prior, np.float64).reshape(self.batch, 7) if prior is not None else None
        self.ctx.check(lib().vloam_laser_odometry(self._h, _ptr(pr), _ptr(pose), _ptr(corr)))
        self.last_pose = self._split_lo(pose, corr)
        return self.last_pose

    def lo_pose(self, prev: bool = False):
        """Pose of the current scan (or, prev=True, of the previous scan while the current one is still in flight)."""
        pose = np.zeros((self.batch, 14))
        corr = np.zeros((self.batch, 2), np.int32)
        fn = lib().vloam_get_lo_pose_prev if prev else lib().vloam_get_lo_pose
        self.ctx.check(fn(self._h, _ptr(pose), _ptr(corr)))
        self.last_pose = self._split_lo(pose, corr)
        return self.last_pose

    @staticmethod
    def _split_lo(pose, corr):
        return {"q_last_curr": pose[:, 0:4].copy(), "t_last_curr": pose[:, 4:7].copy(), "q_w_curr": pose[:, 7:11].copy(),
                "t_w_curr": pose[:, 11:14].copy(), "corner_correspondence": corr[:, 0].copy(),
                "plane_correspondence": corr[:, 1].copy()}

    def set_lo_motion(self, motion):
        m = np.ascontiguousarray(motion, np.float64).reshape(self.batch, 7)
        self.ctx.check(lib().vloam_set_lo_motion(self._h, m.ctypes.data_as(c_dp)))

    def set_lo_pose(self, pose):
        """Overwrite q_w_curr / t_w_curr of laser odometry, (batch, 7) = q(xyzw) t (checkpoint / resume)."""
        m = np.ascontiguousarray(pose, np.float64).reshape(self.batch, 7)
        self.ctx.check(lib().vloam_set_lo_pose(self._h, m.ctypes.data_as(c_dp)))

    # -- lidar_odometry_mapping.cpp:125-154
    def laserMappingIO(self, fetch=True):
        if not fetch:
            self.ctx.check(lib().vloam_laser_mapping(self._h, None))
            return None
        pose = np.zeros((self.batch, 14))
        self.ctx.check(lib().vloam_laser_mapping(self._h, _ptr(pose)))
        self.last_map_pose = {"q_w_curr": pose[:, 0:4].copy(), "t_w_curr": pose[:, 4:7].copy(),
                              "q_wmap_wodom": pose[:, 7:11].copy(), "t_wmap_wodom": pose[:, 11:14].copy()}
        return self.last_map_pose

    # -- read-out helpers (ScanRegistration::output / parity views)
    def stream_status(self):
        s = np.zeros(self.batch, np.int32)
        self.ctx.check(lib().vloam_get_stream_status(self._h, s.ctypes.data_as(c_ip)))
        return s

    def feature_counts(self):
        c = np.zeros((self.batch, 5), np.int32)
        self.ctx.check(lib().vloam_get_feature_counts(self._h, c.ctypes.data_as(c_ip)))
        return c

    def cloud(self, which: int, stream: int = 0):
        n = C.c_int(0)
        self.ctx.check(lib().vloam_get_cloud(self._h, stream, which, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), np.float32)
        if n.value:
            self.ctx.check(lib().vloam_get_cloud(self._h, stream, which, out.ctypes.data_as(c_fp), n.value, C.byref(n)))
        return out

    def curvature(self, stream: int = 0):
        n = C.c_int(0)
        self.ctx.check(lib().vloam_get_curvature(self._h, stream, None, 0, C.byref(n)))
        out = np.empty(n.value, np.float32)
        if n.value:
            self.ctx.check(lib().vloam_get_curvature(self._h, stream, out.ctypes.data_as(c_fp), n.value, C.byref(n)))
        return out

    def labels(self, stream: int = 0):
        n = C.c_int(0)
        self.ctx.check(lib().vloam_get_labels(self._h, stream, None, 0, C.byref(n)))
        out = np.empty(n.value, np.int8)
        if n.value:
            self.ctx.check(lib().vloam_get_labels(self._h, stream, C.c_void_p(out.ctypes.data), n.value, C.byref(n)))
        return out

    def feature_indices(self, which: int, stream: int = 0):
        n = C.c_int(0)
        self.ctx.check(lib().vloam_get_feature_indices(self._h, stream, which, None, 0, C.byref(n)))
        out = np.empty(n.value, np.int32)
        if n.value:
            self.ctx.check(lib().vloam_get_feature_indices(self._h, stream, which, out.ctypes.data_as(c_ip), n.value, C.byref(n)))
        return out

    def lo_trace(self, pass_: int, stream: int = 0):
        corr = np.zeros((MAX_SHARP + MAX_FLAT, 4), np.int32)
        rec = np.zeros((8, 7))
        info = np.zeros(4, np.int32)
        para = np.zeros(7)
        self.ctx.check(lib().vloam_get_lo_trace(self._h, stream, pass_, corr.ctypes.data_as(c_ip), rec.ctypes.data_as(c_dp),
                                                info.ctypes.data_as(c_ip), para.ctypes.data_as(c_dp)))
        cq = corr[:MAX_SHARP]
        pq = corr[MAX_SHARP:]
        ci = np.nonzero(cq[:, 3])[0]
        pi = np.nonzero(pq[:, 3])[0]
        return {"corner": np.column_stack([ci, cq[ci, 0], cq[ci, 1]]).astype(np.int32),
                "plane": np.column_stack([pi, pq[pi, 0], pq[pi, 1], pq[pi, 2]]).astype(np.int32),
                "iterations": rec[: min(int(info[0]), 8)].copy(), "n_records": int(info[0]), "termination": int(info[1]),
                "n_corner": int(info[2]), "n_plane": int(info[3]), "para": para}

    def lm_info(self):
        info = np.zeros((self.batch, 8), np.int32)
        self.ctx.check(lib().vloam_get_lm_info(self._h, info.ctypes.data_as(c_ip)))
        return info

    def map_stats(self):
        """(batch, 2, 10) int32 per stream and kind: points, pool high-water mark, pool index, non-empty cubes, fixed-point
        cubes, cubes rewritten by the last scan, re-packs so far, slab capacity in use, points in the rewritten cubes,
        column-index table slots in use."""
        st = np.zeros((self.batch, 2, 10), np.int32)
        self.ctx.check(lib().vloam_get_map_stats(self._h, st.ctypes.data_as(c_ip)))
        return st

    def lm_trace(self, pass_: int, stream: int = 0):
        rec = np.zeros((8, 7))
        info = np.zeros(4, np.int32)
        para = np.zeros(7)
        self.ctx.check(lib().vloam_get_lm_trace(self._h, stream, pass_, rec.ctypes.data_as(c_dp), info.ctypes.data_as(c_ip),
                                                para.ctypes.data_as(c_dp)))
        return {"iterations": rec[: min(int(info[0]), 8)].copy(), "n_records": int(info[0]), "termination": int(info[1]),
                "n_corner": int(info[2]), "n_plane": int(info[3]), "para": para}

    def lm_status(self):
        """(batch, 2) int32: LM_* bits of the last mapped scan, OR of the bits of every scan so far."""
        st = np.zeros((self.batch, 2), np.int32)
        self.ctx.check(lib().vloam_get_lm_status(self._h, st.ctypes.data_as(c_ip)))
        return st

    LM_COUNTERS = ("scans", "solved", "lm_iterations_pass0", "lm_iterations_pass1", "cubes_indexed_corner", "cubes_indexed_surf",
                   "cubes_rewritten_corner", "cubes_rewritten_surf", "cubes_merged", "cubes_merged_in_place", "cubes_filtered",
                   "cubes_appended", "voxels_inserted", "repacks", "queries", "factors_last_pass", "knn_queries", "knn_candidates")

    def lm_counters(self):
        """(batch, 18) int64 cumulative laser-mapping work counters (names: LM_COUNTERS)."""
        c = np.zeros((self.batch, 18), np.int64)
        self.ctx.check(lib().vloam_get_lm_counters(self._h, c.ctypes.data_as(C.POINTER(C.c_longlong))))
        return c

    def set_debug_stats(self, on: bool = True):
        self.ctx.check(lib().vloam_lidar_set_debug_stats(self._h, int(on)))

    def lm_queries(self, pass_: int, kind: int, stream: int = 0):
        """Indices into the down-sampled corner (kind 0) / surf (kind 1) stack that produced a factor in outer pass `pass_`."""
        n = C.c_int(0)
        self.ctx.check(lib().vloam_get_lm_queries(self._h, stream, pass_, kind, None, 0, C.byref(n)))
        out = np.empty(n.value, np.int32)
        if n.value:
            self.ctx.check(lib().vloam_get_lm_queries(self._h, stream, pass_, kind, out.ctypes.data_as(c_ip), n.value, C.byref(n)))
        return out

    def map_set_cube(self, kind: int, cube: int, xyzi, stream: int = 0):
        a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        self.ctx.check(lib().vloam_map_set_cube(self._h, stream, kind, cube, a.ctypes.data_as(c_fp), a.shape[0]))

    def map_get_cube(self, kind: int, cube: int, stream: int = 0):
        n = C.c_int(0)
        self.ctx.check(lib().vloam_map_get_cube(self._h, stream, kind, cube, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), np.float32)
        if n.value:
            self.ctx.check(lib().vloam_map_get_cube(self._h, stream, kind, cube, out.ctypes.data_as(c_fp), n.value, C.byref(n)))
        return out

    def close(self):
        if self._h:
            lib().vloam_lidar_destroy(self._h)
            self._h = C.c_void_p()
        if self._own_ctx:
            self.ctx.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VisualOdometry:
    """Mirror of the in-scope part of vloam::VisualOdometry for `batch` streams: processImage (Shi-Tomasi detection, ORB
    description, descriptor matching: detKeypoints / descKeypoints / matchDescriptors), then depth association, residual
    construction and the solve (processPointCloud, solveNlsAll)."""

    def __init__(self, ctx: Context | None = None, batch: int = 1, max_points: int = 131072, max_matches: int = 1024,
                 remove_VO_outlier: int = 100, max_num_iterations: int = 100):
        self.ctx = ctx or Context()
        self._own_ctx = ctx is None
        self.batch, self.max_matches = batch, max_matches
        self.remove_VO_outlier, self.max_num_iterations = remove_VO_outlier, max_num_iterations
        self._h = C.c_void_p()
        self.ctx.check(lib().vloam_vo_create(self.ctx._h, batch, max_points, max_matches, C.byref(self._h)))

    # -- visual_odometry.cpp:132-155
    def setUpPointCloud(self, cam_T_velo, rect0_T_cam, P_rect0):
        a = np.ascontiguousarray(cam_T_velo, np.float32).ravel()
        b = np.ascontiguousarray(rect0_T_cam, np.float32).ravel()
        c = np.ascontiguousarray(P_rect0, np.float32).ravel()
        assert a.size == 16 and b.size == 16 and c.size == 12
        self.ctx.check(lib().vloam_vo_set_calibration(self._h, a.ctypes.data_as(c_fp), b.ctypes.data_as(c_fp), c.ctypes.data_as(c_fp)))

    # -- visual_odometry.cpp:86-90
    def reset(self):
        self.ctx.check(lib().vloam_vo_reset(self._h))

    # -- visual_odometry.cpp:157-186
    def processPointCloud(self, point_cloud, n_points=None):
        a = np.ascontiguousarray(point_cloud, np.float32)
        if a.ndim == 2:
            a = a[None]
        assert a.shape[0] == self.batch and a.shape[2] in (3, 4)
        n = np.full(self.batch, a.shape[1], np.int32) if n_points is None else np.ascontiguousarray(n_points, np.int32)
        self.ctx.check(lib().vloam_vo_process_cloud(self._h, _ptr(a), _ptr(n), a.shape[2], a.shape[1]))

    def processPointCloudDevice(self, xyz_dev, n_points_dev, stride: int, slab_points: int):
        """processPointCloud on clouds that already live in device memory ((batch, slab_points, stride) float32)."""
        self.ctx.check(lib().vloam_vo_process_cloud_device(self._h, _ptr(xyz_dev), _ptr(n_points_dev), stride, slab_points))

    def queryDepth(self, xy, slot: int = 0, stream: int = 0):
        q = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros(q.shape[0], np.float32)
        self.ctx.check(lib().vloam_vo_query_depth(self._h, stream, slot, q.ctypes.data_as(c_fp), q.shape[0], out.ctypes.data_as(c_fp)))
        return out

    def buckets(self, slot: int = 0, stream: int = 0):
        nb = 249 * 75
        bx, by, bd = (np.zeros(nb, np.float32) for _ in range(3))
        bc = np.zeros(nb, np.int32)
        self.ctx.check(lib().vloam_vo_get_buckets(self._h, stream, slot, bx.ctypes.data_as(c_fp), by.ctypes.data_as(c_fp),
                                                  bd.ctypes.data_as(c_fp), bc.ctypes.data_as(c_ip)))
        return bx.reshape(249, 75), by.reshape(249, 75), bd.reshape(249, 75), bc.reshape(249, 75)

    # -- visual_odometry.cpp:254-450
    def solveNlsAll(self, prev_uv, curr_uv, n_matches=None, init=None):
        """prev_uv / curr_uv: (batch, m, 2) or (m, 2) matched pixels; init: (batch, 6) angle-axis + t or None."""
        p = np.ascontiguousarray(prev_uv, np.float32)
        c = np.ascontiguousarray(curr_uv, np.float32)
        if p.ndim == 2:
            p, c = p[None], c[None]
        m = p.shape[1]
        assert p.shape == c.shape and p.shape[0] == self.batch and m <= self.max_matches
        pp = np.zeros((self.batch, self.max_matches, 2), np.float32)
        cc = np.zeros((self.batch, self.max_matches, 2), np.float32)
        pp[:, :m] = p
        cc[:, :m] = c
        nm = np.full(self.batch, m, np.int32) if n_matches is None else np.ascontiguousarray(n_matches, np.int32)
        ini = np.ascontiguousarray(init, np.float64).reshape(self.batch, 6) if init is not None else None
        out = np.zeros((self.batch, 8))
        self.ctx.check(lib().vloam_vo_solve(self._h, _ptr(pp), _ptr(cc), _ptr(nm), _ptr(ini), self.remove_VO_outlier,
                                            self.max_num_iterations, out.ctypes.data_as(c_dp)))
        return {"angles_0to1": out[:, 0:3].copy(), "t_0to1": out[:, 3:6].copy(), "counter32": out[:, 6].astype(int),
                "counter22": out[:, 7].astype(int)}

    # -- image_util.cpp:214-296 (BF + NORM_HAMMING + KNN + ratio test: the configuration of visual_odometry.cpp:34-37)
    def matchDescriptors(self, desc_query, desc_train, kp_query=None, kp_train=None, ratio: float = 0.8):
        """desc_*: per stream an (n, 32) uint8 array (the cv::Mat of cv::ORB), or one array for batch 1; kp_*: matching (n, 2)
        float32 keypoint pixels or None.  Returns per stream the accepted (queryIdx, trainIdx, distance) rows in query order
        and the raw 2-NN table (idx0, idx1, d0, d1) per query row."""
        if isinstance(desc_query, np.ndarray):
            desc_query, desc_train = [desc_query], [desc_train]
            kp_query, kp_train = (None if kp_query is None else [kp_query]), (None if kp_train is None else [kp_train])
        B, M = self.batch, self.max_matches
        assert len(desc_query) == B and len(desc_train) == B
        dq = np.zeros((B, M, 32), np.uint8); dt = np.zeros((B, M, 32), np.uint8)
        nq = np.zeros(B, np.int32); nt = np.zeros(B, np.int32)
        kq = kt = None
        if kp_query is not None:
            kq = np.zeros((B, M, 2), np.float32); kt = np.zeros((B, M, 2), np.float32)
        for b in range(B):
            a, t = np.ascontiguousarray(desc_query[b], np.uint8).reshape(-1, 32), np.ascontiguousarray(desc_train[b], np.uint8).reshape(-1, 32)
            nq[b], nt[b] = a.shape[0], t.shape[0]
            dq[b, :nq[b]] = a; dt[b, :nt[b]] = t
            if kq is not None:
                kq[b, :nq[b]] = kp_query[b]; kt[b, :nt[b]] = kp_train[b]
        m = np.zeros((B, M, 3), np.int32)
        nm = np.zeros(B, np.int32)
        self.ctx.check(lib().vloam_vo_match_descriptors(self._h, _ptr(dq), _ptr(nq), _ptr(dt), _ptr(nt), _ptr(kq), _ptr(kt), float(ratio),
                                                        _ptr(m), _ptr(nm)))
        knn = np.zeros((B, M, 4), np.int32)
        self.ctx.check(lib().vloam_vo_get_knn(self._h, _ptr(knn)))
        return [{"matches": m[b, :nm[b]].copy(), "knn": knn[b, :nq[b]].copy()} for b in range(B)]

    def detKeypoints(self, images, max_corners: int = 1024, quality_level: float = 0.03, min_distance: float = 7.5):
        """ImageUtil::detKeypoints with DetectorType::ShiTomasi (image_util.cpp:11-37): images = (batch, H, W) or (H, W) uint8;
        returns per stream the (n, 2) float32 corner coordinates (x, y) in cv::goodFeaturesToTrack's order."""
        a = np.ascontiguousarray(images, np.uint8)
        if a.ndim == 2:
            a = a[None]
        assert a.shape[0] == self.batch, a.shape
        out = np.zeros((self.batch, max_corners, 2), np.float32)
        n = np.zeros(self.batch, np.int32)
        self.ctx.check(lib().vloam_vo_detect_corners(self._h, _ptr(a), a.shape[1], a.shape[2], max_corners, float(quality_level),
                                                     float(min_distance), _ptr(out), _ptr(n)))
        self._det_shape = a.shape[1:]
        return [out[b, :n[b]].copy() for b in range(self.batch)]

    # -- image_util.cpp:162-212 (DescriptorType::ORB)
    def descKeypoints(self, keypoints=None, images=None):
        """ImageUtil::descKeypoints with DescriptorType::ORB: keypoints = per stream an (n, 2) float32 array of cv::KeyPoint::pt
        (one array for batch 1), or None = the corners of the last detKeypoints call (still on the device); images = (batch, H, W)
        or (H, W) uint8, or None = the images of that call.  Returns per stream {"keypoints": the key points ORB keeps (the
        reference's vector after the call), "index": their positions in the input list, "descriptors": (n, 32) uint8}."""
        B, M = self.batch, self.max_matches
        kp = n = img = None
        H = W = 0
        if keypoints is not None:
            if isinstance(keypoints, np.ndarray) and keypoints.ndim == 2:
                keypoints = [keypoints]
            assert len(keypoints) == B
            kp = np.zeros((B, M, 2), np.float32)
            n = np.zeros(B, np.int32)
            for b in range(B):
                a = np.ascontiguousarray(keypoints[b], np.float32).reshape(-1, 2)
                n[b] = a.shape[0]
                kp[b, :min(n[b], M)] = a[:M]
        if images is not None:
            img = np.ascontiguousarray(images, np.uint8)
            if img.ndim == 2:
                img = img[None]
            assert img.shape[0] == B
            H, W = img.shape[1:]
        kxy = np.zeros((B, M, 2), np.float32)
        kidx = np.zeros((B, M), np.int32)
        desc = np.zeros((B, M, 32), np.uint8)
        nk = np.zeros(B, np.int32)
        self.ctx.check(lib().vloam_vo_describe_orb(self._h, _ptr(img), H, W, _ptr(kp), _ptr(n), _ptr(kxy), _ptr(kidx), _ptr(desc), _ptr(nk)))
        return [{"keypoints": kxy[b, :nk[b]].copy(), "index": kidx[b, :nk[b]].copy(), "descriptors": desc[b, :nk[b]].copy()} for b in range(B)]

    # -- visual_odometry.cpp:92-130
    def processImage(self, images, fetch: bool = True):
        """VisualOdometry::processImage: detection, description and (from the second frame on) matching against the previous
        frame on the device; call reset() first, like the reference's frame loop.  fetch: return per stream
        {"n_keypoints", "n_matches"}; the features and matches are read with frame_features() / matches()."""
        if hasattr(images, "data_ptr"):            # a torch tensor: pinned host memory or device memory, used in place
            a = images
            assert a.dtype.itemsize == 1 and a.is_contiguous() and a.dim() == 3
        else:
            a = np.ascontiguousarray(images, np.uint8)
            if a.ndim == 2:
                a = a[None]
        assert a.shape[0] == self.batch, a.shape
        nk = np.zeros(self.batch, np.int32) if fetch else None
        nm = np.zeros(self.batch, np.int32) if fetch else None
        self.ctx.check(lib().vloam_vo_process_image(self._h, _ptr(a), int(a.shape[1]), int(a.shape[2]), _ptr(nk), _ptr(nm)))
        self._det_shape = (int(a.shape[1]), int(a.shape[2]))
        return {"n_keypoints": nk, "n_matches": nm} if fetch else None

    def detect_status(self) -> int:
        """1 when a frame of the last detection overflowed its candidate list (only of interest after processImage(fetch=False))."""
        s = C.c_int(0)
        self.ctx.check(lib().vloam_vo_get_detect_status(self._h, C.byref(s)))
        return s.value

    def frame_features(self, slot: int = 0):
        """keypoints[slot] / descriptors[slot] of the processImage chain (0 = current frame, 1 = previous), per stream."""
        B, M = self.batch, self.max_matches
        kxy = np.zeros((B, M, 2), np.float32)
        desc = np.zeros((B, M, 32), np.uint8)
        nk = np.zeros(B, np.int32)
        self.ctx.check(lib().vloam_vo_get_frame_features(self._h, slot, _ptr(kxy), _ptr(desc), _ptr(nk)))
        return [{"keypoints": kxy[b, :nk[b]].copy(), "descriptors": desc[b, :nk[b]].copy()} for b in range(B)]

    def matches(self):
        """The (queryIdx, trainIdx, distance) rows of the last processImage / matchDescriptors call, per stream."""
        m = np.zeros((self.batch, self.max_matches, 3), np.int32)
        nm = np.zeros(self.batch, np.int32)
        self.ctx.check(lib().vloam_vo_get_matches(self._h, _ptr(m), _ptr(nm)))
        return [m[b, :nm[b]].copy() for b in range(self.batch)]

    def corner_response(self, stream: int = 0):
        """cv::cornerMinEigenVal map of the last detKeypoints call, (H, W) float32."""
        H, W = self._det_shape
        r = np.zeros((H, W), np.float32)
        self.ctx.check(lib().vloam_vo_get_corner_response(self._h, stream, _ptr(r), H * W))
        return r

    def match_buffers(self):
        """(query_uv, train_uv, n_matches) device addresses of the last matchDescriptors call with keypoints."""
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().vloam_vo_get_match_buffers(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def match_uv(self):
        """Host copies of the matched (query, train) pixel pairs, (batch, max_matches, 2) each."""
        q = np.zeros((self.batch, self.max_matches, 2), np.float32)
        t = np.zeros_like(q)
        self.ctx.check(lib().vloam_vo_get_match_uv(self._h, _ptr(q), _ptr(t)))
        return q, t

    def solveNlsAllDevice(self, prev_uv_dev, curr_uv_dev, n_matches_dev, init_dev=None):
        """solveNlsAll on device-resident matches ((batch, max_matches, 2) float32, (batch,) int32); asynchronous."""
        self.ctx.check(lib().vloam_vo_solve_device_async(self._h, _ptr(prev_uv_dev), _ptr(curr_uv_dev), _ptr(n_matches_dev),
                                                         _ptr(init_dev), self.remove_VO_outlier, self.max_num_iterations))

    def result(self):
        out = np.zeros((self.batch, 8))
        self.ctx.check(lib().vloam_vo_get_result(self._h, out.ctypes.data_as(c_dp)))
        return {"angles_0to1": out[:, 0:3].copy(), "t_0to1": out[:, 3:6].copy(), "counter32": out[:, 6].astype(int),
                "counter22": out[:, 7].astype(int)}

    def exportLOPrior(self, velo_T_cam0, prior_dev):
        """VloamTF::VO2VeloAndBase (vloam_tf.cpp:59-63) on the device: writes velo_last_VOT_velo_curr as (batch, 7)
        float64 q(xyzw) t into prior_dev, ready for LidarOdometryMapping.laserOdometryIO(prior=prior_dev, fetch=False)."""
        m = np.ascontiguousarray(velo_T_cam0, np.float64).reshape(4, 4)
        self.ctx.check(lib().vloam_vo_export_lo_prior(self._h, m.ctypes.data_as(c_dp), _ptr(prior_dev)))

    def residuals(self, stream: int = 0):
        t = np.zeros(self.max_matches, np.int32)
        o = np.zeros((self.max_matches, 5))
        self.ctx.check(lib().vloam_vo_get_residuals(self._h, stream, t.ctypes.data_as(c_ip), o.ctypes.data_as(c_dp)))
        return t, o

    def trace(self, stream: int = 0):
        rec = np.zeros((8, 7))
        info = np.zeros(4, np.int32)
        para = np.zeros(7)
        self.ctx.check(lib().vloam_vo_get_trace(self._h, stream, rec.ctypes.data_as(c_dp), info.ctypes.data_as(c_ip), para.ctypes.data_as(c_dp)))
        return {"iterations": rec[: min(int(info[0]), 8)].copy(), "n_records": int(info[0]), "termination": int(info[1]), "para": para}

    def close(self):
        if self._h:
            lib().vloam_vo_destroy(self._h)
            self._h = C.c_void_p()
        if self._own_ctx:
            self.ctx.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
