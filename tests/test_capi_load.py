"""CPU tests: the C-ABI library builds for sm_100a, loads, exports every symbol include/vloam_b200.h declares,
and fails loudly (error code, no fallback) when no CUDA device is present."""
import ctypes
import os

import pytest


def test_library_exports_every_declared_symbol():
    import vloam_b200 as V
    assert os.path.exists(V.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(V.LIB_PATH)
    names = V.exported_symbols_in_header()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_params_default_are_the_reference_launch_values():
    import vloam_b200 as V
    p = V.default_lidar_params()
    # src/lidar_odometry_mapping/launch/loam_velodyne_HDL_64_kitti.launch:3-16, src/vloam_main/launch/vloam_main.launch:4
    assert (p.scan_line, p.minimum_range, p.mapping_skip_frame) == (64, 5.0, 1)
    assert (p.mapping_line_resolution, p.mapping_plane_resolution, p.detach_VO_LO) == (0.4, 0.8, 1)
    assert (p.lo_outer_passes, p.lo_max_iterations, p.lm_outer_passes, p.lm_max_iterations) == (2, 4, 2, 4)


def test_kernel_name_table():
    import vloam_b200 as V
    L = V.lib()
    names = [L.vloam_ctx_kernel_name(i).decode() for i in range(L.vloam_ctx_kernel_count())]
    assert "sr_curvature" in names and "lo_solve" in names and len(set(names)) == len(names)


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA the product path must refuse to run rather than fall back to the oracle."""
    import torch
    import vloam_b200 as V
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(V.VloamError):
        V.Context()
    with pytest.raises(V.VloamError):
        V.LidarOdometryMapping(batch=1, max_points=1024)


def test_oracle_is_not_linked_into_the_product():
    import subprocess
    import vloam_b200 as V
    out = subprocess.run(["nm", "-D", V.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out and "oracle" not in out
