"""CPU tests of the formats on either side of the hot path: KITTI .bin loader, PointCloud2
payload binding, KITTI-format pose dump — Python mirror vs the C++ adapter header vs a direct numpy restatement of the
reference lines (point_cloud_util.cpp:118-146, vloam_main_node.cpp:148, vloam_tf.cpp:77-153)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AD = os.path.join(ROOT, "vloam-cmu-16833_b200", "adapter")
BIN = os.path.join(ROOT, "vloam-cmu-16833_b200", "lib", "wire_formats_check")


@pytest.fixture(scope="module")
def checker():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(AD, "wire_formats_check.cpp"), "-o", BIN])
    return BIN


def test_kitti_bin_loader(tmp_path, checker):
    from vloam_b200 import wire
    rng = np.random.default_rng(0)
    pts = rng.normal(0, 20, (12345, 4)).astype(np.float32)
    path = tmp_path / "000000.bin"
    (tmp_path / "000000.bin").write_bytes(pts.tobytes() + b"\x00\x01")      # trailing partial record is ignored
    got = wire.load_kitti_bin(str(path))
    assert got.shape == (12345, 4) and np.array_equal(got, pts)
    out = subprocess.check_output([checker, "bin", str(path)], text=True).split()
    assert int(out[0]) == 12345
    np.testing.assert_allclose([float(v) for v in out[1:]], pts.astype(np.float64).sum(0), rtol=0, atol=1e-3)
    big = np.zeros((300000, 4), np.float32)                                 # the reference reads at most 1 000 000 floats
    (tmp_path / "big.bin").write_bytes(big.tobytes())
    assert wire.load_kitti_bin(str(tmp_path / "big.bin")).shape == (250000, 4)
    assert int(subprocess.check_output([checker, "bin", str(tmp_path / "big.bin")], text=True).split()[0]) == 250000


@pytest.mark.parametrize("step,ox,oy,oz", [(16, 0, 4, 8), (32, 0, 4, 8), (24, 4, 8, 12), (20, 0, 8, 16)])
def test_pointcloud2_binding(tmp_path, checker, step, ox, oy, oz):
    from vloam_b200 import wire
    rng = np.random.default_rng(step)
    n = 777
    xyz = rng.normal(0, 30, (n, 3)).astype(np.float32)
    msg = np.zeros((n, step), np.uint8)
    for k, o in enumerate((ox, oy, oz)):
        msg[:, o:o + 4] = xyz[:, k:k + 1].copy().view(np.uint8)
    view, zero_copy = wire.pointcloud2_xyz(msg.reshape(-1), n, step, {"x": ox, "y": oy, "z": oz})
    assert np.array_equal(np.asarray(view)[:, :3], xyz)
    assert zero_copy == (oy == ox + 4 and oz == ox + 8 and step % 4 == 0)
    if zero_copy:
        assert np.shares_memory(view, msg)
    path = tmp_path / "msg.raw"
    path.write_bytes(msg.tobytes())
    out = subprocess.check_output([checker, "pc2", str(path), str(step), str(ox), str(oy), str(oz), str(n)], text=True).split()
    assert int(out[0]) == int(zero_copy)
    assert int(out[1]) == (step // 4 if zero_copy else 3)
    np.testing.assert_allclose([float(v) for v in out[2:]], xyz.astype(np.float64).sum(0), atol=1e-3)


def test_kitti_pose_dump_matches_reference_formula(tmp_path, checker):
    from vloam_b200 import wire
    rng = np.random.default_rng(3)

    def rand_qt(scale):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        return q, rng.normal(0, scale, 3)
    bq, bt = rand_qt(1.0)
    poses = [rand_qt(50.0) for _ in range(6)]
    path = tmp_path / "poses.txt"
    with open(path, "w") as f:
        for q, t in [(bq, bt)] + poses:
            f.write(" ".join(repr(float(v)) for v in list(q) + list(t)) + "\n")
    w = wire.Cam0StartFrameWriter(wire.mat4_from_qt(bq, bt))
    assert w.write(None, -1, np.eye(4)) == ""
    lines = [w.write(None, k, wire.mat4_from_qt(q, t)) for k, (q, t) in enumerate(poses)]
    # first line: identity
    np.testing.assert_allclose(np.array(lines[0].split(), float).reshape(3, 4), np.eye(4)[:3], atol=1e-6)
    # direct restatement of vloam_tf.cpp:77-101 with general inverses
    B = wire.mat4_from_qt(bq, bt)
    init = [np.linalg.inv(B) @ wire.mat4_from_qt(q, t) @ B for q, t in poses]
    for k, line in enumerate(lines):
        ref = (np.linalg.inv(init[0]) @ init[k])[:3, :4]
        np.testing.assert_allclose(np.array(line.split(), float).reshape(3, 4), ref, atol=2e-5)
        assert len(line.split()) == 12 and all("." in v and len(v.split(".")[1]) == 6 for v in line.split())
    cpp = subprocess.check_output([checker, "poses", str(path)], text=True).splitlines(keepends=True)
    assert len(cpp) == len(lines)
    for a, b in zip(cpp, lines):
        np.testing.assert_allclose(np.array(a.split(), float), np.array(b.split(), float), atol=2e-6)
