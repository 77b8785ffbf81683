"""The ROS-facing adapters (vloam-cmu-16833_b200/adapter/*_b200.h) are header-compatible with the reference's classes but need
ROS, PCL, tf2 and Eigen headers, none of which exist in this image.  tests/stubs holds minimal stand-ins, so that CI at least
compiles and links them (CPU) and runs them (GPU) through the call sequences of the reference's callers
(vloam_main_node.cpp:134-167 for the façade, lidar_odometry_mapping.cpp:65-154 for the stage classes)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vloam-cmu-16833_b200", "lib")
TOPICS = {"/velodyne_cloud_2", "/laser_cloud_sharp", "/laser_cloud_less_sharp", "/laser_cloud_flat", "/laser_cloud_less_flat",      # scan_registration.cpp:64-68
          "/laser_cloud_corner_last", "/laser_cloud_surf_last", "/velodyne_cloud_3", "/laser_odom_to_init", "/laser_odom_path",       # laser_odometry.cpp:108-112
          "/laser_cloud_surround", "/laser_cloud_map", "/velodyne_cloud_registered", "/aft_mapped_to_init", "/aft_mapped_path"}        # laser_mapping.cpp:103-108


def _build(tmp_path):
    import vloam_b200
    vloam_b200.build()
    exe = str(tmp_path / "adapter_frame_loop")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "tests", "stubs"), os.path.join(ROOT, "tests", "native", "adapter_frame_loop.cpp"),
                           "-o", exe, "-L" + LIB, "-lvloam_b200", "-Wl,-rpath," + LIB, "-L/usr/local/cuda/lib64", "-lcudart"])
    return exe


def test_adapters_compile_and_link_against_stub_ros(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "compile-only"], capture_output=True, text=True, check=True).stdout
    assert {l.split()[1] for l in out.splitlines() if l.startswith("topic")} == TOPICS


@pytest.mark.gpu
def test_adapters_run_the_callers_frame_loop(tmp_path, synth):
    """Façade and stage classes, three frames each: the poses they publish into vloam_tf->world_MOT_base_last equal the
    Python mirror's (same library underneath, bit for bit), every topic of the reference is published once per frame."""
    import vloam_b200 as V
    exe = _build(tmp_path)
    s = synth.ScanStream(71, n_cols=512)
    files = []
    for k in range(3):
        f = str(tmp_path / f"scan{k}.bin")
        s.scan(k).astype(np.float32).tofile(f)
        files.append(f)
    out = subprocess.run([exe, "run"] + files, capture_output=True, text=True, check=True).stdout
    lom = V.LidarOdometryMapping(batch=1, max_points=1 << 18)
    want = []
    for k in range(3):
        lom.reset(); lom.scanRegistrationIO(s.scan(k)); lom.laserOdometryIO(np.array([[0, 0, 0, 1.0, 0, 0, 0]])); mp = lom.laserMappingIO()
        want.append(np.r_[mp["q_w_curr"][0], mp["t_w_curr"][0]])
    counts = lom.feature_counts()[0]
    lom.close()
    for tag in ("facade", "stages"):
        rows = [l.split() for l in out.splitlines() if l.startswith(tag)]
        assert len(rows) == 3
        for k, r in enumerate(rows):
            assert np.array_equal(np.array(r[2:9], float), want[k]), (tag, k)
    last = [l.split() for l in out.splitlines() if l.startswith("stages 2")][0]
    assert int(last[last.index("sharp") + 1]) == counts[1] and int(last[last.index("lessFlat") + 1]) == counts[4]
    published = {l.split()[1]: int(l.split()[2]) for l in out.splitlines() if l.startswith("published")}
    for t in ("/laser_odom_to_init", "/laser_odom_path", "/aft_mapped_to_init", "/aft_mapped_path", "/laser_cloud_sharp"):
        assert published[t] == 6, (t, published)        # three frames through the façade + three through the stage classes
    assert published["/velodyne_cloud_registered"] == 6                  # laser_mapping.cpp:797-805: every frame
    assert published["/laser_cloud_map"] == 2                            # :778: frameCount % map_pub_number (= 2) == 0, once per run of 3
    assert published["/laser_cloud_surround"] == 0                       # advertised, never published (the reference does neither)


# ------------------------------------------------------------------------------------------------ vloam::VisualOdometry
def _build_vo(tmp_path):
    import vloam_b200
    vloam_b200.build()
    exe = str(tmp_path / "vo_adapter_frame_loop")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "tests", "stubs"), os.path.join(ROOT, "tests", "native", "vo_adapter_frame_loop.cpp"),
                           "-o", exe, "-L" + LIB, "-lvloam_b200", "-Wl,-rpath," + LIB, "-L/usr/local/cuda/lib64", "-lcudart"])
    return exe


def test_vo_adapter_compiles_and_links_against_stub_ros_and_opencv(tmp_path):
    """adapter/visual_odometry_b200.h (same class name, methods and public members as the reference's visual_odometry.h:40-121)
    compiles against the stub ROS / PCL / OpenCV headers; its constructor advertises the reference's topic (visual_odometry.cpp:7)."""
    exe = _build_vo(tmp_path)
    out = subprocess.run([exe, "compile-only"], capture_output=True, text=True, check=True).stdout
    assert [l.split()[1] for l in out.splitlines() if l.startswith("topic")] == ["/point_cloud_follow_VO"]


@pytest.mark.gpu
def test_vo_adapter_runs_the_callers_frame_loop(tmp_path, synth):
    """vloam::VisualOdometry (the adapter) through three frames of vloam_main_node.cpp:134-166 — processImage as one device chain
    (the stub ImageUtil would abort if the adapter fell back to OpenCV) — against the Python mirror: the same key points,
    descriptors and matches in the members the caller reads, and the same solved motion, bit for bit."""
    import vloam_b200 as V
    exe = _build_vo(tmp_path)
    g = np.load(os.path.join(ROOT, "tests", "golden", "vo_detect_cv2.npz"))
    base = g["kitti_image"]
    imgs = [base, np.ascontiguousarray(np.roll(base, (-2, 5), axis=(0, 1))), np.ascontiguousarray(np.roll(base, (-3, 9), axis=(0, 1)))]
    H, W = base.shape
    s = synth.ScanStream(83, n_cols=512)
    cam_T_velo, rect0, P = synth.kitti_like_calibration()
    calib = str(tmp_path / "calib.txt")
    with open(calib, "w") as f:
        f.write(" ".join(repr(float(v)) for v in np.r_[cam_T_velo.ravel(), rect0[:3, :3].ravel(), P.ravel()]))
    args, scans = [], []
    for k, im in enumerate(imgs):
        fi, fs = str(tmp_path / f"img{k}.bin"), str(tmp_path / f"scan{k}.bin")
        im.tofile(fi)
        sc = s.scan(k)
        sc = sc[np.isfinite(sc).all(1)].astype(np.float32)
        sc.tofile(fs)
        scans.append(sc)
        args += [fi, fs]
    out = subprocess.run([exe, "run", str(H), str(W), calib] + args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("vo ")]
    assert len(rows) == 3

    def checksum(values):
        s_ = 0
        for v in values:
            s_ = (s_ * 31 + int(v)) & 0xFFFFFFFF
        return s_
    vo = V.VisualOdometry(batch=1, max_points=1 << 18, max_matches=2048)           # the C++ mirror's defaults
    rect_q8 = rect0.copy()
    rect_q8[3, 3] = 0.0                                                            # the ROS path leaves (3, 3) = 0, and so does the adapter
    prev = None
    for k, im in enumerate(imgs):
        vo.reset()
        r = vo.processImage(im)
        vo.setUpPointCloud(cam_T_velo, rect_q8, P)
        vo.processPointCloud(np.c_[scans[k], np.ones(len(scans[k]), np.float32)])   # pcl::PointXYZ: 16-byte records
        cur = vo.frame_features(0)[0]
        m = vo.matches()[0]
        row = rows[k]
        val = {row[i]: i for i in range(len(row))}
        assert int(row[val["keypoints"] + 1]) == len(cur["keypoints"]) == int(row[val["rows"] + 1]) and len(cur["keypoints"]) > 500
        assert int(row[val["desc"] + 1]) == checksum(cur["descriptors"].ravel())
        assert int(row[val["matches"] + 1]) == len(m)
        assert int(row[val["msum"] + 1]) == checksum(m[:, :2].ravel())
        motion = np.array([float(v) for v in row[val["motion"] + 1: val["motion"] + 7]])
        if k == 0:
            assert len(m) == 0 and not motion.any()
        else:
            assert len(m) > 300
            res = vo.solveNlsAll(prev["keypoints"][m[:, 0]], cur["keypoints"][m[:, 1]])
            assert np.array_equal(motion, np.r_[res["angles_0to1"][0], res["t_0to1"][0]]), (k, motion)
            assert res["counter32"][0] + res["counter22"][0] > 50
            # the same solve without the host in between: the matched pixels processImage left on the device go straight in
            vo.solveNlsAllDevice(*vo.match_buffers())
            dres = vo.result()
            assert np.array_equal(dres["angles_0to1"], res["angles_0to1"]) and np.array_equal(dres["t_0to1"], res["t_0to1"])
            assert dres["counter32"][0] == res["counter32"][0] and dres["counter22"][0] == res["counter22"][0]
        prev = cur
    published = {l.split()[1]: int(l.split()[2]) for l in out.stdout.splitlines() if l.startswith("published")}
    assert published == {"/point_cloud_follow_VO": 3, "/visual_odom_to_init": 3, "/visual_odom_path": 3}
    vo.close()
