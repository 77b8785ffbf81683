#!/usr/bin/env python
"""bench.py — scans/sec of the VLOAM per-scan LiDAR hot path on B200 (BASELINE.json metric).

Workload at N = 1: BASELINE.json's metric is "scans/sec ... laserOdometry+Mapping", i.e. configs[2]: scan registration,
laserOdometry and laserMapping (scan-to-submap on a pre-built 1 M-point voxel map, 2 passes x 5 LM iterations = "10 GN
iters") over a synthetic HDL-64 64 x 2048 range-image stream, `--batch` independent streams driven in lock-step
(`--workload sr_lo` = configs[1], scanRegistration + laserOdometry only; `--workload vloam` = configs[3], with visual
odometry feeding the LiDAR prior).  One *step* = one scan of every stream through the whole path.  Three measurements of
the same work:

  value  inputs already resident in HBM (a pool of scans larger than L2), timed with CUDA events on the
         launching stream, barrier + synchronize on both sides, max over ranks;
  e2e    the same steps through the reference-facing API with HOST buffers: pinned host -> device upload
         of every scan and device -> host read of the poses inside the timed region;
  cpu_baseline / --impl reference
         the CPU oracle (oracle/: restatement of the reference's Ceres/PCL path; the reference itself cannot
         be built here) timed on the box's host cores.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the algorithmic-byte model behind `roofline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_RINGS, N_COLS = 64, 2048
TRAJ_SCANS = 64         # consecutive scans of every base sequence: a forward drive of 32..96 m (5..15 m/s at 10 Hz); a run of up
                        # to 64 steps never replays a scan, a longer one drives the same road back (63, 62, ...) and forth
N_BASE = 4              # distinct base sequences per rank, tiled across the batch
BENCH_SEED = 1234


def pingpong(i: int, n: int) -> int:
    p = i % (2 * n - 2)
    return p if p < n else 2 * n - 2 - p


BENCH_YAW_RATE_MAX = 0.01      # rad/s: the 64-scan drives stay on the scene's road (synth.ScanStream)


def scan_index(i: int) -> int:
    """Trajectory scan visited at step i (0, 1, ..., 63, 62, ...)."""
    return pingpong(i, TRAJ_SCANS)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~5 ms while the timed region runs (nvidia-smi's own
    polling is too coarse for a region of a few hundred ms; it stays as the fallback when NVML cannot be loaded)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.idx, self.uuid = gpu_index, uuid
        self.rows, self.sm, self.bits = [], [], 0
        self.proc, self.nv, self.h, self.run = None, None, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid) if self.uuid else pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv, self.run = pynvml, True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self.run = False
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
                    "reasons": sorted(n for b, n in self.REASONS if self.bits & b), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def make_base_scans(rank: int, needed, world: int = 1, batch: int = N_BASE, with_streams: bool = False):
    """N_BASE sequences of this rank; seqs[i][k] = scan k (float32 (n, 3), NaN = no return) for every k in `needed`."""
    from vloam_b200 import dist as D
    from vloam_b200 import synth
    seqs, streams = [], []
    first = D.shard_streams(world * batch, world, rank)[0] if batch > 0 else 0     # this rank's first global stream
    for i in range(N_BASE):
        s = synth.ScanStream(D.stream_seed(BENCH_SEED, first + i, N_BASE) if world > 1 else BENCH_SEED + i, n_cols=N_COLS,
                             yaw_rate_max=BENCH_YAW_RATE_MAX, on_road=True)
        streams.append(s)
        seqs.append({k: s.scan(k) for k in needed})
    return (seqs, streams) if with_streams else seqs


_FRAME_BASE = None


def frame_of(stream_index: int, step: int):
    """configs[3]: the camera frame of stream `stream_index` at replay step `step` — the committed KITTI-sized test image
    (376 x 1241) moved horizontally by a per-stream offset, 4 px further on odd steps.  The frames feed the image front end
    (detection, ORB description, matching: its cost is in the step); they do not depict the LiDAR scene, so the solve takes
    the geometry-consistent synthetic matches of synth.make_matches."""
    global _FRAME_BASE
    if _FRAME_BASE is None:
        _FRAME_BASE = np.load(os.path.join(ROOT, "tests", "golden", "vo_detect_cv2.npz"))["kitti_image"]
    return np.roll(_FRAME_BASE, (stream_index * 37) % 300 + 4 * (step % 2), axis=1)


class CpuFrontEnd:
    """VisualOdometry::processImage as the reference runs it (visual_odometry.cpp:92-130): the three OpenCV calls, one thread."""

    def __init__(self):
        try:
            import cv2
            cv2.setNumThreads(1)
            self.cv2, self.orb, self.bf = cv2, cv2.ORB_create(), cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)
        except ImportError:
            self.cv2 = None
        self.prev = None

    def process(self, img):
        if self.cv2 is None:
            return
        cv2 = self.cv2
        c = cv2.goodFeaturesToTrack(img, 1024, 0.03, 7.5, None, blockSize=5, useHarrisDetector=False, k=0.04)
        c = np.zeros((0, 2), np.float32) if c is None else c.reshape(-1, 2)
        _, d = self.orb.compute(img, [cv2.KeyPoint(float(x), float(y), 5.0) for x, y in c])
        if self.prev is not None and d is not None and len(self.prev) and len(d):
            knn = self.bf.knnMatch(self.prev, d, 2)
            _ = [m[0] for m in knn if len(m) == 2 and m[0].distance < 0.8 * m[1].distance]
        self.prev = d


class CpuChain:
    """One stream through the oracle in the order of vloam_main_node.cpp:125-180 (CPU arm / cpu_baseline).
    workload: sr_lo | sr_lo_lm | vloam (adds VisualOdometry and feeds its result to laserOdometry as the prior)."""

    def __init__(self, O, workload, map_cubes, lm_iterations=4):
        self.O, self.workload = O, workload
        self.t = {"sr_ms": 0.0, "lo_ms": 0.0, "lm_ms": 0.0, "vo_ms": 0.0, "fe_ms": 0.0, "scans": 0}
        if workload == "vloam":
            self.fe = CpuFrontEnd()
            from vloam_b200 import synth
            self.calib = synth.kitti_like_calibration()
            self.velo_T_cam0 = np.linalg.inv(self.calib[0].astype(np.float64))
            self.vo = O.VisualOdometry(*self.calib)
            self.lo = O.LaserOdometry(detach_VO_LO=False)
            self.lm = O.LaserMapping()
            self.frames = 0
        else:
            self.pipe = O.Pipeline()
            self.lm = self.pipe.lm
        self.lm.set_iterations(2, lm_iterations)
        for (kind, cube), pts in map_cubes.items():
            self.lm.set_cube(kind, cube, pts)

    def process(self, scan, matches=None, frame=None):
        O = self.O
        if self.workload != "vloam":
            self.pipe.process(scan, do_mapping=self.workload == "sr_lo_lm")
            self.t.update(self.pipe.timings())
            return
        tf = time.perf_counter()
        if frame is not None:
            self.fe.process(frame)
        t0 = time.perf_counter()
        self.t["fe_ms"] += 1e3 * (t0 - tf)
        self.vo.reset()
        self.vo.process_cloud(scan)
        prior = np.r_[0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
        if self.frames > 0 and matches is not None:
            r = self.vo.solve(matches[0], matches[1])
            prior = O.vo_to_lo_prior(r["angles_0to1"], r["t_0to1"], self.velo_T_cam0)
        t1 = time.perf_counter()
        sr = O.scan_registration(scan)
        t2 = time.perf_counter()
        self.lo.solve(sr, prior_q=prior[:4], prior_t=prior[4:])
        t3 = time.perf_counter()
        self.lm.reset()
        self.lm.input_from_lo(self.lo)
        self.lm.solve()
        t4 = time.perf_counter()
        self.frames += 1
        self.t["vo_ms"] += 1e3 * (t1 - t0); self.t["sr_ms"] += 1e3 * (t2 - t1); self.t["lo_ms"] += 1e3 * (t3 - t2)
        self.t["lm_ms"] += 1e3 * (t4 - t3); self.t["scans"] += 1

    def timings(self):
        return dict(self.t)

    def lm_iterations(self):
        """LM iterations the oracle ran in each outer pass of the last mapped scan (iteration records minus the initial one)."""
        return [int(t["iterations"].shape[0]) - 1 for t in self.lm.trace()]


def cpu_matches(scan_stream, i):
    """(prev_uv, curr_uv) for replay step i of one base sequence (None for the first frame)."""
    from vloam_b200 import synth
    if i == 0:
        return None
    kp, k = scan_index(i - 1), scan_index(i)
    pu, cu, _ = synth.make_matches(scan_stream, max(kp, k), n_matches=800)
    return (pu, cu) if k > kp else (cu, pu)


WORKLOAD_NAME = {
    "sr_lo": ("scanRegistration+laserOdometry", "configs[1]: scanRegistration + laserOdometry on 1xB200, synthetic 64x2048 range-image stream"),
    "sr_lo_lm": ("scanRegistration+laserOdometry+laserMapping",
                 "configs[2]: laserOdometry + laserMapping scan-to-submap (1M-pt voxel map, 10 GN iters) on 1xB200, fed by scanRegistration"),
    "vloam": ("visualOdometry(image front end+depth+solve)+scanRegistration+laserOdometry(VO prior)+laserMapping",
              "configs[3]: full VLOAM, visual odometry (Shi-Tomasi + ORB + BF matching on 1241x376 frames, depth association, solve) "
              "feeding LiDAR odometry and mapping"),
}


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def config_dict(args, streams_per_gpu, handles, world, do_map, extra=None):
    """The `config` object both arms print (same keys, same values for the same command line)."""
    c = {"workload": WORKLOAD_NAME[args.workload][1], "streams_per_gpu": streams_per_gpu, "handles": handles,
         "points_per_scan": N_RINGS * N_COLS, "lo_passes": 2, "lo_iterations_per_pass": 4, "lm_passes": 2,
         "lm_iterations_per_pass": args.lm_iterations, "map_points": args.map_points if do_map else 0, "solver_mode": args.solver_mode, "cuda_graphs": args.graphs,
         "trajectory": f"{N_BASE} seeded base sequences x {TRAJ_SCANS} consecutive scans (forward drive along the scene's road, full-size scans throughout, no replay within {TRAJ_SCANS} steps), "
                       "tiled across the streams"}
    cap = N_RINGS * N_COLS
    # (both arms print these two as well: they describe the B200 arm's run of this configuration — the timed steps read a different
    # slab of the device-resident scan pool each, which exceeds the 126 MB L2 from 84 streams on)
    c["l2_policy"] = (f"inputs larger than L2: a different [{streams_per_gpu}, {cap}, 3] slab ({streams_per_gpu * cap * 12 / 1e6:.0f} MB) of the "
                      "device-resident scan pool every step")
    if args.parallelism == "point":
        c["parallelism"] = (f"point-sharded x{world}: replicated scans, queries split across ranks, partial normal equations (28 doubles per tile, "
                            f"{8 * 28 * 8} bytes per stream and evaluation) all-reduced by NCCL between the accumulate and step launches of laser "
                            "odometry and laser mapping")
    elif args.parallelism == "point-peer":
        c["parallelism"] = (f"point-sharded x{world}: replicated scans, correspondences split across ranks, 28-double normal equations summed inside "
                            "the solve kernel over NVLink peer memory")
    else:
        c["parallelism"] = f"stream-sharded x{world} (no data-path collective)"
    if extra:
        c.update(extra)
    return c


def run_reference(args, rank):
    """The reference's CPU path (oracle port: the reference needs ROS/PCL/Ceres/Eigen, none installed) on all host
    threads: one independent stream per thread; a step = one scan on every thread.  Same trajectory, same pre-built map,
    same iteration limits as the B200 arm."""
    if rank != 0:
        return
    from oracle import pyoracle as O
    from concurrent.futures import ThreadPoolExecutor
    O.build()
    T = max(1, len(os.sched_getaffinity(0)))
    do_map = args.workload in ("sr_lo_lm", "vloam")
    n_steps = args.warmup + args.steps
    needed = sorted({scan_index(i) for i in range(n_steps)})
    seqs, scan_streams = make_base_scans(0, needed, with_streams=True)
    cubes = synth_map_cubes(args.map_points, BENCH_SEED) if do_map else {}
    pipes = [CpuChain(O, args.workload, cubes, args.lm_iterations) for _ in range(T)]
    mt = {(t % N_BASE, i): cpu_matches(scan_streams[t % N_BASE], i) for t in range(min(T, N_BASE)) for i in range(n_steps)} \
        if args.workload == "vloam" else {}

    def one(t, i):
        pipes[t].process(seqs[t % N_BASE][scan_index(i)], mt.get((t % N_BASE, i)), frame_of(t, i) if args.workload == "vloam" else None)

    with ThreadPoolExecutor(T) as ex:
        for i in range(args.warmup):
            list(ex.map(lambda t: one(t, i), range(T)))
        t0 = time.perf_counter()
        for i in range(args.warmup, args.warmup + args.steps):
            list(ex.map(lambda t: one(t, i), range(T)))
        dt = time.perf_counter() - t0
    value = T * args.steps / dt
    tm = pipes[0].timings()
    its = pipes[0].lm_iterations() if do_map else None
    line = {
        "impl": "reference", "metric": "scans/sec (HDL-64, 64x2048 pts) " + WORKLOAD_NAME[args.workload][0],
        "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 points / f64 solve", "data": "synthetic",
        "config": config_dict(args, args.batch, args.handles, max(1, args.gpus), do_map),
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": T, "kind": "port",
                         "sample": f"{T} threads x {args.steps} scans, one stream per thread (the GPU arm's streams_per_gpu streams are a batch "
                                   f"dimension this CPU path does not have); per-thread SR {tm['sr_ms']/max(1,tm['scans']):.1f} ms, "
                                   f"LO {tm['lo_ms']/max(1,tm['scans']):.1f} ms, LM {tm['lm_ms']/max(1,tm['scans']):.1f} ms, VO {tm.get('vo_ms', 0.0)/max(1,tm['scans']):.1f} ms, image front end (OpenCV) {tm.get('fe_ms', 0.0)/max(1,tm['scans']):.1f} ms per scan"
                                   + (f"; LM iterations executed in the last scan's passes: {its}" if its else "")},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def synth_map_cubes(n_points: int, seed: int):
    from vloam_b200 import synth
    return synth.map_cubes(n_points, seed)


# --------------------------------------------------------------------------------------------- our arm (B200)
# timing id (vloam_ctx_kernel_name) -> the __global__ functions launched under it (names as ncu prints them)
NCU_KERNELS = {
    "lm_associate": ["lm_knn"], "lm_voxel": ["lm_voxel_stack"], "lm_index": ["lm_index_build"], "lm_insert": ["lm_insert_keys"],
    "lm_place": ["lm_place", "lm_compact_copy", "lm_write_back"], "lm_misc": ["lm_transform_update", "lm_export_pose"],
    "lm_refilter": ["lm_refilter_merge", "lm_refilter"],
}


def algorithmic_bytes(kernel, c):
    """Algorithmic HBM bytes of ONE launch of `kernel` over the whole batch (DESIGN.md "Measurement").
    c: dict of batch totals: N (input points), Np (kept points), nLF, nLS, nSharp, nFlat (current scan features),
    nLFlast, nLSlast (previous scan's targets)."""
    feats = c["nSharp"] + c["nLS"] + c["nFlat"]
    table = {
        "sr_find_ends": 0,
        "sr_classify": c["N"] * (12 + 1),                                   # xyz read, ring id written
        "sr_scan": 0,
        "sr_scatter": c["N"] * (12 + 1) + c["Np"] * 16,                    # xyz + ring id read, XYZI written
        "sr_curvature": c["Np"] * 21,                                       # XYZI read, curvature + gap flag written
        "sr_pick_features": c["Np"] * (4 + 1 + 1) + feats * 4,              # curvature + gap flag read, label written
        "sr_less_flat_voxel": c["Np"] * (1 + 16) + c["nLF"] * 16,           # label + XYZI read, centroids written
        "sr_pack": c["nLF"] * 32 + feats * (4 + 16 + 16 + 4),
        "lo_build_grid": (c["nLS"] + c["nLF"]) * (16 + 16 + 4),             # cloud read, column-sorted copy + index written
        "lo_associate": (c["nSharp"] + c["nFlat"]) * 32 + (c["nLSlast"] + c["nLFlast"]) * 20,
        "lo_associate_brute": (c["nSharp"] + c["nFlat"]) * 32 + (c["nLSlast"] + c["nLFlast"]) * 16,
        "lo_solve": (c["nSharp"] + c["nFlat"]) * 16 + c["nSharp"] * 48 + c["nFlat"] * 64,
        "lo_export_pose": 0,
        # laser mapping (M = sub-map points, S = down-sampled scan points)
        "lm_prepare": 0, "lm_misc": 0,
        "lm_voxel": (c["nLS"] + c["nLF"]) * (16 + 16),
        "lm_index": 0,                                      # only cubes without a column index are (re)indexed: none in steady state
        # per query: the point (16 B), five positions out (20 B), <= 6 run look-ups (2 x 2 B each + a layer start) and the
        # candidates of its runs (16 B each; count measured by the statistics pass, ~13 on the benchmark map)
        "lm_associate": c.get("S", 0) * (16 + 20 + 6 * 8 + c.get("cand", 13.0) * 16),
        "lm_fit": c.get("S", 0) * (16 + 20 + 5 * 16 + 72),
        "lm_solve": c.get("S", 0) * 80,
        "lm_insert": c.get("S", 0) * 48,
        "lm_refilter": c.get("Mw", 0) * (16 + 16) // 2,     # two launches (merge path / sort path): rewritten cubes read, merged cubes written
        "lm_place": c.get("Mw", 0) * (16 * 3 + 16) // 3,    # three launches; write-back: staged read, slab + sorted copy written
    }
    return table.get(kernel, 0)


def bind_near_gpu(local_rank: int):
    """Best effort: run this process (and so first-touch its pinned host buffers) on the CPUs of the GPU's NUMA node.
    Returns a short description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        devid = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"gpu on numa node {node}, none of its cpus allowed"
        os.sched_setaffinity(0, allowed)
        return f"bound to {len(allowed)} cpus of numa node {node}"
    except Exception as e:          # sysfs not visible in the container, ...
        return f"not bound ({type(e).__name__})"


def run_ours(args, rank, world, local_rank):
    import torch
    import vloam_b200 as V
    from vloam_b200 import dist as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_near_gpu(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # NCCL's own log lines (version banner, the `nranks` evidence at NCCL_DEBUG=INFO) go to stderr: stdout carries exactly
        # one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    do_vo = args.workload == "vloam"
    do_map = args.workload in ("sr_lo_lm", "vloam")
    cap = N_RINGS * N_COLS
    point = args.parallelism in ("point", "point-peer") and world > 1
    point_nccl = args.parallelism == "point"
    n_total = args.warmup + 2 * args.steps + 2           # timed leg + per-kernel timing pass + statistics pass
    needed = sorted({scan_index(i) for i in range(n_total)})
    # point-sharded: every rank replays the SAME streams (rank-independent seeds) and owns a slice of their correspondences
    t_gen = time.perf_counter()
    seqs, scan_streams = make_base_scans(0 if point else rank, needed, 1 if point else world, B, with_streams=True)
    t_gen = time.perf_counter() - t_gen

    # Scan pools.  Host: the N_BASE x len(needed) distinct scans once, pinned; a stream's scan is uploaded from the base
    # sequence it replays (one host buffer per stream, vloam_scan_registration_ptrs).  Device: [B, cap, 3] per trajectory
    # step (stream b replays base sequence b % N_BASE), gathered on the device from the uploaded base scans.
    slot_of = {k: j for j, k in enumerate(needed)}
    host_base = torch.empty((N_BASE, len(needed), cap, 3), dtype=torch.float32).pin_memory()
    hb = host_base.numpy()
    for i in range(N_BASE):
        for k in needed:
            hb[i, slot_of[k]] = seqs[i][k]
    dev_base = host_base.to(dev, non_blocking=True)
    sel = torch.arange(B, device=dev) % N_BASE
    dev_pool = {k: dev_base[sel, slot_of[k]].contiguous() for k in needed}
    n_host = np.full(B, cap, np.int32)
    n_dev = torch.from_numpy(n_host).to(dev)
    pool_bytes = len(needed) * B * cap * 12
    base_ptr = host_base.data_ptr()
    host_ptrs = {k: np.array([base_ptr + (((b % N_BASE) * len(needed) + slot_of[k]) * cap * 12) for b in range(B)], np.uint64) for k in needed}

    # configs[3]: matched keypoint pixels for every (previous scan -> current scan) pair the run visits
    M = 1024
    match_host, match_dev, velo_T_cam0, calib = {}, {}, None, None
    if do_vo:
        from vloam_b200 import synth
        calib = synth.kitti_like_calibration()
        velo_T_cam0 = np.linalg.inv(calib[0].astype(np.float64))
        pairs = sorted({(scan_index(i - 1), scan_index(i)) for i in range(1, n_total)})
        for (kp, k) in pairs:
            hp = torch.zeros((B, M, 2), dtype=torch.float32).pin_memory()
            hc = torch.zeros((B, M, 2), dtype=torch.float32).pin_memory()
            hn = torch.zeros((B,), dtype=torch.int32).pin_memory()
            per_base = []
            for sst in scan_streams:
                pu, cu, _ = synth.make_matches(sst, max(kp, k), n_matches=800)
                per_base.append((pu, cu) if k > kp else (cu, pu))      # a long run also drives backwards in time
            for b in range(B):
                pu, cu = per_base[b % N_BASE]
                hp[b, :len(pu)] = torch.from_numpy(pu); hc[b, :len(cu)] = torch.from_numpy(cu); hn[b] = len(pu)
            match_host[(kp, k)] = (hp, hc, hn)
            match_dev[(kp, k)] = (hp.to(dev), hc.to(dev), hn.to(dev))
        # camera frames: one per stream and step parity (frame_of), pinned on the host and resident on the device
        frames_host = [torch.from_numpy(np.stack([frame_of(rank * B + b, par) for b in range(B)])).pin_memory() for par in range(2)]
        frames_dev = [f.to(dev) for f in frames_host]
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    ctx = V.Context(device=local_rank, cuda_stream=stream.cuda_stream)

    use_fused = bool(args.graphs) and do_map and not do_vo and not point
    map_cubes = synth_map_cubes(args.map_points, BENCH_SEED) if do_map else {}
    map_cap = int(2 ** np.ceil(np.log2(max(1 << 17, 1.3 * args.map_points)))) if do_map else 1 << 17

    class Group:
        """Streams [b0, b1) of the batch: one LidarOdometryMapping handle (+ one VisualOdometry handle for configs[3]) on
        one context / CUDA stream, driven in the order of vloam_main_node.cpp:125-180."""

        def __init__(self, ctx_, b0, b1):
            self.ctx, self.b0, self.b1, self.nb = ctx_, b0, b1, b1 - b0
            self.lom = V.LidarOdometryMapping(ctx_, batch=self.nb, max_points=cap, map_capacity_points=map_cap,
                                              detach_VO_LO=0 if do_vo else 1, lm_max_iterations=args.lm_iterations,
                                              solver_mode=args.solver_mode)
            for (kind, cube), pts in map_cubes.items():      # the same pre-built map under every stream
                for b in range(self.nb):
                    self.lom.map_set_cube(kind, cube, pts, stream=b)
            self.vo, self.prior = None, None
            if do_vo:
                self.vo = V.VisualOdometry(ctx_, batch=self.nb, max_points=cap, max_matches=M)
                self.vo.setUpPointCloud(*calib)
                self.prior = torch.zeros((self.nb, 7), dtype=torch.float64, device=dev)
                self.prior[:, 3] = 1.0
                self.uv = [torch.zeros((self.nb, M, 2), dtype=torch.float32, device=dev) for _ in range(2)]
                self.nm = torch.zeros((self.nb,), dtype=torch.int32, device=dev)
            self.steps = 0

        def _vo(self, i, xyz, n, stride, slab, pu, cu, nm, frames):
            self.vo.reset()
            self.vo.processImage(frames, fetch=False)                # visual_odometry.cpp:92-130: enqueues, no synchronisation
            self.vo.processPointCloudDevice(xyz, n, stride, slab)
            if self.steps > 0:
                self.vo.solveNlsAllDevice(pu, cu, nm)
                self.vo.exportLOPrior(velo_T_cam0, self.prior)

        def step_dev(self, i):
            k = scan_index(i)
            sl = slice(self.b0, self.b1)
            if use_fused:      # one call per frame, replayed as a CUDA graph (not with the VO stage interleaved, configs[3])
                self.lom.processDevice(dev_pool[k][sl], n_dev[sl], 3, cap, use_graph=True)
                self.steps += 1
                return
            self.lom.reset()
            self.lom.scanRegistrationDevice(dev_pool[k][sl], n_dev[sl], 3, cap)
            if do_vo:
                pu, cu, nm = match_dev[(scan_index(i - 1), k)] if i > 0 else (None, None, None)
                self._vo(i, dev_pool[k][sl], n_dev[sl], 3, cap, None if pu is None else pu[sl], None if cu is None else cu[sl],
                         None if nm is None else nm[sl], frames_dev[i % 2][sl])
            self.lom.laserOdometryIO(prior=self.prior, fetch=False)
            if do_map:
                self.lom.laserMappingIO(fetch=False)
            self.steps += 1

        def step_host(self, i, first):
            """One scan in flight: enqueue scan i (pinned host -> device upload on the copy stream + kernels), then read
            scan i-1's poses (device -> host) while scan i runs, so uploads overlap compute."""
            k = scan_index(i)
            sl = slice(self.b0, self.b1)
            if use_fused:
                self.lom.processPtrs(host_ptrs[k][sl], n_host[sl], 3, use_graph=True, keep=host_base)
                self.steps += 1
                return None if first else self.lom.lo_pose(prev=True)
            self.lom.reset()
            self.lom.scanRegistrationPtrs(host_ptrs[k][sl], n_host[sl], 3, keep=host_base)
            if do_vo:
                xyz, n, stride, slab = self.lom.input_device()          # the cloud is uploaded once and read twice
                if i > 0:
                    hp, hc, hn = match_host[(scan_index(i - 1), k)]
                    self.uv[0].copy_(hp[sl], non_blocking=True); self.uv[1].copy_(hc[sl], non_blocking=True)
                    self.nm.copy_(hn[sl], non_blocking=True)
                self._vo(i, xyz, n, stride, slab, self.uv[0], self.uv[1], self.nm, frames_host[i % 2][sl])
                self.lom.input_consumed()
            self.lom.laserOdometryIO(prior=self.prior, fetch=False)
            if do_map:
                self.lom.laserMappingIO(fetch=False)
            self.steps += 1
            return None if first else self.lom.lo_pose(prev=True)

        def close(self):
            if self.vo is not None:
                self.vo.close()
            self.lom.close()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return D.max_over_ranks(ms, dist, dev)

    def all_ranks(x):
        """[x of rank 0, x of rank 1, ...] (floats)."""
        if dist is None:
            return [float(x)]
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)

    # ---------------- leg 1: device-resident inputs -> `value`
    # The B streams are split over H groups, each on its own CUDA stream, so one group's single-CTA-per-stream
    # kernels (the LM solves, ring-end scans) overlap the other groups' wide kernels instead of idling the SMs.
    H = 1 if point else max(1, min(args.handles, B))
    bounds = [(B * h) // H for h in range(H + 1)]
    streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(H - 1)]
    ctxs = [ctx] + [V.Context(device=local_rank, cuda_stream=s_.cuda_stream) for s_ in streams[1:]]
    groups = []
    for h in range(H):
        with torch.cuda.stream(streams[h]):
            groups.append(Group(ctxs[h], bounds[h], bounds[h + 1]))
    loms = [g.lom for g in groups]
    def enable_sharding(lom_):
        if point_nccl:
            D.enable_point_sharding_nccl(lom_, dist, dev)
        else:
            D.enable_point_sharding(lom_, dist, dev)

    def disable_sharding(lom_):
        if point_nccl:
            lom_.shard_nccl_destroy()
        else:
            lom_.shard_disable()
    if point:
        enable_sharding(groups[0].lom)

    def step_dev(i, serial=False):
        for h in range(H):
            with torch.cuda.stream(streams[h]):
                groups[h].step_dev(i)
            if serial:
                torch.cuda.synchronize()

    def join_streams():
        for s_ in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(s_)
            stream.wait_event(ev)

    def counters():
        return np.concatenate([hd.lm_counters() for hd in loms]) if do_map else None

    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step_dev(i)
        barrier()
        launches0 = sum(c_.launch_count for c_ in ctxs)
        cnt0 = counters()
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s_ in streams[1:]:
            s_.wait_event(e0)
        for i in range(args.warmup, args.warmup + args.steps):
            step_dev(i)
        join_streams()
        e1.record(stream)
        barrier()
        ms_rank = e0.elapsed_time(e1)
        ms_ranks = all_ranks(ms_rank)
        ms_dev = max(ms_ranks)
        clocks = sampler.stop() if rank == 0 else None
        launches = sum(c_.launch_count for c_ in ctxs) - launches0
        cnt1 = counters()
        counts = np.concatenate([hd.feature_counts() for hd in loms]).astype(np.int64)
        pose_dev = {kk: np.concatenate([hd.lo_pose()[kk] for hd in loms]) for kk in loms[0].lo_pose()}
        lm_status = np.concatenate([hd.lm_status() for hd in loms]) if do_map else np.zeros((B, 2), np.int32)
        # per-kernel durations: the same steps again with a CUDA-event pair around every launch; with H > 1 the groups
        # are run one after the other here so that the per-kernel times are not inflated by overlap
        for c_ in ctxs:
            c_.enable_timing(True)
            c_.kernel_timings(reset=True)
        for i in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
            step_dev(i, serial=H > 1)
        barrier()
        ktimes = {}
        for c_ in ctxs:
            for kname, (ms_k, n_k) in c_.kernel_timings(reset=True).items():
                a = ktimes.get(kname, (0.0, 0))
                ktimes[kname] = (a[0] + ms_k, a[1] + n_k)
            c_.enable_timing(False)
        # the byte model must describe the SAME scans the durations come from: once more over those scans (ping-pong index),
        # reading the point / feature counts after every step (a read-back per step would disturb the timing pass itself)
        counts_sum, lm_info_sum, n_k_steps = None, None, 0
        for i in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
            step_dev(i)
            c_i = np.concatenate([hd.feature_counts() for hd in loms]).astype(np.int64)
            counts_sum = c_i if counts_sum is None else counts_sum + c_i
            if do_map:
                l_i = np.concatenate([hd.lm_info() for hd in loms]).astype(np.int64)
                lm_info_sum = l_i if lm_info_sum is None else lm_info_sum + l_i
            n_k_steps += 1
        barrier()
        # statistics pass (untimed): candidate points the 5-NN search tests per query
        knn = None
        if do_map:
            for hd in loms:
                hd.set_debug_stats(True)
            c_a = counters()
            for i in range(args.warmup + 2 * args.steps, args.warmup + 2 * args.steps + 2):
                step_dev(i)
            barrier()
            c_b = counters()
            for hd in loms:
                hd.set_debug_stats(False)
            dq, dc = float((c_b[:, 16] - c_a[:, 16]).sum()), float((c_b[:, 17] - c_a[:, 17]).sum())
            knn = {"queries": dq, "candidates": dc, "candidates_per_query": dc / max(dq, 1.0)}
    # whole-job units / max-over-ranks time (helpers shared with the CPU test tests/test_dist_gloo.py)
    value = (B * args.steps / (ms_dev * 1e-3)) if point else D.aggregate_throughput(B * args.steps, ms_rank, dist, dev)

    if args.legs == "device":
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms_dev / args.steps, "gpu_launches": int(launches), "legs": "device",
                              "knn": knn,
                              "kernels": {k: {"avg_us": 1e3 * v[0] / v[1], "launches": v[1]} for k, v in ktimes.items()}}), flush=True)
        return
    lm_info_all = np.concatenate([hd.lm_info() for hd in loms]).astype(np.int64) if do_map else None
    map_stats_all = np.concatenate([hd.map_stats() for hd in loms]).astype(np.int64) if do_map else None
    lm_tr = loms[0].lm_trace(1, 0) if do_map else None
    if point:
        disable_sharding(groups[0].lom)
    for g in groups:
        g.close()               # fresh handles for the e2e leg (same inputs, same number of steps -> same final poses)

    # ---------------- leg 2: host buffers through the public API -> `e2e`
    # The same H handles as the device leg, each uploading its streams' scans from pinned host memory on its own copy
    # stream (one scan in flight per handle) and reading its poses back.
    groups2 = []
    for h in range(H):
        with torch.cuda.stream(streams[h]):
            groups2.append(Group(ctxs[h], bounds[h], bounds[h + 1]))
    g2 = groups2[0]
    if point:
        enable_sharding(g2.lom)

    def step_host(i, first):
        for h in range(H):
            with torch.cuda.stream(streams[h]):
                groups2[h].step_host(i, first)

    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step_host(i, i == 0)
        if args.warmup:
            for g in groups2:
                g.lom.lo_pose()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.perf_counter()
        e0.record(stream)
        for s_ in streams[1:]:
            s_.wait_event(e0)
        for i in range(args.warmup, args.warmup + args.steps):
            step_host(i, i == args.warmup)
        poses_h = [g.lom.lo_pose() for g in groups2]    # drain: the last scan's results are read inside the timed region too
        join_streams()
        e1.record(stream)
        barrier()
        t_host1 = time.perf_counter()
        ms_e2e_rank = max(e0.elapsed_time(e1), 1e3 * (t_host1 - t_host0))
        ms_e2e_ranks = all_ranks(ms_e2e_rank)
        ms_e2e = max(ms_e2e_ranks)
    pose_host = {kk: np.concatenate([p_[kk] for p_ in poses_h]) for kk in poses_h[0]}
    e2e_value = (1 if point else world) * B * args.steps / (ms_e2e * 1e-3)
    h2d = int(B * cap * 12 + B * 4 + (B * (2 * M * 2 * 4 + 4) + int(frames_host[0].numel()) if do_vo else 0))
    d2h = int(B * 16 * 8)
    shard_err = g2.lom.shard_status() if (point and not point_nccl) else 0
    # same inputs, same number of steps -> both legs must end on identical poses
    same = bool(np.array_equal(pose_dev["t_w_curr"], pose_host["t_w_curr"]))

    # The host -> device ceiling of this box for these uploads (every rank at once, nothing else running): first as B separate
    # copies of one 1.57 MB scan from the pinned pool per step (one cudaMemcpyAsync per stream, what round 2 started with),
    # then as contiguous copies of the same bytes.  e2e cannot beat the second; it is printed beside e2e.
    with torch.cuda.stream(stream):
        sink = torch.empty((B, cap, 3), dtype=torch.float32, device=dev)
        reps = 6
        barrier()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(stream)
        for r in range(reps):
            k = needed[r % len(needed)]
            for b in range(B):      # two upload queues, like the library's copy streams
                with torch.cuda.stream(side if b & 1 else stream):
                    sink[b].copy_(host_base[b % N_BASE, slot_of[k]], non_blocking=True)
        stream.wait_stream(side)
        b_.record(stream)
        barrier()
        h2d_percopy_ms = max(all_ranks(a_.elapsed_time(b_) / reps))
        # ... and as large contiguous copies of the same number of bytes: what the link delivers at best, and what the
        # library's batched upload (cudaMemcpyBatchAsync) is measured to reach with one buffer per stream
        flat_host, flat_dev = host_base.view(-1), sink.view(-1)
        chunk = min(flat_host.numel(), flat_dev.numel())
        barrier()
        a_.record(stream)
        for r in range(reps):
            done_ = 0
            while done_ < flat_dev.numel():
                m_ = min(chunk, flat_dev.numel() - done_)
                flat_dev[done_:done_ + m_].copy_(flat_host[:m_], non_blocking=True)
                done_ += m_
        b_.record(stream)
        barrier()
        h2d_ms = max(all_ranks(a_.elapsed_time(b_) / reps))
        del sink
    h2d_ceiling_gbs = B * cap * 12 / (h2d_ms * 1e-3) / 1e9             # per GPU, all ranks copying at once
    h2d_ceiling_scans = (1 if point else world) * B / (h2d_ms * 1e-3)
    h2d_percopy_gbs = B * cap * 12 / (h2d_percopy_ms * 1e-3) / 1e9

    # ---------------- leg 3: single-stream latency (batch = 1), context only
    lat_ms, lat = None, None
    if rank == 0:
        lom1 = V.LidarOdometryMapping(ctx, batch=1, max_points=cap, map_capacity_points=map_cap, lm_max_iterations=args.lm_iterations)
        for (kind, cube), pts in map_cubes.items():
            lom1.map_set_cube(kind, cube, pts)
        n1 = n_dev[0:1].contiguous()
        n_lat = 40
        lat_scan = (lambda i: needed[pingpong(i, len(needed))]) if len(needed) > 1 else (lambda i: needed[0])   # stays inside the pool
        lat = {}
        with torch.cuda.stream(stream):
            def step1(i, graph):
                if do_map and graph is not None:
                    lom1.processDevice(dev_pool[lat_scan(i)][0:1], n1, 3, cap, use_graph=graph)
                    return
                lom1.reset()
                lom1.scanRegistrationDevice(dev_pool[lat_scan(i)][0:1], n1, 3, cap)
                lom1.laserOdometryIO(fetch=False)
                if do_map:
                    lom1.laserMappingIO(fetch=False)
            i_lat = 0
            for name, graph in (("stage_calls", None), ("one_call_cuda_graph", True)):
                for _ in range(6):
                    step1(i_lat, graph); i_lat += 1
                torch.cuda.synchronize()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t_w0 = time.perf_counter()
                a.record(stream)
                for _ in range(n_lat):
                    step1(i_lat, graph); i_lat += 1
                b_.record(stream)
                torch.cuda.synchronize()
                lat[name] = {"ms_per_scan_device": a.elapsed_time(b_) / n_lat, "ms_per_scan_wall": 1e3 * (time.perf_counter() - t_w0) / n_lat}
            lat_ms = min(v["ms_per_scan_device"] for v in lat.values())
        lom1.close()

    if rank != 0:
        return

    # ---------------- per-kernel table and the roofline of the dominant kernel
    peaks, peak_kind = load_peaks()
    cm = counts_sum / max(n_k_steps, 1)                          # mean per scan over the per-kernel timing pass
    tot = {"N": int(B * cap), "Np": int(cm[:, 0].sum()), "nSharp": int(cm[:, 1].sum()), "nLS": int(cm[:, 2].sum()),
           "nFlat": int(cm[:, 3].sum()), "nLF": int(cm[:, 4].sum())}
    tot["nLSlast"], tot["nLFlast"] = tot["nLS"], tot["nLF"]
    if do_map:
        info = lm_info_sum / max(n_k_steps, 1)
        tot["M"] = int(info[:, 4].sum() + info[:, 5].sum())
        tot["S"] = int(info[:, 6].sum() + info[:, 7].sum())
        tot["Mw"] = int(map_stats_all[:, :, 8].sum())
        tot["cand"] = knn["candidates_per_query"] if knn else 0.0
    kern = {}
    for name, (ms, cnt) in ktimes.items():
        by = int(algorithmic_bytes(name, tot)) // H          # one launch covers one handle's B/H streams
        kern[name] = {"ms_total": ms, "launches": cnt, "avg_us": 1e3 * ms / cnt, "share": None,
                      "alg_bytes_per_launch": by, "gbs": (by / (ms / cnt * 1e-3) / 1e9) if ms > 0 else None}
    ksum = sum(v["ms_total"] for v in kern.values()) or 1.0
    for v in kern.values():
        v["share"] = v["ms_total"] / ksum
    dom = max(kern, key=lambda k: kern[k]["ms_total"])
    # measured DRAM traffic / issue utilisation of the same kernels from the committed ncu --set full capture
    # (profiles/ncu_traffic.json: per launch and per stream there; scaled to this run's streams per launch)
    tj = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)["kernels"]
    except Exception:
        tj = {}

    def ncu_of(timing_id):
        members = [m_ for m_ in NCU_KERNELS.get(timing_id, [timing_id]) if m_ in tj]
        if not members:
            return None, None
        tr_ = sum(tj[m_]["dram_bytes_per_launch_per_stream"] for m_ in members) / len(members) * (B / H)
        ia = [tj[m_].get("issue_active_pct") for m_ in members if tj[m_].get("issue_active_pct") is not None]
        return tr_, (sum(ia) / len(ia) if ia else None)
    traffic, issue_active = ncu_of(dom)
    dur_s = kern[dom]["ms_total"] / kern[dom]["launches"] * 1e-3
    roof = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": (kern[dom]["gbs"] or 0.0) / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
            "alg_bytes_per_launch": kern[dom]["alg_bytes_per_launch"],
            "dram_gbs_from_ncu_traffic": (traffic / dur_s / 1e9) if traffic else None,
            "dram_frac_from_ncu_traffic": (traffic / dur_s / 1e9 / peaks["hbm_gbs"]) if traffic else None,
            "issue_active_pct_ncu": issue_active,
            "note": "achieved = algorithmic bytes per launch (DESIGN.md section 6) / CUDA-event duration; traffic = ncu dram read+write bytes "
                    "per launch (profiles/ncu_traffic.json).  The irregular searches (lm_associate, lo_associate) are served from L1/L2 and "
                    "bound by instruction issue, not by HBM: for them read issue_active_pct_ncu and candidates_per_query, not frac."}
    named = {}
    for nm_ in ("sr_curvature", "lo_solve", "lm_solve", "lm_accumulate", "lo_accumulate"):      # the kernels the north star names
        if nm_ in kern and kern[nm_]["gbs"] is not None:
            t_, ia_ = ncu_of(nm_)
            named[nm_] = {"gbs": kern[nm_]["gbs"], "frac": kern[nm_]["gbs"] / peaks["hbm_gbs"], "avg_us": kern[nm_]["avg_us"],
                          "share": kern[nm_]["share"], "ncu_dram_bytes_per_launch": t_}
    curv = kern.get("sr_curvature")

    # ---------------- CPU baseline: oracle, 1 thread, bounded sample of the same trajectory
    from oracle import pyoracle as O
    O.build()
    pipe = CpuChain(O, args.workload, map_cubes, args.lm_iterations)
    n_cpu = args.cpu_scans if not do_map else max(4, args.cpu_scans // 4)      # ~10-15 s of CPU work either way
    n_cpu = min(n_cpu, n_total)
    cpu_m = [cpu_matches(scan_streams[0], i) for i in range(n_cpu)] if do_vo else [None] * n_cpu
    t0 = time.perf_counter()
    for i in range(n_cpu):
        pipe.process(seqs[0][scan_index(i)], cpu_m[i], frame_of(0, i) if do_vo else None)
    cpu_dt = time.perf_counter() - t0
    tm = pipe.timings()
    cpu_value = n_cpu / cpu_dt

    lm_work = None
    if do_map:
        d = (cnt1 - cnt0).astype(np.float64)
        scans_t = max(d[:, 0].sum(), 1.0)
        per = lambda col: float(d[:, col].sum() / scans_t)       # noqa: E731  (mean per scan and stream over the timed region)
        lm_work = {
            "scans_mapped": int(d[:, 0].sum()), "scans_solved": int(d[:, 1].sum()),
            "lm_iterations_executed_per_pass": [per(2), per(3)], "lm_iterations_limit_per_pass": args.lm_iterations,
            "cubes_entering_window_indexed_per_scan": [per(4), per(5)], "cubes_rewritten_per_scan": [per(6), per(7)],
            "cubes_merged_per_scan": per(8), "of_which_patched_in_place": per(9), "cubes_fully_refiltered_per_scan": per(10),
            "cubes_appended_per_scan": per(11), "voxels_inserted_per_scan": per(12), "repacks": int(d[:, 13].sum()),
            "knn_queries_per_scan_per_pass": per(14), "residual_blocks_last_pass_per_scan": per(15),
            "knn_candidates_per_query": knn["candidates_per_query"] if knn else None,
            "streams_with_map_errors": int((lm_status[:, 1] != 0).sum()),
        }

    line = {
        "metric": "scans/sec (HDL-64, 64x2048 pts) " + WORKLOAD_NAME[args.workload][0],
        "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if point else "weak", "vs_baseline": None,
        "dtype": "f32 points / f64 solve",
        "data": f"synthetic ({N_BASE} seeded base sequences x {TRAJ_SCANS}-scan forward trajectories tiled across the batch; generated in {t_gen:.1f} s)",
        "config": config_dict(args, B, args.handles, world, do_map),
        "scan_pool_gb": pool_bytes / 1e9, "shard_status": int(shard_err),
        "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "poses_identical_to_device_leg": same,
                "h2d_gbs_per_gpu": h2d / (ms_e2e / args.steps * 1e-3) / 1e9,
                "h2d_ceiling_gbs_per_gpu": h2d_ceiling_gbs, "h2d_ceiling_scans_per_s": h2d_ceiling_scans,
                "h2d_one_memcpy_per_stream_gbs_per_gpu": h2d_percopy_gbs,
                "upload": "one pinned host buffer per stream, all streams of a handle in one cudaMemcpyBatchAsync",
                "ms_per_step_per_rank": [m_ / args.steps for m_ in ms_e2e_ranks], "host_numa": numa},
        "ms_per_step_per_rank": [m_ / args.steps for m_ in ms_ranks],
        "exchange": exchange_report(args, world, do_map) if point else None,
        "gpu_launches": int(launches),
        "roofline": roof,
        "north_star_kernels": named,
        "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in kern.items()},
        "curvature_kernel": None if curv is None else {"gbs": curv["gbs"], "frac": curv["gbs"] / peaks["hbm_gbs"]},
        "single_stream_latency_ms": lat_ms, "single_stream_latency": lat,
        "laser_mapping_work": lm_work,
        "map": None if not do_map else {
            "points_per_stream": float(map_stats_all[:, :, 0].sum() / B), "cubes_per_stream": float(map_stats_all[:, :, 3].sum() / B),
            "cubes_rewritten_last_scan_per_stream": float(map_stats_all[:, :, 5].sum() / B),
            "points_rewritten_last_scan_per_stream": float(map_stats_all[:, :, 8].sum() / B),
            "fixed_point_cubes_per_stream": float(map_stats_all[:, :, 4].sum() / B), "repacks": int(map_stats_all[:, :, 6].sum()),
            "queries_per_scan_stream0": [int(lm_info_all[0, 6]), int(lm_info_all[0, 7])],
            "residual_blocks_last_pass_stream0": {"edge": int(lm_tr["n_corner"]), "plane": int(lm_tr["n_plane"]), "lm_records": int(lm_tr["n_records"])},
            "note": "valid cubes that are fixed points of their voxel filter and received no point are not re-filtered "
                    "(the reference filters them again and gets the same cloud back)"},
        "cpu_baseline": {"value": cpu_value, "unit": "scans/s", "cores": 1, "kind": "port",
                         "sample": f"{n_cpu} scans of one stream, 1 thread: SR {tm['sr_ms']/n_cpu:.1f} ms + LO {tm['lo_ms']/n_cpu:.1f} ms"
                                   + (f" + LM {tm['lm_ms']/n_cpu:.1f} ms" if do_map else "") + (f" + VO {tm['vo_ms']/n_cpu:.1f} ms + image front end (OpenCV, 1 thread) {tm['fe_ms']/n_cpu:.1f} ms" if do_vo else "") + " per scan"
                                   + (f"; LM iterations executed in the last scan's passes: {pipe.lm_iterations()}" if do_map else "")},
        "clocks": clocks,
    }
    if world == 1 and args.legs == "all" and args.workload == "sr_lo_lm":
        line["vo_frontend"] = frontend_leg(V, local_rank)
    print(json.dumps(line), flush=True)


def frontend_leg(V, device_index, frames_per_call=32, reps=10):
    """Supplementary line item (not part of `value` / `e2e`): VisualOdometry::processImage on the device (vloam_vo_process_image:
    Shi-Tomasi detection, ORB description, descriptor matching) on `frames_per_call` KITTI-sized frames
    per call, wall time through the host API with the image upload inside, per-kernel CUDA-event times, and the same three OpenCV
    calls on one host thread when cv2 is importable.  Any failure is reported in the object instead of failing the bench."""
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "vo_detect_cv2.npz"))
        base = g["kitti_image"]
        rng = np.random.default_rng(5)
        shifts = [int(rng.integers(0, 300)) for _ in range(frames_per_call)]
        frames = [np.stack([np.roll(base, (0, s_ + 4 * k), axis=(0, 1)) for s_ in shifts]) for k in range(2)]
        ctx = V.Context(device=device_index)
        vo = V.VisualOdometry(ctx, batch=frames_per_call, max_points=1024, max_matches=1024)
        for k in range(4):
            vo.reset(); vo.processImage(frames[k % 2])
        ctx.enable_timing(True)
        t0 = time.perf_counter()
        for k in range(reps):
            vo.reset(); r = vo.processImage(frames[k % 2])
        dt = (time.perf_counter() - t0) / reps
        kt = ctx.kernel_timings()
        out = {"call": "vloam_vo_process_image (detection + ORB description + matching), image upload inside the timed region",
               "frames_per_call": frames_per_call, "image": list(base.shape), "ms_per_call": dt * 1e3, "frames_per_s": frames_per_call / dt,
               "kernel_ms_per_call": {k_: v_[0] / reps for k_, v_ in kt.items()}, "gpu_launches_per_call": sum(v_[1] for v_ in kt.values()) / reps,
               "keypoints_stream0": int(r["n_keypoints"][0]), "matches_stream0": int(r["n_matches"][0]), "h2d_bytes_per_call": int(frames[0].nbytes)}
        vo.close(); ctx.close()
        fe = CpuFrontEnd()
        out["cv2_ms_per_frame_1_thread"] = None
        if fe.cv2 is not None:
            n_cpu = min(frames_per_call, 4)
            t0 = time.perf_counter()
            for b_ in range(n_cpu):
                fe.prev = None
                for k in range(3):                    # (the first frame of a stream has nothing to match against)
                    fe.process(frames[k % 2][b_])
            out["cv2_ms_per_frame_1_thread"] = (time.perf_counter() - t0) / (3 * n_cpu) * 1e3
            out["cv2_version"] = fe.cv2.__version__
        return out
    except Exception as e:      # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def exchange_report(args, world, do_map):
    """Point-sharded layouts: what crosses NVLink per scan and stream.  One Levenberg-Marquardt evaluation exchanges the partial
    normal equations (21 upper J'J + 6 J'r + cost = 28 doubles); a pass runs at most max_iterations + 1 evaluations."""
    lo_eval = 2 * (4 + 1)
    lm_eval = 2 * (args.lm_iterations + 1) if do_map else 0
    if args.parallelism == "point":
        payload = 8 * 28 * 8                                  # kGnTiles tile partials per stream, summed after the all-reduce
        n_coll = lo_eval + lm_eval + 2                        # + the correspondence counts of the two odometry passes
        return {"mechanism": "ncclAllReduce between the accumulate and step launches (laser odometry and laser mapping)",
                "payload_bytes_per_stream_per_evaluation": payload, "collectives_per_scan_max": n_coll,
                "payload_bytes_per_stream_per_scan_max": (lo_eval + lm_eval) * payload + 2 * 16,
                "nvlink_bytes_per_rank_per_stream_per_scan_max": int(2 * (world - 1) / world * ((lo_eval + lm_eval) * payload + 2 * 16)),
                "note": "ring all-reduce volume 2(N-1)/N x payload per rank; every collective carries all streams of the handle"}
    payload = 28 * 8 + 8                                      # one partial + its sequence number per stream
    return {"mechanism": "peer-memory loads inside the running solve kernel (laser odometry only)",
            "payload_bytes_per_stream_per_evaluation": payload, "collectives_per_scan_max": 0,
            "payload_bytes_per_stream_per_scan_max": lo_eval * payload,
            "nvlink_bytes_per_rank_per_stream_per_scan_max": (world - 1) * lo_eval * payload,
            "note": "every rank reads the other ranks' slots: (N-1) x payload per evaluation"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="independent streams per GPU (default: 192 for sr_lo_lm, 256 for sr_lo, 128 for vloam)")
    ap.add_argument("--workload", default="sr_lo_lm", choices=["sr_lo", "sr_lo_lm", "vloam"],
                    help="sr_lo_lm (default) = BASELINE.json's metric, laserOdometry+Mapping on a 1 M-point map (configs[2], with the scan "
                         "registration that feeds it); sr_lo = configs[1]; vloam = configs[3]")
    ap.add_argument("--lm-iterations", type=int, default=0,
                    help="LM iterations per laserMapping pass (default: 5 for sr_lo_lm = configs[2]'s '10 GN iters' over the 2 passes; "
                         "4 = the reference's setting, laser_mapping.cpp:612, for the others)")
    ap.add_argument("--cpu-scans", type=int, default=200, help="scans timed for cpu_baseline (1 thread)")
    ap.add_argument("--map-points", type=int, default=1000000, help="size of the pre-built map for --workload sr_lo_lm")
    ap.add_argument("--handles", type=int, default=0,
                    help="split the batch over this many handles / CUDA streams (device leg); default: 64 streams per handle for "
                         "sr_lo_lm (3 handles), 2 handles otherwise")
    ap.add_argument("--parallelism", default="stream", choices=["stream", "point", "point-peer"],
                    help="N > 1: stream = independent streams per rank (weak scaling, headline); point = BASELINE configs[4] as worded: "
                         "every rank holds all streams and a slice of each stream's queries, the partial normal equations of every "
                         "Levenberg-Marquardt evaluation (odometry and mapping) are summed by ncclAllReduce; point-peer = the same "
                         "split for laser odometry with the sum formed inside the solve kernel through NVLink peer memory")
    ap.add_argument("--solver-mode", type=int, default=0, choices=[0, 1, 2],
                    help="vloam_lidar_params::solver_mode: 0 = by batch size, 1 = one CTA (cluster) per stream, 2 = wide accumulate + step launches")
    ap.add_argument("--graphs", type=int, default=0, choices=[0, 1],
                    help="1: drive every frame through vloam_lidar_process (one call, launch sequence replayed as a CUDA graph)")
    ap.add_argument("--legs", default="all", choices=["all", "device"], help="device: only the HBM-resident timed leg (for ncu runs)")
    args = ap.parse_args()
    if args.lm_iterations <= 0:
        args.lm_iterations = 5 if args.workload == "sr_lo_lm" else 4
    if args.batch <= 0:
        args.batch = {"sr_lo": 256, "sr_lo_lm": 192, "vloam": 128}[args.workload]
    if args.handles <= 0:
        args.handles = 3 if args.workload == "sr_lo_lm" else 2
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
