"""CPU proofs-by-test of the two claims the incremental map re-filter relies on, checked against the
oracle's pcl::VoxelGrid restatement (oracle/voxel_grid.hpp), which is what the reference runs over every valid cube after
every scan (laser_mapping.cpp:689-702):

  1. fixed point: if every centroid of a filter's output lies in the voxel it was averaged over, filtering the output again
     returns it bit for bit (so a cube that received no point can be skipped);
  2. merge: for such a cube, VoxelGrid(old ++ new) equals "patch the occupied voxels with ((0 + old) + new_1 + ...) / n,
     insert the empty ones, keep lattice order" — the restatement below is the algorithm of lm_refilter_merge.
"""
import numpy as np
import pytest


def _lattice(p, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    return np.floor(p[:, :3] * inv).astype(np.int64)


def _is_fixed_point_form(y, leaf):
    """One point per voxel, in lexicographic (z, y, x) lattice order."""
    v = _lattice(y, leaf)
    order = np.lexsort((v[:, 0], v[:, 1], v[:, 2]))
    return np.array_equal(order, np.arange(len(y))) and len(np.unique(v, axis=0)) == len(y)


def _merge(old, new, leaf):
    """old: fixed-point cube; new: points in any order.  Returns VoxelGrid(old ++ new) without sorting old."""
    vo, vn = _lattice(old, leaf), _lattice(new, leaf)
    order = np.lexsort((np.arange(len(new)), vn[:, 0], vn[:, 1], vn[:, 2]))        # by voxel, then input (stack) order
    new, vn = new[order], vn[order]
    key = lambda v: (int(v[2]), int(v[1]), int(v[0]))
    okeys = [key(v) for v in vo]
    pos = {k: i for i, k in enumerate(okeys)}
    out = {k: old[i].copy() for i, k in enumerate(okeys)}
    i = 0
    while i < len(new):
        j = i
        while j < len(new) and np.array_equal(vn[j], vn[i]):
            j += 1
        k = key(vn[i])
        s = np.zeros(4, np.float32)
        cnt = 0
        if k in pos:
            s = s + old[pos[k]]; cnt = 1
        for t in range(i, j):
            s = (s + new[t]).astype(np.float32); cnt += 1
        out[k] = (s / np.float32(cnt)).astype(np.float32)
        i = j
    keys = sorted(out)            # (z, y, x) tuples: lattice order
    return np.stack([out[k] for k in keys]).astype(np.float32)


@pytest.mark.parametrize("leaf,seed", [(0.4, 0), (0.8, 1), (0.8, 2), (0.2, 3)])
def test_fixed_point_and_merge(oracle, leaf, seed):
    rng = np.random.default_rng(seed)
    # a cube's worth of points: surfaces + clutter inside [-25, 25)^3 shifted to an arbitrary cube
    shift = np.array([150.0, -100.0, 0.0]) if seed % 2 else np.zeros(3)
    raw = np.c_[rng.uniform(-25, 25, (6000, 2)), rng.normal(0, 1.5, 6000).clip(-24, 24), rng.uniform(0, 60, 6000)]
    raw[:, :3] += shift
    y = oracle.voxel_grid(raw.astype(np.float32), leaf)
    # claim 1
    if _is_fixed_point_form(y, leaf):
        y2 = oracle.voxel_grid(y, leaf)
        assert y2.shape == y.shape and np.array_equal(y2.view(np.uint32), y.view(np.uint32))
    else:                                  # a centroid left its voxel (possible, rare): one more pass must settle or keep shrinking
        y = oracle.voxel_grid(y, leaf)
        assert _is_fixed_point_form(y, leaf)
    # claim 2: scans keep adding points, some into occupied voxels, some into new ones
    for step in range(3):
        new = np.c_[rng.uniform(-25, 25, (900, 2)), rng.normal(0, 1.5, 900).clip(-24, 24), rng.uniform(0, 60, 900)]
        new[:, :3] += shift
        new = new.astype(np.float32)
        ref = oracle.voxel_grid(np.concatenate([y, new]), leaf)
        got = _merge(y, new, leaf)
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"step {step}"
        y = ref
        if not _is_fixed_point_form(y, leaf):
            y = oracle.voxel_grid(y, leaf)
        assert _is_fixed_point_form(y, leaf)


def test_lattice_order_is_independent_of_the_bounding_box(oracle):
    """pcl::VoxelGrid keys are bounding-box relative; the ORDER they induce is not: adding far-away points must not reorder
    the voxels that were already there."""
    rng = np.random.default_rng(5)
    a = np.c_[rng.uniform(-5, 5, (500, 3)), np.zeros(500)].astype(np.float32)
    ya = oracle.voxel_grid(a, 0.8)
    far = np.array([[-24.0, -24.0, -24.0, 0], [24.0, 24.0, 24.0, 0]], np.float32)
    yb = oracle.voxel_grid(np.concatenate([a, far]), 0.8)
    core = yb[1:-1]                        # the two far voxels sort first and last
    assert np.array_equal(core.view(np.uint32), ya.view(np.uint32))


def _voxel_grid_by_runs(p, leaf):
    """pcl::VoxelGrid restated over RUNS of consecutive equal voxel keys instead of points (the next step planned for
    sr_less_flat_voxel): detect runs, order the runs by (key, first index) — a voxel's runs then sit
    side by side in index order — and chain the float sums through them point by point."""
    f = np.float32
    inv = f(1.0) / f(leaf)
    mn, mx = p[:, :3].min(0), p[:, :3].max(0)
    minb = np.floor(mn * inv).astype(np.int64)
    divb = np.floor(mx * inv).astype(np.int64) - minb + 1
    ijk = (np.floor(p[:, :3] * inv) - minb.astype(f)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * divb[0] + ijk[:, 2] * divb[0] * divb[1]
    starts = np.r_[0, np.nonzero(key[1:] != key[:-1])[0] + 1]
    ends = np.r_[starts[1:], len(p)]
    order = np.lexsort((starts, key[starts]))                # runs by (key, first index)
    out = []
    i = 0
    while i < len(order):
        k = key[starts[order[i]]]
        s = np.zeros(4, f)
        cnt = 0
        while i < len(order) and key[starts[order[i]]] == k:
            for t in range(starts[order[i]], ends[order[i]]):
                s = (s + p[t]).astype(f); cnt += 1
            i += 1
        out.append((s / f(cnt)).astype(f))
    return np.stack(out), len(starts)


def test_voxel_grid_over_runs_equals_voxel_grid_over_points(oracle, synth):
    s = synth.ScanStream(77, n_cols=1024)
    sr = oracle.scan_registration(s.scan(1))
    cloud, lab = sr.laserCloud, sr.label
    rings = cloud[:, 3].astype(int)
    runs = pts = 0
    for r in range(0, 51, 5):
        idx = np.nonzero(rings == r)[0]
        if len(idx) < 20:
            continue
        idx = idx[5:-6]
        idx = idx[lab[idx] <= 0]                             # the ring's less-flat candidates (scan_registration.cpp:424-430)
        p = np.ascontiguousarray(cloud[idx]).astype(np.float32)
        ref = oracle.voxel_grid(p, 0.2)
        got, nr = _voxel_grid_by_runs(p, 0.2)
        assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32)), r
        runs += nr; pts += len(p)
    assert pts > 3 * runs / 2                                # runs are what make it worthwhile: > 1.5 points per run here
