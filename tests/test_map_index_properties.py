"""CPU check of the exactness argument behind the per-cube column index of laserMapping (lm_knn):
every map point within 1 m of a query (float32 squared distance < 1, the reference's acceptance test at
laser_mapping.cpp:479 / :547) lies in the 3 x 3 block of 1.001 m columns around the query in one of the cubes the query's
+-1.001 m box touches.  The index arithmetic is restated here in numpy with the kernel's float32 / double expressions
(cube_coord, cube_min_coord, cube_cell[_clamped] in lm_kernels.cu); the candidate set it yields is compared with a brute-force
scan.  Also covers the z-binned refinement (4 m bins inside a column) planned as the next step."""
import numpy as np
import pytest

F = np.float32
CELL_INV = F(1.0) / F(1.001)
ZBIN_INV = F(1.0) / F(4.0)


def cube_coord(v, cen):                      # laser_mapping.cpp:207-216 / 643-652, double arithmetic
    v = np.asarray(v, np.float64)
    c = ((v + 25.0) / 50.0).astype(np.int64) + cen          # truncation toward zero like the C cast
    return np.where(v + 25.0 < 0, c - 1, c)


def cube_min(idx, cen):
    return F((idx - cen) * 50.0 - 25.0)


def cell(v, mn):
    return np.floor((np.asarray(v, F) - F(mn)) * CELL_INV).astype(np.int64)


def zbin(z, mnz):
    return np.clip(np.floor((np.asarray(z, F) - F(mnz)) * ZBIN_INV).astype(np.int64), 0, 12)


def sqdist_f(a, b):                          # ((dx*dx + dy*dy) + dz*dz) in float32
    d = (a.astype(F) - b.astype(F)).astype(F)
    return ((d[..., 0] * d[..., 0]).astype(F) + (d[..., 1] * d[..., 1]).astype(F)).astype(F) + (d[..., 2] * d[..., 2]).astype(F)


@pytest.mark.parametrize("seed,use_zbins", [(0, False), (1, False), (2, True), (3, True)])
def test_column_block_contains_every_point_within_one_metre(seed, use_zbins):
    rng = np.random.default_rng(seed)
    cen = (10, 10, 5)
    # a dense cloud around a corner where eight cubes meet, so that many queries sit at cube borders
    corner = np.array([25.0, 25.0, 25.0])
    pts = (corner + rng.uniform(-6, 6, (60000, 3))).astype(F)
    # file every point under its cube and its column (and z-bin)
    ci, cj, ck = (cube_coord(pts[:, a], cen[a]) for a in range(3))
    mnx = np.array([cube_min(i, cen[0]) for i in ci]); mny = np.array([cube_min(j, cen[1]) for j in cj]); mnz = np.array([cube_min(k, cen[2]) for k in ck])
    px = np.clip(np.floor((pts[:, 0] - mnx) * CELL_INV).astype(np.int64), 0, 49)
    py = np.clip(np.floor((pts[:, 1] - mny) * CELL_INV).astype(np.int64), 0, 49)
    pz = np.clip(np.floor((pts[:, 2] - mnz) * ZBIN_INV).astype(np.int64), 0, 12)
    index = {}
    for n, key in enumerate(zip(ci.tolist(), cj.tolist(), ck.tolist(), px.tolist(), py.tolist(), pz.tolist() if use_zbins else [0] * len(pts))):
        index.setdefault(key, []).append(n)
    queries = (corner + rng.uniform(-5, 5, (400, 3))).astype(F)
    queries[:40] = (corner + rng.uniform(-1.2, 1.2, (40, 3))).astype(F)            # right at the corner
    queries[40:60, 0] = F(25.0)                                                    # exactly on a cube face
    checked = 0
    for q in queries:
        cand = set()
        r = {}
        for a, c in enumerate(cen):
            r[a] = range(int(cube_coord(np.float64(q[a]) - 1.001, c)), int(cube_coord(np.float64(q[a]) + 1.001, c)) + 1)
        for k in r[2]:
            for j in r[1]:
                for i in r[0]:
                    qx = int(cell(q[0], cube_min(i, cen[0]))); qy = int(cell(q[1], cube_min(j, cen[1])))
                    mz = cube_min(k, cen[2])
                    zs = range(int(zbin(q[2] - F(1.001), mz)), int(zbin(q[2] + F(1.001), mz)) + 1) if use_zbins else [0]
                    for y in range(max(qy - 1, 0), min(qy + 1, 49) + 1):
                        for x in range(max(qx - 1, 0), min(qx + 1, 49) + 1):
                            for z in zs:
                                cand.update(index.get((i, j, k, x, y, z), ()))
        d = sqdist_f(pts, q[None, :])
        near = set(np.nonzero(d < F(1.0))[0].tolist())
        assert near <= cand, (q, len(near - cand))
        checked += len(near)
    assert checked > 20000


def test_lo_grid_safe_radius_bounds_every_unvisited_point():
    """laserOdometry's column grid (lo_kernels.cu: cell_coord, grid_safe_radius): after the (2k+1)^2 block of columns around
    the query has been visited, every point filed under another column is farther than the safe radius R_k, in the float32
    distance the search compares (so stopping at best <= R_k^2 is exact, including the index tie-break)."""
    rng = np.random.default_rng(7)
    pts = np.c_[rng.uniform(-60, 60, (40000, 2)), rng.uniform(-3, 8, 40000)].astype(F)
    minx, miny = F(pts[:, 0].min()), F(pts[:, 1].min())
    c = F(1.0)
    inv_c = F(1.0) / c
    nx = int(np.floor((pts[:, 0].max() - minx) / c)) + 1
    ny = int(np.floor((pts[:, 1].max() - miny) / c)) + 1
    ix = np.clip(np.floor((pts[:, 0] - minx) * inv_c).astype(np.int64), 0, nx - 1)
    iy = np.clip(np.floor((pts[:, 1] - miny) * inv_c).astype(np.int64), 0, ny - 1)
    queries = np.c_[rng.uniform(-65, 65, (300, 2)), rng.uniform(-3, 8, 300)].astype(F)       # some outside the grid
    queries[:50, :2] = (np.round(queries[:50, :2]) - (minx % F(1.0))).astype(F)                # near column borders
    worst = np.inf
    for q in queries:
        qx = int(np.floor((q[0] - minx) * inv_c)); qy = int(np.floor((q[1] - miny) * inv_c))
        d = sqdist_f(pts, q[None, :])
        for k in range(1, 6):
            xl = F(q[0] - (minx + F(qx - k) * c)); xr = F((minx + F(qx + k + 1) * c) - q[0])
            yl = F(q[1] - (miny + F(qy - k) * c)); yr = F((miny + F(qy + k + 1) * c) - q[1])
            R = F(min(xl, xr, yl, yr) * (F(1.0) - F(1e-5)) - F(1e-4))
            if R <= 0:
                continue
            outside = (np.abs(ix - qx) > k) | (np.abs(iy - qy) > k)
            if outside.any():
                m = d[outside].min()
                assert m > F(R * R), (q, k, m, R)
                worst = min(worst, float(m) - float(R * R))
    assert np.isfinite(worst)


def test_lo_ring_window_interval_equals_the_literal_walks():
    """lo_associate's ring-window pass (lo_kernels.cu, "phase 2"): the reference walks the ring-major target cloud away from
    the closest point, forwards until the first point with int(intensity) > id + 2.5 and backwards until the first with
    int(intensity) < id - 2.5 (laser_odometry.cpp:279-324 / 368-417).  With int(intensity) equal to the point's true ring R
    or R - 1 (negative relTime), the indices the two walks visit are exactly [lo_j, hi_j) minus the closest point, with
    lo_j / hi_j from the true ring offsets and the per-ring tables firstFull / lastLow (GridHeader)."""
    rng = np.random.default_rng(11)
    R = 64
    for trial in range(30):
        counts = rng.integers(0, 40, R)
        counts[rng.integers(0, R, 6)] = 0                                   # some empty rings
        ring_start = np.r_[0, np.cumsum(counts)]                            # [R + 1]
        n = int(ring_start[-1])
        true_ring = np.repeat(np.arange(R), counts)
        low = (rng.random(n) < 0.3) & (true_ring > 0)                       # int(intensity) == R - 1 for these
        rid = true_ring - low.astype(int)
        first_full = np.full(R + 1, np.iinfo(np.int32).max, np.int64)
        last_low = np.full(R + 1, -1, np.int64)
        for j in range(n):
            r = true_ring[j]
            if rid[j] == r:
                first_full[r] = min(first_full[r], j)
            else:
                last_low[r] = max(last_low[r], j)
        rs = np.r_[ring_start, n]                                           # ringStart[64] = ringStart[65] = n
        for closest in rng.integers(0, n, 60):
            idc = int(rid[closest])
            fwd = []
            for j in range(closest + 1, n):
                if rid[j] > idc + 2.5:
                    break
                fwd.append(j)
            bwd = []
            for j in range(closest - 1, -1, -1):
                if rid[j] < idc - 2.5:
                    break
                bwd.append(j)
            hi_j, lo_j = n, 0
            if idc + 3 <= R - 1:
                hi_j = min(first_full[idc + 3], rs[min(idc + 4, R)])
            if idc - 2 >= 0:
                lo_j = max(last_low[idc - 2], rs[idc - 2] - 1) + 1
            want = set(fwd) | set(bwd)
            got = set(range(int(lo_j), int(hi_j))) - {int(closest)}
            assert got == want, (trial, closest, idc, sorted(got ^ want)[:5])
