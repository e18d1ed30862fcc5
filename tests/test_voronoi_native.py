"""Native Voronoi neighbour generation against voro++'s own lists.

The committed fixtures tests/golden/grid_*.npz hold the NeighbourMatrix that the reference's voro++ driver
(rt_preprocessing/output_sites, run by tests/make_golden.py) printed for those sites — output of the reference's own
native component, so this row IS pinned by the reference.  Parity bar: per site the multiset of neighbour ids (walls -5 /
-6 included) is identical; the order inside a row is not defined by the reference.

not-gpu part: the cell geometry (voronoirt_b200/csrc/voronoi_cell.cuh, host/device code) compiled into a CPU harness
(tests/voronoi_harness.cpp, test infrastructure only).  gpu part: the same through the C ABI (vrt_voronoi_neighbours).
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import load_grid

HERE = os.path.dirname(os.path.abspath(__file__))
GRIDS = ("grid_unit300", "grid_unit1000", "grid_strat3000")


def same_sets(a, b):
    """rows of two NeighbourMatrices (n, ld) hold the same multisets"""
    if a.shape[0] != b.shape[0] or not np.array_equal(a[:, 0], b[:, 0]):
        return False
    for i in range(a.shape[0]):
        k = int(a[i, 0])
        if sorted(a[i, 1:1 + k].tolist()) != sorted(b[i, 1:1 + k].tolist()):
            return False
    return True


@pytest.fixture(scope="module")
def harness():
    d = tempfile.mkdtemp(prefix="vrt_vc_")
    so = os.path.join(d, "libvc_harness.so")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "voronoi_harness.cpp")], check=True)
    return C.CDLL(so)


def run_harness(L, pos, b, cells_per_site=0.25):
    n = pos.shape[1]
    p = np.ascontiguousarray(pos.T)                       # memory of the (3, n) column-major array
    vol = (b[1] - b[0]) * (b[3] - b[2]) * (b[5] - b[4])
    h = (vol / (n * cells_per_site)) ** (1 / 3)
    g = [max(1, int(round((b[2 * k + 1] - b[2 * k]) / h))) for k in (1, 2, 0)]
    nbr = np.zeros((n, 64), dtype=np.int64)
    st = np.zeros(n, dtype=np.int32)
    bad = L.vc_harness(C.c_int64(n), p.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), g[0], g[1], g[2],
                       nbr.ctypes.data_as(C.c_void_p), C.c_int64(64), st.ctypes.data_as(C.c_void_p))
    return nbr, st, bad


@pytest.mark.parametrize("name", GRIDS)
def test_cell_geometry_matches_voro(harness, name):
    pos, gold, b = load_grid(name)
    nbr, st, bad = run_harness(harness, pos, b)
    assert bad == 0 and not st.any()
    assert same_sets(nbr, np.asarray(gold))


def test_cell_geometry_is_independent_of_the_search_grid(harness):
    """coarser and finer search grids visit the candidates in another order: same cells"""
    pos, gold, b = load_grid("grid_unit1000")
    for cps in (0.02, 1.0, 3.0):
        nbr, st, bad = run_harness(harness, pos, b, cps)
        assert bad == 0 and same_sets(nbr, np.asarray(gold))


def test_symmetry_and_walls(harness):
    pos, gold, b = load_grid("grid_strat3000")
    nbr, st, bad = run_harness(harness, pos, b)
    n = pos.shape[1]
    sets = [set(nbr[i, 1:1 + nbr[i, 0]].tolist()) for i in range(n)]
    for i in range(n):
        for j in sets[i]:
            if j > 0:
                assert (i + 1) in sets[j - 1]             # Voronoi adjacency is symmetric
            else:
                assert j in (-5, -6)
    assert all(len(s) >= 4 for s in sets)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GRIDS)
def test_gpu_neighbours_match_voro(name):
    import voronoirt_b200 as V
    pos, gold, b = load_grid(name)
    nbr = V.voronoi_neighbours(pos, *b)
    assert nbr.shape[1] == int(np.asarray(gold)[:, 0].max()) + 1
    assert same_sets(np.asarray(nbr), np.asarray(gold))


@pytest.mark.gpu
def test_gpu_neighbours_drive_the_solver_like_voro_lists():
    """the generated matrix is a drop-in input of read_cell: same layers as with voro++'s matrix (BFS layers depend on the
    sets only), and a formal solution runs on it"""
    import voronoirt_b200 as V
    pos, gold, b = load_grid("grid_strat3000")
    n = pos.shape[1]
    nbr = V.voronoi_neighbours(pos, *b)
    cell_a = V.read_cell(nbr, n, pos, b[2], b[3], b[4], b[5])
    cell_b = V.read_cell(gold, n, pos, b[2], b[3], b[4], b[5])
    assert np.array_equal(cell_a[3], cell_b[3]) and np.array_equal(cell_a[4], cell_b[4])      # layers_up / layers_down
    assert np.array_equal(cell_a[5], cell_b[5]) and np.array_equal(cell_a[6], cell_b[6])      # perm_up / perm_down


def brute_force_nn(pos, q):
    """argmin over the sites of (dz^2 + dx^2) + dy^2, first minimum; pos (3, n), q (3, m)"""
    d = (pos.T[None, :, :] - q.T[:, None, :]) ** 2
    d2 = (d[:, :, 0] + d[:, :, 1]) + d[:, :, 2]
    return d2.argmin(axis=1) + 1, np.sqrt(d2.min(axis=1))


def queries(b, m, rng):
    """points inside the box, on its faces and a little outside"""
    q = np.stack([rng.uniform(b[0], b[1], m), rng.uniform(b[2], b[3], m), rng.uniform(b[4], b[5], m)])
    q[0, :50] = b[0]; q[1, 50:100] = b[3]; q[2, 100:150] = b[5]
    q[:, 150:200] += (np.array([b[1] - b[0], b[3] - b[2], b[5] - b[4]]) * 0.05)[:, None]
    return np.asfortranarray(q)


def test_nearest_site_core_equals_brute_force(harness):
    pos, gold, b = load_grid("grid_strat3000")
    q = queries(b, 4000, np.random.default_rng(3))
    idx = np.zeros(q.shape[1], dtype=np.int64)
    d2 = np.zeros(q.shape[1])
    p = np.ascontiguousarray(pos.T)
    qq = np.ascontiguousarray(q.T)
    harness.vc_nn_harness(C.c_int64(pos.shape[1]), p.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 8, 8, 13, C.c_int64(q.shape[1]),
                          qq.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p))
    ref, dist = brute_force_nn(pos, q)
    assert np.array_equal(idx, ref) and np.array_equal(np.sqrt(d2), dist)


@pytest.mark.gpu
def test_gpu_nearest_site_and_raster():
    import voronoirt_b200 as V
    pos, gold, b = load_grid("grid_strat3000")
    q = queries(b, 4000, np.random.default_rng(4))
    idx, dist = V.nearest_site(pos, b, q)
    ref, rdist = brute_force_nn(pos, q)
    assert np.array_equal(idx, ref) and np.array_equal(dist, rdist)
    # Voronoi_to_Raster: rasters of per-site fields (voronoi_utils.jl:436-455)
    n = pos.shape[1]
    cell = V.read_cell(gold, n, pos, b[2], b[3], b[4], b[5])
    rng = np.random.default_rng(5)
    T, S = rng.uniform(4e3, 1e4, n), rng.uniform(0, 1, (7, n))
    sites = V.VoronoiSites(*cell, T, T, T, T, T, T, b[0], b[1], b[2], b[3], b[4], b[5], n)
    z, x, y = np.linspace(b[0], b[1], 9), np.linspace(b[2], b[3], 6), np.linspace(b[4], b[5], 5)
    ridx, Tg, Sg = V.Voronoi_to_Raster(sites, z, x, y, T, S)
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    want, _ = brute_force_nn(pos, np.stack([Z.ravel(), X.ravel(), Y.ravel()]))
    assert np.array_equal(ridx.ravel(), want)
    assert np.array_equal(Tg, T[ridx - 1]) and np.array_equal(Sg, S[:, ridx - 1]) and Sg.shape == (7, 9, 6, 5)


def brute_force_knn(pos, q, k):
    d = (pos.T[None, :, :] - q.T[:, None, :]) ** 2
    d2 = (d[:, :, 0] + d[:, :, 1]) + d[:, :, 2]
    idx = np.argsort(d2, axis=1, kind="stable")[:, :k]
    return (idx + 1).T, np.sqrt(np.take_along_axis(d2, idx, axis=1)).T


def test_k_nearest_sites_core_equals_brute_force(harness):
    pos, gold, b = load_grid("grid_strat3000")
    q = queries(b, 3000, np.random.default_rng(8))
    k = 4
    idx = np.zeros((q.shape[1], k), dtype=np.int64)
    d2 = np.zeros((q.shape[1], k))
    p = np.ascontiguousarray(pos.T)
    qq = np.ascontiguousarray(q.T)
    harness.vc_knn_harness(C.c_int64(pos.shape[1]), p.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 8, 8, 13, C.c_int64(q.shape[1]),
                           qq.ctypes.data_as(C.c_void_p), k, idx.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p))
    ref, dist = brute_force_knn(pos, q, k)
    assert np.array_equal(idx.T, ref) and np.array_equal(np.sqrt(d2).T, dist)


@pytest.mark.gpu
def test_gpu_k_nearest_and_inverse_distance_raster():
    import voronoirt_b200 as V
    pos, gold, b = load_grid("grid_strat3000")
    q = queries(b, 3000, np.random.default_rng(9))
    for k in (1, 2, 5):
        idx, dist = V.nearest_sites(pos, b, q, k)
        ref, rdist = brute_force_knn(pos, q, k)
        assert np.array_equal(idx, ref) and np.array_equal(dist, rdist)
    # Voronoi_to_Raster_inv_dist (voronoi_utils.jl:773-816, p = 1, n_k = 2) for the populations (n, 3)
    n = pos.shape[1]
    cell = V.read_cell(gold, n, pos, b[2], b[3], b[4], b[5])
    rng = np.random.default_rng(10)
    T = rng.uniform(4e3, 1e4, n)
    pops = rng.uniform(1.0, 2.0, (n, 3))
    sites = V.VoronoiSites(*cell, T, T, T, T, T, T, b[0], b[1], b[2], b[3], b[4], b[5], n)
    z, x, y = np.linspace(b[0], b[1], 7), np.linspace(b[2], b[3], 5), np.linspace(b[4], b[5], 4)
    grid = V.Voronoi_to_Raster_inv_dist(sites, z, x, y, pops)
    assert grid.shape == (7, 5, 4, 3)
    for (kk, ii, jj) in ((0, 0, 0), (3, 2, 1), (6, 4, 3)):
        d = np.sqrt(((pos.T - np.array([z[kk], x[ii], y[jj]])) ** 2).sum(axis=1))
        o = np.argsort(d, kind="stable")[:2]
        inv = 1.0 / d[o]
        want = (pops[o[0]] * inv[0] + pops[o[1]] * inv[1]) / (inv[0] + inv[1])
        assert np.allclose(grid[kk, ii, jj], want, rtol=1e-14)


def tiny_cases():
    d = np.load(os.path.join(HERE, "golden", "tiny_voro.npz"))
    return [(np.asfortranarray(d[f"pos_{c}"]), d[f"nbr_{c}"].astype(np.int64)) for c in range(int(d["n_cases"]))]


def test_nearly_empty_periodic_boxes(harness):
    """1..40 sites in the unit box (tests/make_tiny_golden.py): cells bounded by periodic images, also of the site itself
    (voro++ then lists the site's own id); bisectors through cell edges must not become faces"""
    b = np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0])
    for pos, gold in tiny_cases():
        nbr, st, bad = run_harness(harness, pos, b)
        assert bad == 0 and same_sets(nbr, gold), pos.shape


@pytest.mark.gpu
def test_gpu_nearly_empty_periodic_boxes():
    import voronoirt_b200 as V
    for pos, gold in tiny_cases():
        nbr = V.voronoi_neighbours(pos, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
        assert same_sets(np.asarray(nbr), gold), pos.shape


def _lattice(m):
    ax = (np.arange(m) + 0.5) / m
    Z, X, Y = np.meshgrid(ax, ax, ax, indexing="ij")
    pos = np.asfortranarray(np.stack([Z.ravel(), X.ravel(), Y.ravel()]))
    idx = np.arange(m ** 3).reshape(m, m, m)                       # site id - 1 of lattice point (iz, ix, iy)
    want = []
    for iz in range(m):
        for ix in range(m):
            for iy in range(m):
                row = [idx[iz, (ix + 1) % m, iy] + 1, idx[iz, (ix - 1) % m, iy] + 1, idx[iz, ix, (iy + 1) % m] + 1, idx[iz, ix, (iy - 1) % m] + 1]
                row.append(idx[iz - 1, ix, iy] + 1 if iz > 0 else -5)
                row.append(idx[iz + 1, ix, iy] + 1 if iz < m - 1 else -6)
                want.append(sorted(row))
    return pos, want


def test_cubic_lattice_has_six_faces_per_cell(harness):
    """exactly degenerate input: the bisectors towards edge and corner neighbours only touch the cell; those zero-area faces
    are dropped, as voro++ drops them (6 faces per cell: periodic in x and y, walls in z)"""
    pos, want = _lattice(6)
    nbr, st, bad = run_harness(harness, pos, np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0]))
    assert bad == 0 and np.all(nbr[:, 0] == 6)
    assert [sorted(nbr[i, 1:7].tolist()) for i in range(nbr.shape[0])] == want


@pytest.mark.gpu
def test_gpu_cubic_lattice_has_six_faces_per_cell():
    import voronoirt_b200 as V
    pos, want = _lattice(8)
    nbr = np.asarray(V.voronoi_neighbours(pos, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0))
    assert np.all(nbr[:, 0] == 6)
    assert [sorted(nbr[i, 1:7].tolist()) for i in range(nbr.shape[0])] == want
