"""trilinear / initialise (functions.jl:207-248, voronoi_utils.jl:687-708): the oracle against a numpy
restatement written from the Julia source, and the CUDA kernel against the oracle bit for bit."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    """tests/voronoi_harness.cpp: the host/device set-up code (voronoi_cell.cuh, sampling.cuh) compiled for the CPU"""
    d = tempfile.mkdtemp(prefix="vrt_vc_")
    so = os.path.join(d, "libvc_harness.so")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "voronoi_harness.cpp")], check=True)
    L = C.CDLL(so)
    L.vc_sample_harness.restype = C.c_int64
    L.vc_trilinear_harness.restype = C.c_int64
    return L


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def problem(rng, nz=9, nx=7, ny=8, n=500):
    z = np.cumsum(rng.uniform(0.5, 2.0, nz))
    x = np.cumsum(rng.uniform(0.5, 2.0, nx))
    y = np.cumsum(rng.uniform(0.5, 2.0, ny))
    vals = np.asfortranarray(rng.normal(size=(nz, nx, ny)) * 10.0 ** rng.uniform(-3, 6, (nz, nx, ny)))
    pos = np.asfortranarray(np.stack([rng.uniform(z[0], z[-1], n), rng.uniform(x[0], x[-1], n), rng.uniform(y[0], y[-1], n)]))
    pos[:, 0] = (z[3], x[2], y[5])                       # exactly on grid points: lower corner is the previous point
    pos[:, 1] = (z[-1], x[-1], y[-1])                    # upper corner of the box
    return z, x, y, vals, pos


def numpy_trilinear(z, x, y, vals, pos):
    out = np.zeros(pos.shape[1])
    for k in range(pos.shape[1]):
        zm, xm, ym = pos[:, k]
        iz, ix, iy = (np.searchsorted(a, v, side="left") - 1 for a, v in ((z, zm), (x, xm), (y, ym)))     # searchsortedfirst - 1, 0-based
        x_d = (xm - x[ix]) / (x[ix + 1] - x[ix]); y_d = (ym - y[iy]) / (y[iy + 1] - y[iy]); z_d = (zm - z[iz]) / (z[iz + 1] - z[iz])
        c00 = vals[iz, ix, iy] * (1 - x_d) + vals[iz, ix + 1, iy] * x_d
        c01 = vals[iz + 1, ix, iy] * (1 - x_d) + vals[iz + 1, ix + 1, iy] * x_d
        c10 = vals[iz, ix, iy + 1] * (1 - x_d) + vals[iz, ix + 1, iy + 1] * x_d
        c11 = vals[iz + 1, ix, iy + 1] * (1 - x_d) + vals[iz + 1, ix + 1, iy + 1] * x_d
        c0 = c00 * (1 - y_d) + c10 * y_d
        c1 = c01 * (1 - y_d) + c11 * y_d
        out[k] = c0 * (1 - z_d) + c1 * z_d
    return out


def test_oracle_trilinear_equals_the_julia_expressions(oracle):
    z, x, y, vals, pos = problem(np.random.default_rng(1))
    out, bad = oracle.trilinear(z, x, y, vals, pos)
    assert bad == 0 and np.array_equal(out, numpy_trilinear(z, x, y, vals, pos))
    # a field that is linear in the coordinates is reproduced
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    lin, _ = oracle.trilinear(z, x, y, 2 * Z - 3 * X + 0.5 * Y + 1, pos)
    assert np.allclose(lin, 2 * pos[0] - 3 * pos[1] + 0.5 * pos[2] + 1, rtol=1e-13, atol=1e-12)
    # outside the axes Julia throws: flagged, NaN
    p2 = pos.copy(); p2[0, 5] = z[0] - 1.0; p2[1, 6] = x[-1] + 1.0; p2[2, 7] = y[0]       # y[0] itself: searchsortedfirst - 1 = 0 in Julia
    out2, bad2 = oracle.trilinear(z, x, y, vals, p2)
    assert bad2 == 3 and np.isnan(out2[[5, 6, 7]]).all()


@pytest.mark.gpu
def test_gpu_trilinear_bit_exact(oracle):
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib
    z, x, y, vals, pos = problem(np.random.default_rng(2), n=20000)
    atm = V.Atmosphere(z, x, y, *(vals * s for s in (1.0, 2.0, 3.0, -1.0, 0.5, 7.0)))
    ref, bad = oracle.trilinear(z, x, y, vals, pos)
    assert bad == 0
    assert np.array_equal(V.trilinear(pos, atm, vals), ref)
    six = V.initialise(pos, atm)
    for s, got in zip((1.0, 2.0, 3.0, -1.0, 0.5, 7.0), six):
        assert np.array_equal(got, oracle.trilinear(z, x, y, vals * s, pos)[0])
    p2 = pos.copy(); p2[0, 5] = z[0] - 1.0
    with pytest.raises(_lib.VRTError, match="outside"):
        V.trilinear(p2, atm, vals)


def harness_sample(L, n, z, x, y, q, seed):
    pos = np.zeros((n, 3))
    trials = L.vc_sample_harness(C.c_int64(n), C.c_int64(len(z)), C.c_int64(len(x)), C.c_int64(len(y)), ptr(z), ptr(x), ptr(y), ptr(q),
                                 C.c_uint64(seed), C.c_double(q.min()), C.c_double(q.max() - q.min()), ptr(pos))
    return np.asfortranarray(pos.T), trials


def density_box(rng):
    z = np.linspace(0.0, 4.0, 17)
    x = np.linspace(-1.0, 1.0, 9)
    y = np.linspace(0.0, 3.0, 7)
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    q = np.asfortranarray(np.exp(-Z) * (1.5 + np.sin(2 * X) * np.cos(Y)))           # strongly stratified in z
    return z, x, y, q


def test_philox_known_answers(harness):
    """Random123's kat_vectors for philox4x32, 10 rounds"""
    for ctr, key, want in (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
                           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))):
        a = (C.c_uint32 * 4)(*ctr)
        harness.vc_philox(a, C.c_uint32(key[0]), C.c_uint32(key[1]))
        assert tuple(a) == want


def test_shared_trilinear_equals_oracle(harness, oracle):
    z, x, y, vals, pos = problem(np.random.default_rng(6))
    out = np.zeros(pos.shape[1])
    bad = harness.vc_trilinear_harness(C.c_int64(len(z)), C.c_int64(len(x)), C.c_int64(len(y)), ptr(z), ptr(x), ptr(y), ptr(vals),
                                       C.c_int64(pos.shape[1]), ptr(np.ascontiguousarray(pos.T)), ptr(out))
    ref, rbad = oracle.trilinear(z, x, y, vals, pos)
    assert bad == rbad == 0 and np.array_equal(out, ref)


def test_rejection_sampling_follows_the_density(harness, oracle):
    """functions.jl:100-118: accepted points are distributed with density proportional to q - q_min inside the box"""
    z, x, y, q = density_box(np.random.default_rng(0))
    n = 40000
    pos, trials = harness_sample(harness, n, z, x, y, q, 2022)
    assert trials > n and (pos[0] > z[0]).all() and (pos[0] <= z[-1]).all() and (pos[1] > x[0]).all() and (pos[2] <= y[-1]).all()
    pos2, _ = harness_sample(harness, n, z, x, y, q, 2022)
    pos3, _ = harness_sample(harness, n, z, x, y, q, 7)
    assert np.array_equal(pos, pos2) and not np.array_equal(pos, pos3)                 # reproducible, seed-dependent
    # expected marginal along z: integral over x, y of (q_interp - q_min), by fine quadrature of the trilinear interpolant
    zz = np.linspace(z[0], z[-1], 257)[1:]
    xx = np.linspace(x[0], x[-1], 41)[1:]
    yy = np.linspace(y[0], y[-1], 31)[1:]
    Z, X, Y = np.meshgrid(zz, xx, yy, indexing="ij")
    pts = np.asfortranarray(np.stack([Z.ravel(), X.ravel(), Y.ravel()]))
    dens = (oracle.trilinear(z, x, y, q, pts)[0] - q.min()).reshape(Z.shape)
    marg = dens.sum(axis=(1, 2))
    edges = np.linspace(z[0], z[-1], 17)
    expected = np.array([marg[(zz > a) & (zz <= b)].sum() for a, b in zip(edges[:-1], edges[1:])])
    expected *= n / expected.sum()
    got, _ = np.histogram(pos[0], bins=edges)
    chi2 = ((got - expected) ** 2 / np.maximum(expected, 1.0)).sum()
    assert chi2 < 60.0, (chi2, got, expected)                                          # 15 degrees of freedom; quadrature error included
    # acceptance rate = mean(q - q_min) / (q_max - q_min)
    rate = dens.mean() / (q.max() - q.min())
    assert abs(n / trials - rate) < 0.02 * rate + 0.005


@pytest.mark.gpu
def test_gpu_rejection_sampling_equals_the_host_evaluation(harness):
    import voronoirt_b200 as V
    z, x, y, q = density_box(np.random.default_rng(0))
    atm = V.Atmosphere(z, x, y)
    n = 20000
    pos = V.rejection_sampling(n, atm, q, seed=99)
    ref, trials = harness_sample(harness, n, z, x, y, q, 99)
    assert np.array_equal(pos, ref)
    assert abs(V.rejection_sampling.mean_trials - trials / n) < 1e-9


def numpy_nearest_corner(z, x, y, vals, pos):
    """voronoi_utils.jl:727-760: corners in the order (idz, idx, idy), (idz, idx, idy+1), (idz, idx+1, idy), ...; argmin of euclidean"""
    out = np.zeros(pos.shape[1])
    for k in range(pos.shape[1]):
        zm, xm, ym = pos[:, k]
        iz, ix, iy = (np.searchsorted(a, v, side="left") - 1 for a, v in ((z, zm), (x, xm), (y, ym)))
        best = None
        for a in (0, 1):
            for b in (0, 1):
                for c in (0, 1):
                    d = np.sqrt((z[iz + a] - zm) ** 2 + (x[ix + b] - xm) ** 2 + (y[iy + c] - ym) ** 2)
                    if best is None or d < best[0]:
                        best = (d, vals[iz + a, ix + b, iy + c])
        out[k] = best[1]
    return out


def test_nearest_corner_core(harness):
    z, x, y, vals, pos = problem(np.random.default_rng(12), n=1500)
    harness.vc_corner_harness.restype = C.c_int64
    out = np.zeros(pos.shape[1])
    bad = harness.vc_corner_harness(C.c_int64(len(z)), C.c_int64(len(x)), C.c_int64(len(y)), ptr(z), ptr(x), ptr(y), ptr(vals),
                                    C.c_int64(pos.shape[1]), ptr(np.ascontiguousarray(pos.T)), ptr(out))
    assert bad == 0 and np.array_equal(out, numpy_nearest_corner(z, x, y, vals, pos))


@pytest.mark.gpu
def test_gpu_initialiseII():
    import voronoirt_b200 as V
    z, x, y, vals, pos = problem(np.random.default_rng(13), n=3000)
    atm = V.Atmosphere(z, x, y, *(vals * s for s in (1.0, 2.0, 3.0, -1.0, 0.5, 7.0)))
    ref = numpy_nearest_corner(z, x, y, vals, pos)
    for s, got in zip((1.0, 2.0, 3.0, -1.0, 0.5, 7.0), V.initialiseII(pos, atm)):
        assert np.array_equal(got, ref * s)
