"""trilinear / initialise (functions.jl:207-248, voronoi_utils.jl:687-708): the oracle against a numpy
restatement written from the Julia source, and the CUDA kernel against the oracle bit for bit."""
import numpy as np
import pytest


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
