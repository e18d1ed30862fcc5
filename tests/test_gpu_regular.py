"""Parity of the CUDA regular-grid short-characteristics solver (vrt_regular_formal_solve) through the
C ABI: against the reference's own golden searchlight vectors (tests/golden/ref_I_*_regular.npy, see
test_regular_golden.py for their provenance) and against the oracle on seeded inputs.

Tolerance: I within 1e-9 relative (BASELINE.json north_star), plus 1e-13 * max|I| absolute for values that are sums of
terms of either sign in the last bits; the branch taken per plane (an index decision) must be identical.
"""
import os

import numpy as np
import pytest

from test_regular_golden import GOLD, kvec, searchlight_regular_inputs

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
QUAD = np.loadtxt(os.path.join(HERE, "..", "voronoirt_b200", "quadratures", "ul7n12.dat"))


def close(a, b, rel=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max() if b.size else 0.0
    return np.all(np.abs(a - b) <= rel * np.abs(b) + 1e-13 * scale)


def random_box(rng, nz, nx, ny, nlam, stretch=True):
    """non-uniform z (so that the branch changes from plane to plane), uniform x, y with ghost columns, opacities
    spanning the three linear_weights branches (dtau < 5e-4, in between, > 50)"""
    z = np.cumsum(rng.uniform(0.3, 3.0, nz)) if stretch else np.linspace(0, 1.0, nz) * nz
    x = (np.arange(nx) - 1) * 1.1
    y = (np.arange(ny) - 1) * 0.9
    shape = (nlam, nz, nx, ny)
    S = np.asfortranarray(rng.uniform(0.1, 2.0, shape))
    alpha = np.asfortranarray(10.0 ** rng.uniform(-5, 2, shape))
    I0 = np.asfortranarray(rng.uniform(0.0, 1.0, (nlam, nx, ny)))
    # periodic ghost columns as the reference's Atmosphere has them
    for a in (S, alpha):
        a[:, :, 0, :] = a[:, :, -2, :]; a[:, :, -1, :] = a[:, :, 1, :]
        a[:, :, :, 0] = a[:, :, :, -2]; a[:, :, :, -1] = a[:, :, :, 1]
    I0[:, 0, :] = I0[:, -2, :]; I0[:, -1, :] = I0[:, 1, :]
    I0[:, :, 0] = I0[:, :, -2]; I0[:, :, -1] = I0[:, :, 1]
    return z, x, y, S, alpha, I0


def solve(V, z, x, y, k, down, S, I0, alpha, n_sweeps=3):
    fn = V.short_characteristics_down if down else V.short_characteristics_up
    return fn(k, S, I0, alpha, V.Atmosphere(z, x, y), n_sweeps, return_branches=True)


@pytest.mark.parametrize("theta,phi,down,fname", [
    (160.0, 45.0, 0, "ref_I_160_45_regular.npy"),
    (20.0, 195.0, 1, "ref_I_20_15_regular.npy"),
])
def test_reference_golden_searchlight(theta, phi, down, fname):
    import voronoirt_b200 as V
    ax, S, alpha, I0 = searchlight_regular_inputs()
    I, branch = solve(V, ax, ax, ax, kvec(theta, phi), down, S, I0, alpha)
    out = (I[0] if down else I[-1])[1:-1, 1:-1]
    gold = np.load(os.path.join(GOLD, fname))
    assert np.abs(out - gold).max() <= 1e-13
    assert abs(out.sum() - 80.0) < 1e-10                 # compare_searchlight.jl:209
    assert set(branch.tolist()) == {0, 1}


def test_all_directions_all_branches_vs_oracle(oracle):
    import voronoirt_b200 as V
    rng = np.random.default_rng(7)
    z, x, y, S, alpha, I0 = random_box(rng, 14, 19, 23, 1)
    seen = set()
    for _, theta, phi in QUAD:
        down = int(theta < 90)
        k = kvec(theta, phi)
        ref, ref_branch = oracle.short_characteristics(z, x, y, k, down, S[0], I0[0], alpha[0])
        I, branch = solve(V, z, x, y, k, down, S[0], I0[0], alpha[0])
        assert np.array_equal(branch, ref_branch)
        assert close(I, ref), (theta, phi, np.abs(I - ref).max())
        seen |= set(branch.tolist())
    assert seen == {0, 1, 2, 3}


@pytest.mark.parametrize("n_sweeps", [1, 2, 5])
def test_sweep_counts_vs_oracle(oracle, n_sweeps):
    import voronoirt_b200 as V
    rng = np.random.default_rng(11)
    z, x, y, S, alpha, I0 = random_box(rng, 9, 12, 17, 1, stretch=False)
    alpha *= 1e-2                                        # thin enough for the carried row to matter across sweeps
    for theta, phi in ((109.7, 193.6), (78.2, 55.4), (70.3, 346.4), (114.9, 80.2)):
        down = int(theta < 90)
        k = kvec(theta, phi)
        ref, ref_branch = oracle.short_characteristics(z, x, y, k, down, S[0], I0[0], alpha[0], n_sweeps)
        I, branch = solve(V, z, x, y, k, down, S[0], I0[0], alpha[0], n_sweeps)
        assert np.array_equal(branch, ref_branch) and set(branch.tolist()) - {0, 1}
        assert close(I, ref), (theta, phi, np.abs(I - ref).max())


def test_wavelength_batch_chunks_and_device_pointers(oracle, monkeypatch):
    import torch
    import voronoirt_b200 as V
    rng = np.random.default_rng(3)
    nlam = 5
    z, x, y, S, alpha, I0 = random_box(rng, 10, 35, 14, nlam)
    for theta, phi in ((152.7, 315.5), (67.2, 155.8), (101.8, 235.4)):
        down = int(theta < 90)
        k = kvec(theta, phi)
        ref = np.stack([oracle.short_characteristics(z, x, y, k, down, S[l], I0[l], alpha[l])[0] for l in range(nlam)])
        I, _ = solve(V, z, x, y, k, down, S, I0, alpha)
        assert I.shape == S.shape and close(I, ref)
        monkeypatch.setenv("VRT_REG_LAM_CHUNK", "2")     # 2 + 2 + 1 wavelengths, strided staging copies
        I2, _ = solve(V, z, x, y, k, down, S, I0, alpha)
        monkeypatch.delenv("VRT_REG_LAM_CHUNK")
        assert np.array_equal(I, I2)
        # device-resident arrays in the same (Julia) memory layout
        dev = [torch.from_numpy(np.ascontiguousarray(a.transpose(*range(a.ndim)[::-1]))).cuda() for a in (S, I0, alpha)]
        out = torch.zeros_like(dev[0])
        atm = V.Atmosphere(z, x, y)
        from voronoirt_b200 import _lib
        from voronoirt_b200.api import _ptr
        kk = np.ascontiguousarray(k)
        _lib.check(_lib.lib().vrt_regular_formal_solve(len(z), len(x), len(y), _ptr(atm.z), _ptr(atm.x), _ptr(atm.y), _ptr(kk), down, 3,
                                                       nlam, _ptr(dev[0]), _ptr(dev[2]), _ptr(dev[1]), _ptr(out), None))
        torch.cuda.synchronize()
        I3 = out.cpu().numpy().transpose(*range(S.ndim)[::-1])
        assert np.array_equal(I, I3)


def test_bad_arguments():
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib
    ax = np.linspace(0, 1, 6)
    with pytest.raises(ValueError):
        V.short_characteristics_up(V.direction(160, 45), np.zeros((6, 6, 5)), np.zeros((6, 6)), np.zeros((6, 6, 6)), V.Atmosphere(ax, ax, ax))
    with pytest.raises(_lib.VRTError):
        V.short_characteristics_up(np.array([0.0, 1.0, 0.0]), np.zeros((6, 6, 6)), np.zeros((6, 6)), np.zeros((6, 6, 6)), V.Atmosphere(ax, ax, ax))


def test_workspace_release_and_reuse(oracle):
    import voronoirt_b200 as V
    rng = np.random.default_rng(5)
    z, x, y, S, alpha, I0 = random_box(rng, 6, 9, 40, 2)
    k = kvec(112.8, 335.8)
    a, _ = solve(V, z, x, y, k, 0, S, I0, alpha)
    V.regular_release_workspace()
    b, _ = solve(V, z, x, y, k, 0, S, I0, alpha)          # re-allocates
    z2, x2, y2, S2, alpha2, I02 = random_box(rng, 8, 44, 11, 1)   # grows in one axis, shrinks in another
    c, _ = solve(V, z2, x2, y2, k, 0, S2[0], I02[0], alpha2[0])
    ref = oracle.short_characteristics(z2, x2, y2, k, 0, S2[0], I02[0], alpha2[0])[0]
    assert np.array_equal(a, b) and close(c, ref)
    V.regular_release_workspace()


def test_wide_rows_use_the_1024_thread_recurrence(oracle):
    """rows wider than 514 points run the second instantiation of the recurrence kernel (prefetch depth 4)"""
    import voronoirt_b200 as V
    rng = np.random.default_rng(13)
    z, x, y, S, alpha, I0 = random_box(rng, 5, 7, 603, 1)
    for theta, phi in ((70.3, 346.4), (109.7, 193.6)):
        down = int(theta < 90)
        k = kvec(theta, phi)
        ref, ref_branch = oracle.short_characteristics(z, x, y, k, down, S[0], I0[0], alpha[0])
        I, branch = solve(V, z, x, y, k, down, S[0], I0[0], alpha[0])
        assert np.array_equal(branch, ref_branch) and 2 in branch
        assert close(I, ref)


def continuum_box(rng, nz, nx, ny):
    """periodic alpha, eps, B0 for the regular continuum Λ-iteration (lambda_continuum.jl:58-107)"""
    z = np.cumsum(rng.uniform(0.3, 2.0, nz))
    x = (np.arange(nx) - 1) * 1.1
    y = (np.arange(ny) - 1) * 0.9
    alpha = np.asfortranarray(10.0 ** rng.uniform(-2, 1, (nz, nx, ny)))
    eps_l = np.asfortranarray(10.0 ** rng.uniform(-5, 0, (nz, nx, ny)))      # both sides of the `thick` threshold 1e-4
    B0 = np.asfortranarray(rng.uniform(1.0, 2.0, (nz, nx, ny)))
    for a in (alpha, eps_l, B0):
        a[:, 0, :] = a[:, -2, :]; a[:, -1, :] = a[:, 1, :]
        a[:, :, 0] = a[:, :, -2]; a[:, :, -1] = a[:, :, 1]
    return z, x, y, alpha, eps_l, B0


def test_J_regular_vs_oracle_and_batched(oracle, monkeypatch):
    import voronoirt_b200 as V
    rng = np.random.default_rng(21)
    nlam = 3
    z, x, y, S, alpha, I0 = random_box(rng, 9, 13, 16, nlam)
    atm = V.Atmosphere(z, x, y)
    quad = (QUAD[:, 0], QUAD[:, 1], QUAD[:, 2])
    ref = np.stack([oracle.J_regular(z, x, y, *quad, S[l], alpha[l], I0[l]) for l in range(nlam)])
    J = V.J_λ_regular(S, alpha, atm, quad, I_0=I0)
    assert close(J, ref), np.abs(J - ref).max()
    J1 = V.J_λ_regular(S[1], alpha[1], atm, quad, I_0=I0[1])                 # one wavelength, 3-D arrays
    assert np.array_equal(J1, J[1])
    monkeypatch.setenv("VRT_REG_LAM_CHUNK", "2")
    J2 = V.J_λ_regular(S, alpha, atm, quad, I_0=I0)
    monkeypatch.delenv("VRT_REG_LAM_CHUNK")
    assert np.array_equal(J, J2)
    # J is the weighted sum of the single-direction solutions of the formal solver
    acc = np.zeros_like(S[0])
    for w, theta, phi in QUAD:
        down = int(theta < 90)
        bnd = np.zeros_like(I0[0]) if down else I0[0]
        acc += w * solve(V, z, x, y, kvec(theta, phi), down, S[0], bnd, alpha[0])[0]
    assert np.abs(acc - J[0]).max() <= 1e-13 * np.abs(J[0]).max()
    V.regular_release_workspace()


def test_lambda_regular_vs_oracle(oracle):
    import voronoirt_b200 as V
    rng = np.random.default_rng(22)
    z, x, y, alpha, eps_l, B0 = continuum_box(rng, 8, 11, 12)
    quad = (QUAD[:, 0], QUAD[:, 1], QUAD[:, 2])
    Jr, Sr, conv, it = oracle.lambda_regular(z, x, y, *quad, alpha, eps_l, B0, eps=1e-4, maxiter=60)
    seen = []
    J, S, a = V.Λ_regular(1e-4, 60, V.Atmosphere(z, x, y), quad, alpha, eps_l, B0, callback=lambda rec: seen.append(rec["diff"]))
    res = V.Λ_regular.last
    assert res["iterations"] == it and res["converged"] and 0 < it < 60
    assert close(S, Sr) and close(J, Jr)
    assert np.allclose(seen, conv[:it], rtol=1e-9) and abs(res["diff"] - conv[it]) <= 1e-9 * conv[it]
    # maxiter reached: not converged, same state as the oracle after 3 iterations
    Jr3, Sr3, conv3, it3 = oracle.lambda_regular(z, x, y, *quad, alpha, eps_l, B0, eps=1e-12, maxiter=3)
    J3, S3, _ = V.Λ_regular(1e-12, 3, V.Atmosphere(z, x, y), quad, alpha, eps_l, B0)
    assert V.Λ_regular.last["iterations"] == it3 == 3 and not V.Λ_regular.last["converged"]
    assert close(S3, Sr3) and close(J3, Jr3)
    V.regular_release_workspace()
