"""Regular-grid short-characteristics oracle against the reference's own golden vectors.

Provenance of the fixtures: tests/golden/ref_I_160_45_regular.npy and ref_I_20_15_regular.npy are byte copies of the
reference's data/searchlight_data/I_160_45_regular.npy and I_20_15_regular.npy (49 x 49 float64 each, written by
src/compare_searchlight.jl:154-225 `searchlight_regular`).  They are DATA produced by the reference's Julia code, the
only numerical output of the reference solver that ships with it, so this is the one place where the oracle is pinned
by the reference itself rather than by a second restatement.  The recipe below restates compare_searchlight.jl:154-190.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def searchlight_regular_inputs(n=51, R0=0.1):
    """compare_searchlight.jl:159-190: unit box, LinRange(0,1,n) axes, S = alpha = 0, disk of radius R0 on the
    boundary plane with the reference's own (1-based index)/n centring."""
    ax = np.linspace(0.0, 1.0, n)
    S = np.zeros((n, n, n), order="F")
    alpha = np.zeros((n, n, n), order="F")
    I0 = np.zeros((n, n), order="F")
    for i in range(1, n + 1):
        for j in range(1, n + 1):
            if np.sqrt((i / n - 0.5) ** 2 + (j / n - 0.5) ** 2) < R0:
                I0[i - 1, j - 1] = 1.0
    return ax, S, alpha, I0


def kvec(theta, phi):
    t, p = theta * np.pi / 180, phi * np.pi / 180
    return np.array([np.cos(t), np.cos(p) * np.sin(t), np.sin(p) * np.sin(t)])


@pytest.mark.parametrize("theta,phi,down,fname", [
    (160.0, 45.0, 0, "ref_I_160_45_regular.npy"),
    # the file name carries the older opposite-azimuth convention: phi = 15 + 180
    (20.0, 195.0, 1, "ref_I_20_15_regular.npy"),
])
def test_oracle_reproduces_reference_searchlight(oracle, theta, phi, down, fname):
    ax, S, alpha, I0 = searchlight_regular_inputs()
    I, planes = oracle.short_characteristics(ax, ax, ax, kvec(theta, phi), down, S, I0, alpha)
    out = (I[0] if down else I[-1])[1:-1, 1:-1]
    gold = np.load(os.path.join(GOLD, fname))
    assert out.shape == gold.shape == (49, 49)
    assert np.abs(out - gold).max() <= 1e-15          # measured 2.2e-16 / 6.4e-16
    assert abs(out.sum() - gold.sum()) < 1e-12
    assert set(planes.tolist()) == {0, 1}             # both goldens exercise the xy branch only


def test_oracle_regular_flux_conservation_all_branches(oracle):
    """compare_searchlight.jl:209 prints sum(I_top); with S = alpha = 0 every branch conserves sum(I_0) = 80 on the
    interior (bilinear weights sum to one, periodic wrap).  ul7n12 covers the xy, yz and xz branches."""
    ax, S, alpha, I0 = searchlight_regular_inputs()
    total = I0[1:-1, 1:-1].sum()
    quad = np.loadtxt(os.path.join(HERE, "..", "voronoirt_b200", "quadratures", "ul7n12.dat"))
    seen = set()
    for _, theta, phi in quad:
        down = int(theta < 90)
        I, planes = oracle.short_characteristics(ax, ax, ax, kvec(theta, phi), down, S, I0, alpha)
        out = (I[0] if down else I[-1])[1:-1, 1:-1]
        seen |= set(planes.tolist())
        assert abs(out.sum() - total) < 1e-9
        assert out.min() >= -1e-15 and out.max() <= 1 + 1e-12
    assert seen == {0, 1, 2, 3}


def test_oracle_regular_thick_limit(oracle):
    """S = const, large alpha: I -> S on every plane but the boundary (linear_weights' dtau > 50 branch: a + b = 1)."""
    n = 12
    ax = np.linspace(0.0, 1.0, n)
    S = np.full((n, n, n), 3.5, order="F")
    alpha = np.full((n, n, n), 1e4, order="F")
    I0 = np.zeros((n, n), order="F")
    for theta, phi in ((152.7, 315.5), (109.7, 193.6), (101.8, 235.4), (27.3, 135.5), (70.3, 346.4), (78.2, 55.4)):
        down = int(theta < 90)
        I, _ = oracle.short_characteristics(ax, ax, ax, kvec(theta, phi), down, S, I0, alpha)
        inner = I[:-1] if down else I[1:]
        assert np.abs(inner - 3.5).max() < 1e-12


def test_oracle_J_regular_is_the_weighted_sum_and_lambda_converges(oracle):
    """lambda_continuum.jl:1-24, :58-107: J = Σ w_i I_i (up from I_0, down from zero); with ε = 1 the source function
    stays B_0 and the loop stops after one iteration with criterion 0."""
    rng = np.random.default_rng(4)
    n = 8
    ax = np.linspace(0.0, 1.0, n)
    quad = np.loadtxt(os.path.join(HERE, "..", "voronoirt_b200", "quadratures", "ul7n12.dat"))
    S = np.asfortranarray(rng.uniform(0.5, 1.5, (n, n, n)))
    al = np.asfortranarray(rng.uniform(0.5, 5.0, (n, n, n)))
    I0 = np.asfortranarray(rng.uniform(0.0, 1.0, (n, n)))
    J = oracle.J_regular(ax, ax, ax, quad[:, 0], quad[:, 1], quad[:, 2], S, al, I0)
    acc = np.zeros_like(S)
    for w, theta, phi in quad:
        down = int(theta < 90)
        acc += w * oracle.short_characteristics(ax, ax, ax, kvec(theta, phi), down, S, np.zeros_like(I0) if down else I0, al)[0]
    assert np.abs(acc - J).max() <= 1e-15
    ones = np.ones((n, n, n), order="F")
    Jc, Sc, conv, it = oracle.lambda_regular(ax, ax, ax, quad[:, 0], quad[:, 1], quad[:, 2], al, ones, S, eps=1e-3, maxiter=10)
    assert it == 1 and conv[0] == 1.0 and conv[1] == 0.0 and np.array_equal(Sc, S)


def test_oracle_regular_line_J_reduces_to_the_continuum_form(oracle):
    """with zero level populations the line opacity vanishes (line.jl:219-225), so lambda_iteration.jl:1-58 must give, per
    wavelength, what lambda_continuum.jl:1-24 gives for α_cont with the boundary B_λ(λ, T[1,:,:]): two separately written
    oracle routines, same numbers."""
    import sys
    sys.path.insert(0, HERE)
    from regular_box import regular_line_box
    from voronoirt_b200 import atom
    P = regular_line_box(oracle, nz=7, nx=6, ny=7, nbb=4, nbf=2)
    line, shape, n = P["line"], P["shape"], P["n"]
    quad = np.loadtxt(os.path.join(HERE, "..", "voronoirt_b200", "quadratures", "ul7n12.dat"))
    oq = oracle.make_quadrature(quad[:, 0], quad[:, 1], quad[:, 2])
    T = P["flat"]["temperature"]
    S = atom.B_λ(line.λ[:, None], T[None, :]).T * 0.7                                   # (n, nλ)
    J, damping = oracle.J_lambda_regular(P["z"], P["x"], P["y"], line.as_struct(), line.λ, P["sd"], oq, S, np.zeros((3, n)))
    a3 = P["α_cont"].reshape(shape, order="F")
    for l in (0, len(line.λ) // 2, len(line.λ) - 1):
        I0 = atom.B_λ(line.λ[l], P["fields"]["temperature"][0])
        Jc = oracle.J_regular(P["z"], P["x"], P["y"], quad[:, 0], quad[:, 1], quad[:, 2], S[:, l].reshape(shape, order="F"), a3, I0)
        assert np.abs(Jc.ravel(order="F") - J[:, l]).max() <= 1e-14 * np.abs(J[:, l]).max()
