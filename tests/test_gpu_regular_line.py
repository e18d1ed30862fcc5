"""NLTE line path on the regular grid through the generic Λ-iteration engine (vrt_regular_grid_create + vrt_solver_*):
J_λ_regular (lambda_iteration.jl:1-58), calculate_R / get_revised_populations on the regular arrays (rates.jl:96-143,
populations.jl:147-182) and Λ_regular (lambda_iteration.jl:116-205) against the oracle.  Tolerances as on the Voronoi
path: I, J, S 1e-9 relative; populations 1e-6 (+1e-13 N_H)."""
import numpy as np
import pytest

from regular_box import regular_line_box
from test_gpu_parity import pointwise_rel, pops_close, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    import voronoirt_b200 as V
    return V


def atmosphere_of(V, P):
    f = P["fields"]
    return V.Atmosphere(P["z"], P["x"], P["y"], f["temperature"], f["electron_density"], f["hydrogen_density"], f["velocity_z"],
                        f["velocity_x"], f["velocity_y"])


def test_J_lambda_regular_line_rates_populations(V, oracle):
    from voronoirt_b200 import atom
    P = regular_line_box(oracle)
    line, n, shape = P["line"], P["n"], P["shape"]
    atm = atmosphere_of(V, P)
    qp = V.quadrature_path("ul7n12")
    w, th, ph, nq = V.read_quadrature(qp)
    T = P["flat"]["temperature"]
    S = np.asfortranarray(atom.B_λ(line.λ[:, None], T[None, :]))                    # (nλ, n) == (nλ, nz, nx, ny) memory
    pops4 = P["lte"].reshape(shape + (3,), order="F")
    J, damping = V.J_λ_regular(S.reshape((len(line.λ),) + shape, order="F"), P["α_cont"], pops4, atm, line, qp)
    assert J.shape == (len(line.λ),) + shape
    oq = oracle.make_quadrature(w, th, ph)
    Jref, dref = oracle.J_lambda_regular(P["z"], P["x"], P["y"], line.as_struct(), line.λ, P["sd"], oq, S.T, P["lte"].T)
    Jf = J.reshape((len(line.λ), n), order="F")
    df = damping.reshape((len(line.λ), n), order="F")
    assert pointwise_rel(df.T, dref, 1e-300) < 1e-12
    for l in range(Jf.shape[0]):
        assert rel_err(Jf[l], Jref[:, l]) < 1e-9, l
    # rates and statistical equilibrium per cell on the same handle
    R = V.calculate_R(atm, line, Jf, df, P["lte"], qp)
    Rref = oracle.calculate_R(line.as_struct(), line.λ, T, line.ΔD, Jref, dref, P["lte"].T)
    Rt = np.ascontiguousarray(R.T)
    for (a_, b_) in ((0, 1), (1, 0), (0, 2), (2, 0), (1, 2), (2, 1)):
        assert pointwise_rel(Rt[:, b_, a_], Rref[:, b_, a_], 1e-300) < 1e-9, (a_, b_)
    pops = V.get_revised_populations(R, P["C"], P["flat"]["hydrogen_density"])
    pref = oracle.get_revised_populations(Rref, np.ascontiguousarray(P["C"].T), P["flat"]["hydrogen_density"])
    assert pops_close(pops.T, pref, P["flat"]["hydrogen_density"])


@pytest.mark.parametrize("lam_chunk", [0, 7])
def test_lambda_regular_line(V, oracle, lam_chunk):
    from voronoirt_b200 import atom
    P = regular_line_box(oracle)
    line, n, shape = P["line"], P["n"], P["shape"]
    atm = atmosphere_of(V, P)
    qp = V.quadrature_path("ul7n12")
    w, th, ph, nq = V.read_quadrature(qp)
    maxiter = 3
    kw = dict(lam_chunk=lam_chunk) if lam_chunk else {}
    J, S, α_cont, pops = V.Λ_regular(1e-3, maxiter, atm, line, qp, None, α_cont=P["α_cont"], ελ=P["ελ"], C=P["C"], LTE_pops=P["lte"], **kw)
    res = V.Λ_regular.last
    assert J.shape == (len(line.λ),) + shape and pops.shape == shape + (3,)
    S0 = atom.B_λ(line.λ[:, None], P["flat"]["temperature"][None, :]).T
    oq = oracle.make_quadrature(w, th, ph)
    Jr, Sr, pr, conv, it = oracle.lambda_regular_line(P["z"], P["x"], P["y"], line.as_struct(), line.λ, P["sd"], oq, S0, P["lte"].T,
                                                      eps=1e-3, maxiter=maxiter)
    assert res["iterations"] == it == maxiter
    nl = len(line.λ)
    assert rel_err(S.reshape((nl, n), order="F").T, Sr) < 1e-9 and rel_err(J.reshape((nl, n), order="F").T, Jr) < 1e-9
    assert pops_close(pops.reshape((n, 3), order="F").T, pr, P["flat"]["hydrogen_density"])
    diffs = [h["diff"] for h in res["history"]] + [res["diff"]]
    assert np.allclose(diffs, conv[:len(diffs)], rtol=1e-9) and diffs[0] == 1.0


def test_regular_handle_rejects_voronoi_only_queries(V):
    from voronoirt_b200 import _lib
    import ctypes as C
    ax = np.linspace(0, 1, 6)
    atm = V.Atmosphere(ax, ax, ax)
    L = C.c_int64()
    assert _lib.lib().vrt_grid_num_layers(atm._grid.h, 0, C.byref(L)) == -5          # VRT_E_STATE
    k = np.ascontiguousarray(V.direction(160, 45))
    z = np.zeros(216)
    rc = _lib.lib().vrt_formal_solve(atm._grid.h, C.c_void_p(k.ctypes.data), 0, 7.0, 3, 1, C.c_void_p(z.ctypes.data),
                                     C.c_void_p(z.ctypes.data), None, C.c_void_p(z.ctypes.data))
    assert rc == -5 and b"regular" in _lib.lib().vrt_last_error()
