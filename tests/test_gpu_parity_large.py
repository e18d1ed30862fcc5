"""Parity of the CUDA path against the CPU oracle at BASELINE sizes (the oracle needs seconds to a minute here):
  * 1 M sites sampled from the synthetic atmosphere and tessellated on the GPU, NLTE line opacities, 2 directions x 4
    wavelengths of J_λ_voronoi (lambda_iteration.jl:60-113);
  * BASELINE configs[0], the searchlight of compare_searchlight.jl:10-152: 51^3 uniform sites in the unit box, S = 0, α = 0,
    unit beam of radius 0.1, every direction of ul7n12, p = 7;
  * its Hayek variant (compare_searchlight.jl:227-356): 100^3 sites, θ = 151.9°, ϕ = 45°, p = 50, I_0 = 1 on x, y <= 0.3.
Tolerance 1e-9 relative to the largest value (BASELINE.json north_star); the measured errors go to gpurun_out/parity_large.json.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, oracle_sites

pytestmark = pytest.mark.gpu


def record(key, value):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    f = os.path.join(d, "parity_large.json")
    cur = json.load(open(f)) if os.path.exists(f) else {}
    cur[key] = value
    json.dump(cur, open(f, "w"), indent=1)


def test_J_line_1m_sites_against_oracle(oracle):
    import voronoirt_b200 as V
    from voronoirt_b200 import atom, synth
    n = 1_000_000
    pos, a = synth.native_sites(n, seed=7)
    B = synth.BOX
    nbr = V.voronoi_neighbours(pos, B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"])
    bounds = [B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"]]
    cell = V.read_cell(nbr, n, pos, B["x_min"], B["x_max"], B["y_min"], B["y_max"])
    sites = V.VoronoiSites(*cell, a["temperature"], a["electron_density"], a["hydrogen_density"], a["velocity_z"], a["velocity_x"],
                           a["velocity_y"], *bounds, n)
    line, lte, α_cont, ελ, Cr = synth.line_inputs(a["temperature"], a["electron_density"], a["hydrogen_density"], 50, 20)
    w, th, ph, nq = V.read_quadrature(V.quadrature_path("ul7n12"))
    pick = [2, 9]                      # one upward (θ > 90) and one downward ray of ul7n12
    assert th[pick[0]] > 90 > th[pick[1]]
    l0, nl = 24, 4                     # line core: the opacity spans many decades there
    osites = oracle_sites(oracle, pos, nbr, np.array(bounds))
    # bit-exact set-up at this size too
    for down in (0, 1):
        perm, off = osites.layers(down)
        assert np.array_equal(perm, sites.perm_down if down else sites.perm_up)
        assert np.array_equal(off, sites.layers_down if down else sites.layers_up)
    sd = oracle.make_site_data(temperature=a["temperature"], electron_density=a["electron_density"], hydrogen_density=a["hydrogen_density"],
                               velocity_z=a["velocity_z"], velocity_x=a["velocity_x"], velocity_y=a["velocity_y"], doppler_width=line.ΔD,
                               alpha_cont=α_cont, destruction=ελ, C=np.ascontiguousarray(Cr.T), lte_pops=np.ascontiguousarray(lte.T))
    S = np.ascontiguousarray(atom.B_λ(line.λ[None, :], a["temperature"][:, None]))
    oracle.set_num_threads(os.cpu_count() or 1)
    Jo, _ = oracle.J_lambda_voronoi(osites, line.as_struct(), line.λ, sd, oracle.make_quadrature(w[pick], th[pick], ph[pick]), S, lte.T,
                                    l0=l0, l1=l0 + nl)
    Jo = Jo[:, l0:l0 + nl]
    solver = V.Solver(sites, (w[pick], th[pick], ph[pick]), line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte, lam_range=(l0, l0 + nl))
    Jg = solver.mean_intensity(np.asfortranarray(S[:, l0:l0 + nl].T), lte).T
    solver.close()
    err = float(np.abs(Jg - Jo).max() / np.abs(Jo).max())
    pos_ = Jo > 0
    pw = float((np.abs(Jg - Jo)[pos_] / Jo[pos_]).max())
    record("J_line_1m", {"sites": n, "directions": 2, "wavelengths": nl, "max_rel_err": err, "max_pointwise_rel_err": pw})
    assert err < 1e-9, err


def uniform_grid(V, n, seed=2022):
    rng = np.random.default_rng(seed)
    pos = np.asfortranarray(rng.random((3, n)))
    nbr = V.voronoi_neighbours(pos, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    cell = V.read_cell(nbr, n, pos, 0.0, 1.0, 0.0, 1.0)
    z = np.zeros(n)
    return pos, nbr, V.VoronoiSites(*cell, z, z, z, z, z, z, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, n)


def test_searchlight_51_cubed_all_directions(oracle):
    import voronoirt_b200 as V
    n = 51 ** 3
    pos, nbr, sites = uniform_grid(V, n)
    osites = oracle_sites(oracle, pos, nbr, np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0]))
    w, th, ph, nq = V.read_quadrature(V.quadrature_path("ul7n12"))
    worst = 0.0
    for t, p in zip(th, ph):
        k = V.direction(t, p)
        down = int(t < 90)
        perm, off = (sites.perm_down, sites.layers_down) if down else (sites.perm_up, sites.layers_up)
        n1 = off[1] - 1
        c = perm[:n1] - 1
        I0 = (((pos[1, c] - 0.5) ** 2 + (pos[2, c] - 0.5) ** 2) < 0.1 ** 2).astype(np.float64)       # compare_searchlight.jl:76-99
        Ig = (V.Delaunay_downII if down else V.Delaunay_upII)(k, np.zeros(n), I0, np.zeros(n), sites, 3)
        Io = osites.formal_solve(k, down, np.zeros(n), np.zeros(n), I0)[:, 0]
        assert 0.0 <= Ig.min() and Ig.max() <= 1.0 + 1e-12       # convex combinations of the boundary values
        worst = max(worst, float(np.abs(Ig - Io).max()))
    record("searchlight_51", {"sites": n, "directions": int(nq), "max_abs_err_over_max_I0": worst})
    assert worst < 1e-9, worst


def test_searchlight_hayek_p50(oracle):
    import voronoirt_b200 as V
    n = 100 ** 3
    pos, nbr, sites = uniform_grid(V, n)
    osites = oracle_sites(oracle, pos, nbr, np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0]))
    k = V.direction(151.9, 45.0)                                                                  # compare_searchlight.jl:240-246
    n1 = sites.layers_up[1] - 1
    c = sites.perm_up[:n1] - 1
    I0 = ((pos[1, c] <= 0.3) & (pos[2, c] <= 0.3)).astype(np.float64)
    Ig = V.Delaunay_upII(k, np.zeros(n), I0, np.zeros(n), sites, 3, 50.0)
    Io = osites.formal_solve(k, 0, np.zeros(n), np.zeros(n), I0, p=50.0)[:, 0]
    err = float(np.abs(Ig - Io).max())
    record("searchlight_hayek_p50", {"sites": n, "max_abs_err_over_max_I0": err})
    assert err < 1e-9, err
