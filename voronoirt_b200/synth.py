"""Synthetic "Bifrost-shaped" inputs for tests and benchmarks.

There is no network and no Bifrost cube in this environment, so the atmosphere is an analytic 1-D
stratification times a smooth periodic 3-D perturbation, sampled directly at the Voronoi sites.  Sites are
drawn with the reference's rejection scheme for `sample_from_invNH_invT` (src/sample_grids.jl:223-230,
src/functions.jl:79-121): a candidate uniform in the box is accepted when q = log10(N_H)^-2 * T^-2/5
exceeds U(q_min, q_max).  Neighbour lists come from the reference's own voro++ driver
(rt_preprocessing/output_sites, run as a subprocess exactly like src/functions.jl:13-23).

The collisional rates and the continuum extinction are stand-ins with Bifrost-like magnitudes for the
Transparency.jl recipes the reference evaluates once on the host; they are INPUTS of the device engine.
"""
import os
import tempfile

import numpy as np

from . import api, atom

BOX = dict(z_min=-0.5e6, z_max=10.0e6, x_min=0.0, x_max=6.0e6, y_min=0.0, y_max=6.0e6)


def atmosphere(z, x, y, seed=2022):
    """-> dict of per-point temperature [K], hydrogen_density, electron_density [m^-3], velocity_z/x/y [m/s]"""
    rng = np.random.default_rng(seed)
    Lx = BOX["x_max"] - BOX["x_min"]
    Ly = BOX["y_max"] - BOX["y_min"]
    Lz = BOX["z_max"] - BOX["z_min"]
    # 8 seeded periodic Fourier modes, +-20 %
    pert = np.zeros_like(z)
    vel = [np.zeros_like(z) for _ in range(3)]
    for _ in range(8):
        kx, ky = rng.integers(1, 4, size=2)
        kz = rng.integers(1, 6)
        ph = rng.uniform(0, 2 * np.pi, size=4)
        arg = 2 * np.pi * (kx * x / Lx + ky * y / Ly) + np.pi * kz * (z - BOX["z_min"]) / Lz
        pert += np.sin(arg + ph[0]) / 8
        for a in range(3):
            vel[a] += np.sin(arg + ph[a + 1])
    pert *= 0.2 * np.sqrt(8.0)
    zz = z / 1e6  # Mm
    # T(z): 1.5e4 K at the bottom -> 4.4e3 K minimum near 0.5 Mm -> ~8e3 K chromosphere -> 1e6 K corona above ~2.2 Mm
    T_phot = 4400.0 + 10600.0 * np.clip((0.5 - zz) / 1.0, 0, None) ** 1.5
    T_chrom = 4400.0 + 3600.0 * (1 - np.exp(-np.clip(zz - 0.5, 0, None) / 0.4))
    T_low = np.where(zz < 0.5, T_phot, T_chrom)
    s = 0.5 * (1 + np.tanh((zz - 2.2) / 0.08))
    T = np.exp((1 - s) * np.log(T_low) + s * np.log(1.0e6)) * (1 + pert)
    NH = np.maximum(1.2e23 * np.exp(-(z - BOX["z_min"]) / 150e3), 1e15) * (1 + pert)
    xion = np.clip(1e-4 + 1.0 / (1 + np.exp(-(T - 9000.0) / 1200.0)), 1e-4, 1.0)
    ne = NH * xion
    sc = 5e3 / np.sqrt(4.0)
    return dict(temperature=T, hydrogen_density=NH, electron_density=ne,
                velocity_z=vel[0] * sc, velocity_x=vel[1] * sc, velocity_y=vel[2] * sc)


def sample_sites(n, seed=2022, stratified=True):
    """-> positions (3, n) Fortran array, rows (z, x, y).  stratified=False gives uniform sites in the box."""
    rng = np.random.default_rng(seed)
    lo = np.array([BOX["z_min"], BOX["x_min"], BOX["y_min"]])
    hi = np.array([BOX["z_max"], BOX["x_max"], BOX["y_max"]])
    if not stratified:
        return np.asfortranarray((lo[:, None] + (hi - lo)[:, None] * rng.random((3, n))))
    # q on a coarse probe to get (q_min, q_max), as the reference does on the atmosphere grid
    probe = lo[:, None] + (hi - lo)[:, None] * rng.random((3, 200000))
    a = atmosphere(probe[0], probe[1], probe[2])
    q = np.log10(a["hydrogen_density"]) ** -2.0 * a["temperature"] ** (-2.0 / 5.0)
    qmin, qmax = q.min(), q.max()
    out = np.empty((3, n), order="F")
    got = 0
    while got < n:
        m = max(4 * (n - got), 100000)
        cand = lo[:, None] + (hi - lo)[:, None] * rng.random((3, m))
        a = atmosphere(cand[0], cand[1], cand[2])
        q = np.log10(a["hydrogen_density"]) ** -2.0 * a["temperature"] ** (-2.0 / 5.0)
        keep = q > rng.uniform(qmin, qmax, size=m)
        sel = cand[:, keep][:, : n - got]
        out[:, got:got + sel.shape[1]] = sel
        got += sel.shape[1]
    return out


FIELDS = ("temperature", "electron_density", "hydrogen_density", "velocity_z", "velocity_x", "velocity_y")


def cube_atmosphere(nz=216, nx=129, ny=129, seed=2022):
    """The synthetic atmosphere on a regular cube spanning BOX (what the reference reads from its Bifrost HDF5 file,
    src/io.jl get_atmos): -> (api.Atmosphere, quantity) with quantity = log10(N_H)^-2 T^-2/5, the density the reference
    samples its sites from (src/sample_grids.jl:223-230)."""
    z = np.linspace(BOX["z_min"], BOX["z_max"], nz)
    x = np.linspace(BOX["x_min"], BOX["x_max"], nx)
    y = np.linspace(BOX["y_min"], BOX["y_max"], ny)
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    a = atmosphere(Z.ravel(), X.ravel(), Y.ravel(), seed)
    f = {k: np.asfortranarray(a[k].reshape(nz, nx, ny)) for k in FIELDS}
    atm = api.Atmosphere(z, x, y, f["temperature"], f["electron_density"], f["hydrogen_density"], f["velocity_z"], f["velocity_x"],
                         f["velocity_y"])
    q = np.asfortranarray(np.log10(f["hydrogen_density"]) ** -2.0 * f["temperature"] ** (-2.0 / 5.0))
    return atm, q


def native_sites(n, seed=2022, cube=None):
    """The reference's set-up pipeline on the GPU (compare_line.jl:49-110 without voro++): rejection sampling of n sites
    from the cube (functions.jl:79-121 -> vrt_rejection_sampling) and trilinear initialisation of the six per-site fields
    (voronoi_utils.jl:687-708 -> vrt_trilinear).  -> positions (3, n), dict of per-site fields"""
    atm, q = cube or cube_atmosphere(seed=seed)
    pos = api.rejection_sampling(n, atm, q, seed)
    vals = api.initialise(pos, atm)
    return pos, dict(zip(FIELDS, vals))


def _trilinear_np(atm, vals, pos):
    """numpy restatement of functions.jl:207-248 on the uniform cube of cube_atmosphere (CPU arm of bench.py only)"""
    def cell(ax, p):
        i = np.clip(((p - ax[0]) / (ax[1] - ax[0])).astype(np.int64), 0, len(ax) - 2)
        return i, (p - ax[i]) / (ax[i + 1] - ax[i])
    iz, tz = cell(atm.z, pos[0])
    ix, tx = cell(atm.x, pos[1])
    iy, ty = cell(atm.y, pos[2])
    out = 0.0
    for dz, wz in ((0, 1 - tz), (1, tz)):
        for dx, wx in ((0, 1 - tx), (1, tx)):
            for dy, wy in ((0, 1 - ty), (1, ty)):
                out = out + wz * wx * wy * vals[iz + dz, ix + dx, iy + dy]
    return out


def native_sites_cpu(n, seed=2022, cube=None):
    """the same pipeline in numpy (same cube, same acceptance rule, numpy's random stream): the sites of the CPU baseline arm"""
    atm, q = cube or cube_atmosphere(seed=seed)
    rng = np.random.default_rng(seed)
    lo = np.array([atm.z[0], atm.x[0], atm.y[0]])
    hi = np.array([atm.z[-1], atm.x[-1], atm.y[-1]])
    qmin, qmax = q.min(), q.max()
    out = np.empty((3, n), order="F")
    got = 0
    while got < n:
        m = max(4 * (n - got), 100000)
        cand = lo[:, None] + (hi - lo)[:, None] * rng.random((3, m))
        keep = _trilinear_np(atm, q, cand) > rng.uniform(qmin, qmax, size=m)
        sel = cand[:, keep][:, : n - got]
        out[:, got:got + sel.shape[1]] = sel
        got += sel.shape[1]
    fields = (atm.temperature, atm.electron_density, atm.hydrogen_populations, atm.velocity_z, atm.velocity_x, atm.velocity_y)
    return out, {k: _trilinear_np(atm, f, out) for k, f in zip(FIELDS, fields)}


def voronoi_neighbours(positions, bounds=None, voro_exec=None, workdir=None, keep=False, parse=None):
    """write_arrays + voro + the parsing half of read_cell -> NeighbourMatrix (n, ld).
    parse(fname, n) -> (n, ld) matrix replaces api.read_neighbours (the CPU baseline arm of bench.py parses with the
    oracle so that its process never loads libvrt.so)."""
    b = bounds or BOX
    voro_exec = voro_exec or api.default_voro_exec()
    if voro_exec is None:
        raise FileNotFoundError("voro++ driver not found (VORO_EXEC, /root/reference/rt_preprocessing/output_sites, baseline/_ref/output_sites)")
    n = positions.shape[1]
    tmp = workdir or tempfile.mkdtemp(prefix="vrt_voro_")
    sites_file = os.path.join(tmp, "sites.txt")
    nb_file = os.path.join(tmp, "neighbours.txt")
    # src/compare_line.jl:91-94: x = positions[2,:], y = positions[3,:], z = positions[1,:]
    ids = np.arange(1, n + 1)
    np.savetxt(sites_file, np.column_stack([ids, positions[1], positions[2], positions[0]]), fmt=["%d", "%.17g", "%.17g", "%.17g"], delimiter="\t")
    api.voro(voro_exec, sites_file, nb_file, b["x_min"], b["x_max"], b["y_min"], b["y_max"], b["z_min"], b["z_max"])
    nbr = (parse or api.read_neighbours)(nb_file, n)
    if not keep and workdir is None:
        for f in (sites_file, nb_file):
            os.remove(f)
        os.rmdir(tmp)
    return nbr


def make_sites(n, seed=2022, stratified=True, voro_exec=None):
    """positions + voro++ + read_cell + atmosphere -> api.VoronoiSites (mirrors compare_line.jl:49-110)"""
    pos = sample_sites(n, seed, stratified)
    nbr = voronoi_neighbours(pos, voro_exec=voro_exec)
    return sites_from(pos, nbr, seed)


def sites_from(pos, nbr, seed=2022, bounds=None):
    b = bounds or BOX
    n = pos.shape[1]
    a = atmosphere(pos[0], pos[1], pos[2], seed)
    cell = api.read_cell(nbr, n, pos, b["x_min"], b["x_max"], b["y_min"], b["y_max"])
    return api.VoronoiSites(*cell, a["temperature"], a["electron_density"], a["hydrogen_density"],
                            a["velocity_z"], a["velocity_x"], a["velocity_y"],
                            b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"], n)


def line_inputs(T, ne, NH, nλ_bb=50, nλ_bf=20):
    """What Λ_voronoi computes before its loop (src/lambda_iteration.jl:216-247), with stand-in recipes for the
    Transparency.jl parts: -> (line, LTE_pops (n,3), α_cont (n), ελ (n), C (3,3,n))."""
    T, ne, NH = (np.asarray(a, dtype=np.float64) for a in (T, ne, NH))
    n = len(T)
    line = atom.HydrogenicLine(*atom.test_atom(nλ_bb, nλ_bf), T)
    lte = atom.LTE_populations(line, T, ne, NH)
    # stand-in collisional rates (Seaton / van Regemorter magnitude), x BOOST like src/rates.jl:550
    def up(dE, g, ups):
        return ne * 8.63e-12 * 1e-6 / (g * np.sqrt(T)) * ups * np.exp(-np.minimum(dE / (atom.k_B * T), 600.0))
    C = np.zeros((3, 3, n), order="F")
    C12 = up(line.χj - line.χi, line.gi, 0.6)
    C13 = up(line.χ_inf - line.χi, line.gi, 0.2)
    C23 = up(line.χ_inf - line.χj, line.gj, 2.0)
    tiny = 1e-300
    C[0, 1] = C12
    C[1, 0] = C12 * lte[:, 0] / np.maximum(lte[:, 1], tiny)
    C[0, 2] = C13
    C[2, 0] = C13 * lte[:, 0] / np.maximum(lte[:, 2], tiny)
    C[1, 2] = C23
    C[2, 1] = C23 * lte[:, 1] / np.maximum(lte[:, 2], tiny)
    C *= atom.BOOST
    C = np.asfortranarray(C)
    ελ = atom.destruction(lte, C[1, 0], T, line)
    # stand-in continuum extinction: Thomson + a neutral-hydrogen term; 1e-3 m^-1 at the bottom .. 1e-12 in the corona
    α_cont = 6.652e-29 * ne + 1.0e-26 * (lte[:, 0] + lte[:, 1])
    return line, lte, α_cont, ελ, C


def continuum_inputs(T, ne, NH):
    """stand-ins for src/lambda_continuum.jl:117-136 -> (α_cont, ε_λ, B_0)"""
    T, ne, NH = (np.asarray(a, dtype=np.float64) for a in (T, ne, NH))
    α_s = 6.652e-29 * ne + 1e-32 * NH
    α_a = 1.0e-26 * NH * np.clip(T / 6000.0, 0.2, 5.0) ** 2
    α_cont = α_s + α_a
    return α_cont, α_a / α_cont, atom.B_λ(500.0, T)


def tile_grid(pos, nbr, kx, ky, bounds=None):
    """Replicate a grid that is periodic in x and y into a kx x ky times larger periodic box.

    The Voronoi tessellation of the replicated point set is exactly the replicated tessellation, so the
    neighbour lists (and their order) carry over; only the tile of each neighbour has to be found, from the
    nearest periodic image.  Used to reach benchmark sizes (1 M - 16 M sites) from a base grid that the
    single-threaded voro++ driver can tessellate in seconds.  -> (positions (3, n), NeighbourMatrix (n, ld), bounds dict)
    """
    b = dict(bounds or BOX)
    n0, ld = nbr.shape
    Lx, Ly = b["x_max"] - b["x_min"], b["y_max"] - b["y_min"]
    x, y = pos[1], pos[2]
    ids = nbr[:, 1:]
    valid = ids > 0
    j = np.where(valid, ids - 1, 0)
    sx = np.where(valid, np.rint((x[:, None] - x[j]) / Lx), 0).astype(np.int64)
    sy = np.where(valid, np.rint((y[:, None] - y[j]) / Ly), 0).astype(np.int64)
    n = n0 * kx * ky
    out_pos = np.empty((3, n), order="F")
    out_nbr = np.zeros((n, ld), dtype=np.int64, order="F")
    for a in range(kx):
        for c in range(ky):
            t = a * ky + c
            sl = slice(t * n0, (t + 1) * n0)
            out_pos[0, sl] = pos[0]
            out_pos[1, sl] = x + a * Lx
            out_pos[2, sl] = y + c * Ly
            tt = ((a + sx) % kx) * ky + ((c + sy) % ky)
            out_nbr[sl, 0] = nbr[:, 0]
            out_nbr[sl, 1:] = np.where(valid, ids + tt * n0, ids)
    b["x_max"] = b["x_min"] + kx * Lx
    b["y_max"] = b["y_min"] + ky * Ly
    return out_pos, out_nbr, b
