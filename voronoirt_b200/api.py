"""Host-side mirror of the reference's Julia interface for the irregular-grid path.

Same names, argument order and meaning as the functions the entry scripts call
(compare_searchlight.jl, compare_continuum.jl, compare_line.jl), implemented as thin calls into the
C ABI of libvrt.so (include/vrt.h).  Arrays use the reference's logical shapes in Fortran order, i.e.
exactly the memory layout of the Julia arrays: positions (3, n) rows (z, x, y); NeighbourMatrix
(n, ld); S_λ, J_λ, α, damping (nλ, n); populations (n, 3); R, C (3, 3, n).  Ids are 1-based.
All compute happens on the GPU; there is no CPU path in this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import _abi
from ._lib import check, lib, last_stats  # noqa: F401
from .atom import HydrogenicLine  # noqa: F401


def _f(a, dtype=np.float64):
    return np.asfortranarray(a, dtype=dtype)


def _ptr(a):
    """address of a numpy array, a torch CUDA tensor, or a raw int device pointer (None -> NULL)"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


# ------------------------------------------------------------------ functions.jl
def read_quadrature(fname):
    """src/functions.jl:33-63 -> (weights, θ_array, ϕ_array, n_points).  The point count comes from the table,
    not from the digits after the first 'n' of the path (Q14)."""
    tab = np.atleast_2d(np.loadtxt(fname, dtype=np.float64))
    return tab[:, 0].copy(), tab[:, 1].copy(), tab[:, 2].copy(), tab.shape[0]


QUADRATURE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quadratures")


def quadrature_path(name):
    """packaged copy of the reference's quadratures/<name>.dat (n1, n2, ul2n3, ul7n12, ul9n20)"""
    return os.path.join(QUADRATURE_DIR, name + ".dat")


def default_voro_exec():
    """the reference's voro++ driver: $VORO_EXEC, the reference tree, or the copy staged by baseline/Makefile"""
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.environ.get("VORO_EXEC"), "/root/reference/rt_preprocessing/output_sites",
              os.path.join(here, "..", "baseline", "_ref", "output_sites")):
        if p and os.path.exists(p) and os.access(p, os.X_OK):
            return p
    return None


def write_arrays(x, y, z, fname):
    """src/io.jl:8-40: one line per site, `id\\tx\\ty\\tz`, 1-based id."""
    x, y, z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
    ids = np.arange(1, len(x) + 1)
    with open(fname, "w") as f:
        for i, a, b, c in zip(ids, x, y, z):
            f.write(f"{i}\t{a!r}\t{b!r}\t{c!r}\n")


def voro(voro_executable, sites_file, neighbours_file, x_min, x_max, y_min, y_max, z_min, z_max):
    """src/functions.jl:13-23: run the voro++ driver (rt_preprocessing/output_sites.cc) as a subprocess."""
    subprocess.run([voro_executable, sites_file, neighbours_file, repr(float(x_min)), repr(float(x_max)),
                    repr(float(y_min)), repr(float(y_max)), repr(float(z_min)), repr(float(z_max))],
                   check=True, stdout=subprocess.DEVNULL)


# ------------------------------------------------------------------ voronoi_utils.jl
def voronoi_neighbours(positions, z_min, z_max, x_min, x_max, y_min, y_max):
    """Native replacement of write_arrays + voro + the parsing half of read_cell (src/io.jl:8-40,
    rt_preprocessing/output_sites.cc, src/voronoi_utils.jl:42-70): positions (3, n) rows (z, x, y) -> NeighbourMatrix
    (n, ld) int64 with the same neighbour sets as voro++ (row order differs, see include/vrt.h)."""
    pos = _f(positions)
    n = pos.shape[1]
    b = np.array([z_min, z_max, x_min, x_max, y_min, y_max], dtype=np.float64)
    nbr = np.zeros((n, 64), dtype=np.int64, order="F")
    need = C.c_int64()
    check(lib().vrt_voronoi_neighbours(n, _ptr(pos), _ptr(b), _ptr(nbr), 64, C.byref(need)))
    return np.asfortranarray(nbr[:, :need.value])


def nearest_site(positions, bounds, points):
    """nn(KDTree(positions), p) for every column of points (3, m) (src/voronoi_utils.jl:441-444) -> (idx 1-based (m,), dist (m,))"""
    pos, q = _f(positions), _f(points)
    b = np.ascontiguousarray(bounds, dtype=np.float64)
    m = q.shape[1]
    idx = np.zeros(m, dtype=np.int64)
    dist = np.zeros(m)
    check(lib().vrt_nearest_site(pos.shape[1], _ptr(pos), _ptr(b), m, _ptr(q), _ptr(idx), _ptr(dist)))
    return idx, dist


def Voronoi_to_Raster(sites, z, x, y, *fields):
    """The resampling core of src/voronoi_utils.jl:407-471: nearest site of every raster point (z[k], x[i], y[j]), then the
    gathers.  fields: per-site arrays (n,) or (nλ, n) -> rasters (nz, nx, ny) or (nλ, nz, nx, ny); returns (idx, *rasters)."""
    z, x, y = (np.asarray(a, dtype=np.float64) for a in (z, x, y))
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    pts = np.asfortranarray(np.stack([Z.ravel(order="F"), X.ravel(order="F"), Y.ravel(order="F")]))
    b = [sites.z_min, sites.z_max, sites.x_min, sites.x_max, sites.y_min, sites.y_max]
    idx, _ = nearest_site(sites.positions, b, pts)
    shape = (len(z), len(x), len(y))
    out = []
    for f in fields:
        f = np.asarray(f)
        g = f[..., idx - 1]
        out.append(g.reshape(f.shape[:-1] + shape, order="F"))
    return (idx.reshape(shape, order="F"),) + tuple(out)


def nearest_sites(positions, bounds, points, k):
    """knn(KDTree(positions), p, k) for every column of points (3, m) -> (idx (k, m) 1-based, dist (k, m)), ascending"""
    pos, q = _f(positions), _f(points)
    b = np.ascontiguousarray(bounds, dtype=np.float64)
    m = q.shape[1]
    idx = np.zeros((k, m), dtype=np.int64, order="F")
    dist = np.zeros((k, m), order="F")
    check(lib().vrt_nearest_sites(pos.shape[1], _ptr(pos), _ptr(b), m, _ptr(q), int(k), _ptr(idx), _ptr(dist)))
    return idx, dist


def Voronoi_to_Raster_inv_dist(sites, z, x, y, values, p=1.0, n_k=2):
    """The resampling core of src/voronoi_utils.jl:773-816: inverse-distance interpolation (inv_dist_itp, :848-860) over the
    n_k nearest sites of every raster point.  values (n,) or (n, c) -> raster (nz, nx, ny) or (nz, nx, ny, c)."""
    z, x, y = (np.asarray(a, dtype=np.float64) for a in (z, x, y))
    Z, X, Y = np.meshgrid(z, x, y, indexing="ij")
    pts = np.asfortranarray(np.stack([Z.ravel(order="F"), X.ravel(order="F"), Y.ravel(order="F")]))
    b = [sites.z_min, sites.z_max, sites.x_min, sites.x_max, sites.y_min, sites.y_max]
    idx, dist = nearest_sites(sites.positions, b, pts, n_k)
    vals = np.asarray(values, dtype=np.float64)
    v2 = vals.reshape(vals.shape[0], -1)
    avg = np.zeros(pts.shape[1])
    f = np.zeros((pts.shape[1], v2.shape[1]))
    for t in range(n_k):                                   # the loop of inv_dist_itp, in its order
        inv = 1.0 / dist[t] ** p
        avg = avg + inv
        f = f + v2[idx[t] - 1] * inv[:, None]
    f = f / avg[:, None]
    shape = (len(z), len(x), len(y))
    return f.reshape(shape + vals.shape[1:], order="F") if vals.ndim > 1 else f[:, 0].reshape(shape, order="F")


def initialiseII(p_vec, atmos):
    """src/voronoi_utils.jl:716-770: nearest-corner values of the six atmosphere fields at the sites"""
    pos = _f(p_vec)
    out = []
    for fld in (atmos.temperature, atmos.electron_density, atmos.hydrogen_populations, atmos.velocity_z, atmos.velocity_x, atmos.velocity_y):
        v = _f(fld)
        o = np.zeros(pos.shape[1])
        check(lib().vrt_nearest_corner(atmos.shape[0], atmos.shape[1], atmos.shape[2], _ptr(atmos.z), _ptr(atmos.x), _ptr(atmos.y), _ptr(v),
                                       pos.shape[1], _ptr(pos), _ptr(o)))
        out.append(o)
    return tuple(out)


def trilinear(positions, atmos, vals):
    """src/functions.jl:207-248 broadcast over the sites: positions (3, n) rows (z, x, y), vals (nz, nx, ny) -> (n,)"""
    pos = _f(positions)
    v = _f(vals)
    if v.shape != atmos.shape:
        raise ValueError("vals must be (nz, nx, ny)")
    out = np.zeros(pos.shape[1])
    check(lib().vrt_trilinear(atmos.shape[0], atmos.shape[1], atmos.shape[2], _ptr(atmos.z), _ptr(atmos.x), _ptr(atmos.y), _ptr(v),
                              pos.shape[1], _ptr(pos), _ptr(out)))
    return out


def rejection_sampling(n_sites, atmos, quantity, seed=2022):
    """src/functions.jl:79-121 -> p_vec (3, n_sites) rows (z, x, y); quantity (nz, nx, ny) on the atmosphere axes"""
    q = _f(quantity)
    if q.shape != atmos.shape:
        raise ValueError("quantity must be (nz, nx, ny)")
    pos = np.zeros((3, int(n_sites)), order="F")
    mean = C.c_double()
    check(lib().vrt_rejection_sampling(int(n_sites), atmos.shape[0], atmos.shape[1], atmos.shape[2], _ptr(atmos.z), _ptr(atmos.x),
                                       _ptr(atmos.y), _ptr(q), int(seed), _ptr(pos), C.byref(mean)))
    rejection_sampling.mean_trials = mean.value
    return pos


def initialise(p_vec, atmos):
    """src/voronoi_utils.jl:687-708 -> (temperature, N_e, N_H, velocity_z, velocity_x, velocity_y) at the sites"""
    return tuple(trilinear(p_vec, atmos, f) for f in (atmos.temperature, atmos.electron_density, atmos.hydrogen_populations,
                                                      atmos.velocity_z, atmos.velocity_x, atmos.velocity_y))


class _Grid:
    """owner of a vrt_grid handle"""

    def __init__(self, positions, nbr, bounds):
        self.positions = _f(positions)
        self.nbr = _f(nbr, np.int64)
        self.n = self.positions.shape[1]
        self.ld = self.nbr.shape[1]
        b = np.ascontiguousarray(bounds, dtype=np.float64)
        h = C.c_void_p()
        check(lib().vrt_grid_create(self.n, _ptr(self.positions), _ptr(self.nbr), self.ld, _ptr(b), C.byref(h)))
        self.h = h
        self.solvers = {}

    def __del__(self):
        try:
            for s in self.solvers.values():
                s.close()
            if self.h:
                lib().vrt_grid_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def layers(self, down):
        L = C.c_int64()
        check(lib().vrt_grid_num_layers(self.h, down, C.byref(L)))
        perm = np.zeros(self.n, dtype=np.int64)
        off = np.zeros(L.value + 1, dtype=np.int64)
        check(lib().vrt_grid_get_layers(self.h, down, _ptr(perm), _ptr(off)))
        return perm, off

    def delaunay_lines(self):
        mx = C.c_int64()
        check(lib().vrt_grid_size(self.h, None, C.byref(mx)))
        out = np.zeros((3, mx.value, self.n), order="F")
        check(lib().vrt_grid_get_delaunay_lines(self.h, _ptr(out)))
        return out

    def stencil(self, k, p=7.0):
        k = np.ascontiguousarray(k, dtype=np.float64)
        up = np.zeros((2, self.n), dtype=np.int64, order="F")
        dots, w, r = (np.zeros((2, self.n), order="F") for _ in range(3))
        check(lib().vrt_grid_get_stencil(self.h, _ptr(k), p, _ptr(up), _ptr(dots), _ptr(w), _ptr(r)))
        return up, dots, w, r

    def schedule(self, k, down, n_sweeps=3, prune=1):
        k = np.ascontiguousarray(k, dtype=np.float64)
        cls = np.zeros((2, self.n), dtype=np.int32, order="F")
        sub = np.zeros(self.n, dtype=np.int32)
        stab = np.zeros(self.n, dtype=np.int32)
        ns, nv = C.c_int64(), C.c_int64()
        check(lib().vrt_grid_get_schedule(self.h, _ptr(k), int(down), n_sweeps, prune, _ptr(cls), _ptr(sub), _ptr(stab),
                                          C.byref(ns), C.byref(nv)))
        return cls, sub, stab, ns.value, nv.value


class _NeighbourMatrix(np.ndarray):
    """NeighbourMatrix that remembers the device grid built from it by read_cell"""
    _grid = None


def read_neighbours(fname, n_sites):
    """the parsing half of read_cell (src/voronoi_utils.jl:42-70) -> NeighbourMatrix (n, max_nb+1)"""
    ld = C.c_int64()
    check(lib().vrt_read_neighbours(fname.encode(), n_sites, None, 0, C.byref(ld)))
    nbr = np.zeros((n_sites, ld.value), dtype=np.int64, order="F")
    check(lib().vrt_read_neighbours(fname.encode(), n_sites, _ptr(nbr), ld.value, C.byref(ld)))
    return nbr


def read_cell(fname, n_sites, positions, x_min, x_max, y_min, y_max):
    """src/voronoi_utils.jl:36-85 -> (positions, NeighbourMatrix, Delaunay_lines, layers_up, layers_down, perm_up, perm_down).

    `fname` may also be an already parsed NeighbourMatrix.  Delaunay_lines is returned lazily as None-free
    array only when small (< 2e6 entries); larger grids get an empty placeholder (the device keeps its own)."""
    nbr = read_neighbours(fname, n_sites) if isinstance(fname, (str, bytes)) else _f(fname, np.int64)
    positions = _f(positions)
    g = _Grid(positions, nbr, [0.0, 0.0, x_min, x_max, y_min, y_max])
    perm_up, layers_up = g.layers(0)
    perm_down, layers_down = g.layers(1)
    nm = nbr.view(_NeighbourMatrix)
    nm._grid = g
    lines = g.delaunay_lines() if 3 * n_sites * nbr.shape[1] < 2_000_000 else np.zeros((3, 0, n_sites), order="F")
    return positions, nm, lines, layers_up, layers_down, perm_up, perm_down


class VoronoiSites:
    """src/voronoi_utils.jl:7-28 (same field names and order)."""

    def __init__(self, positions, neighbours, Delaunay_lines, layers_up, layers_down, perm_up, perm_down,
                 temperature, electron_density, hydrogen_populations, velocity_z, velocity_x, velocity_y,
                 z_min, z_max, x_min, x_max, y_min, y_max, n):
        self.positions = positions
        self.neighbours = neighbours
        self.Delaunay_lines = Delaunay_lines
        self.layers_up, self.layers_down = layers_up, layers_down
        self.perm_up, self.perm_down = perm_up, perm_down
        self.temperature = np.asarray(temperature, dtype=np.float64)
        self.electron_density = np.asarray(electron_density, dtype=np.float64)
        self.hydrogen_populations = np.asarray(hydrogen_populations, dtype=np.float64)
        self.velocity_z = np.asarray(velocity_z, dtype=np.float64)
        self.velocity_x = np.asarray(velocity_x, dtype=np.float64)
        self.velocity_y = np.asarray(velocity_y, dtype=np.float64)
        self.z_min, self.z_max, self.x_min, self.x_max, self.y_min, self.y_max = z_min, z_max, x_min, x_max, y_min, y_max
        self.n = n
        g = getattr(neighbours, "_grid", None)
        if g is None:
            g = _Grid(positions, neighbours, [z_min, z_max, x_min, x_max, y_min, y_max])
        self._grid = g


# ------------------------------------------------------------------ io.jl: output / checkpoint file
class OutputFile:
    """vrt_outfile handle: the HDF5 file of create_output_file (src/io.jl:159-225), written without the HDF5 library"""

    def __init__(self, h, path):
        self.h, self.path = h, path

    def write(self, name, array):
        """write_to_file for any dataset: `array` is a numpy array (Julia shape, column-major) or a torch CUDA tensor"""
        a = array if hasattr(array, "data_ptr") else (np.asfortranarray(array, dtype=np.int64) if name in ("n_bb", "n_bf") else _f(array))
        nbytes = a.numel() * a.element_size() if hasattr(a, "data_ptr") else a.nbytes
        check(lib().vrt_output_write(self.h, name.encode(), _ptr(a), int(nbytes)))

    def write_convergence(self, iteration, difference):
        check(lib().vrt_output_write_convergence(self.h, int(iteration), float(difference)))

    def write_state(self, solver):
        """source_function and populations straight from the solver's device state (lambda_iteration.jl:280-281)"""
        check(lib().vrt_output_write_state(self.h, solver.h))

    def close(self):
        if self.h:
            h, self.h = self.h, None
            check(lib().vrt_output_close(h))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def create_output_file(output_path, nλ, size, maxiter):
    """src/io.jl:159-225: size = n_sites (Voronoi) or (nz, nx, ny) (regular grid) -> OutputFile"""
    h = C.c_void_p()
    if isinstance(size, (tuple, list)):
        nz, nx, ny = size
        check(lib().vrt_output_create_regular(str(output_path).encode(), int(nλ), int(nz), int(nx), int(ny), int(maxiter), C.byref(h)))
    else:
        check(lib().vrt_output_create(str(output_path).encode(), int(nλ), int(size), int(maxiter), C.byref(h)))
    return OutputFile(h, str(output_path))


def write_to_file(what, out, *args):
    """the write_to_file methods of src/io.jl:57-157 on an OutputFile: VoronoiSites / HydrogenicLine / (difference, iteration) /
    (n, field) / a source-function or population array"""
    if isinstance(what, VoronoiSites):
        out.write("positions", what.positions)
        for f in ("temperature", "electron_density", "hydrogen_populations", "velocity_z", "velocity_x", "velocity_y"):
            out.write(f, getattr(what, f))
        out.write("boundaries", np.array([what.z_min, what.z_max, what.x_min, what.x_max, what.y_min, what.y_max], dtype=np.float64))
    elif isinstance(what, HydrogenicLine):
        out.write("wavelength", what.λ)
        out.write("line_center", np.array([what.λ0]))
    elif isinstance(what, float) and args:
        out.write_convergence(args[0], what)
    elif isinstance(what, (int, np.integer)) and args:
        out.write(args[0], np.array([what], dtype=np.int64))
    else:
        a = np.asarray(what)
        out.write("populations" if a.shape[-1] == 3 and a.ndim in (2, 4) else "source_function", a)


def nccl_unique_id():
    """128 bytes to be handed to every member of a process group (vrt_nccl_unique_id)"""
    raw = C.create_string_buffer(128)
    check(lib().vrt_nccl_unique_id(raw))
    return raw.raw


def direction(θ, ϕ):
    """k = [cos θ, cos ϕ sin θ, sin ϕ sin θ] on (z, x, y), degrees (src/lambda_iteration.jl:87)"""
    t, p = θ * np.pi / 180, ϕ * np.pi / 180
    return np.array([np.cos(t), np.cos(p) * np.sin(t), np.sin(p) * np.sin(t)])


# ------------------------------------------------------------------ irregular_ray_tracing.jl
def _formal(k, S, I_0, α, sites, n_sweeps, p, down):
    S = _f(S)
    α = _f(α)
    nlam = 1 if S.ndim == 1 else S.shape[0]
    I_0 = _f(I_0)
    out = np.zeros_like(S, order="F")
    k = np.ascontiguousarray(k, dtype=np.float64)
    check(lib().vrt_formal_solve(sites._grid.h, _ptr(k), down, float(p), int(n_sweeps), nlam, _ptr(S), _ptr(α), _ptr(I_0), _ptr(out)))
    return out


def Delaunay_upII(k, S, I_0, α, sites, n_sweeps, p=7.0):
    """src/irregular_ray_tracing.jl:15-82.  S, α: (n,) or (nλ, n); I_0: (n₁,) or (nλ, n₁) on perm_up[1:n₁]."""
    return _formal(k, S, I_0, α, sites, n_sweeps, p, 0)


def Delaunay_downII(k, S, I_0, α, sites, n_sweeps, p=7.0):
    """src/irregular_ray_tracing.jl:96-163."""
    return _formal(k, S, I_0, α, sites, n_sweeps, p, 1)


# ------------------------------------------------------------------ characteristics.jl (regular grid)
class _RegularGrid:
    """owner of a vrt_grid handle over a regular atmosphere (vrt_regular_grid_create)"""

    def __init__(self, z, x, y):
        h = C.c_void_p()
        check(lib().vrt_regular_grid_create(len(z), len(x), len(y), _ptr(z), _ptr(x), _ptr(y), C.byref(h)))
        self.h = h
        self.n = len(z) * len(x) * len(y)
        self.solvers = {}

    def __del__(self):
        try:
            for s in self.solvers.values():
                s.close()
            if self.h:
                lib().vrt_grid_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Atmosphere:
    """src/atmosphere.jl:22-31 (same field names and order); x and y carry the periodic ghost columns.  The 3-D fields are
    (nz, nx, ny) arrays; their column-major memory is the per-cell vector the C ABI takes."""

    def __init__(self, z, x, y, temperature=None, electron_density=None, hydrogen_populations=None,
                 velocity_z=None, velocity_x=None, velocity_y=None):
        self.z, self.x, self.y = (np.ascontiguousarray(a, dtype=np.float64) for a in (z, x, y))
        self.temperature, self.electron_density, self.hydrogen_populations = temperature, electron_density, hydrogen_populations
        self.velocity_z, self.velocity_x, self.velocity_y = velocity_z, velocity_x, velocity_y
        self.shape = (len(self.z), len(self.x), len(self.y))
        self.n = self.shape[0] * self.shape[1] * self.shape[2]
        self._g = None

    @property
    def _grid(self):
        if self._g is None:
            self._g = _RegularGrid(self.z, self.x, self.y)
        return self._g


def _short_characteristics(k, S_0, I_0, α, atmos, n_sweeps, down, return_branches=False):
    nz, nx, ny = len(atmos.z), len(atmos.x), len(atmos.y)
    S_0 = _f(S_0)
    α = _f(α)
    I_0 = _f(I_0)
    if S_0.ndim == 3:
        nlam = 1
        if S_0.shape != (nz, nx, ny) or α.shape != (nz, nx, ny) or I_0.shape != (nx, ny):
            raise ValueError("S_0, α must be (nz, nx, ny) and I_0 (nx, ny)")
    else:
        nlam = S_0.shape[0]
        if S_0.shape != (nlam, nz, nx, ny) or α.shape != S_0.shape or I_0.shape != (nlam, nx, ny):
            raise ValueError("S_0, α must be (nλ, nz, nx, ny) and I_0 (nλ, nx, ny)")
    out = np.zeros_like(S_0, order="F")
    branch = np.zeros(nz, dtype=np.int32)
    k = np.ascontiguousarray(k, dtype=np.float64)
    check(lib().vrt_regular_formal_solve(nz, nx, ny, _ptr(atmos.z), _ptr(atmos.x), _ptr(atmos.y), _ptr(k), down, int(n_sweeps), nlam,
                                         _ptr(S_0), _ptr(α), _ptr(I_0), _ptr(out), _ptr(branch)))
    return (out, branch) if return_branches else out


def regular_release_workspace():
    """free the device workspace vrt_regular_formal_solve keeps between calls"""
    check(lib().vrt_regular_release_workspace())


def short_characteristics_up(k, S_0, I_0, α, atmos, n_sweeps=3, return_branches=False):
    """src/characteristics.jl:19-95.  S_0, α: (nz, nx, ny) or (nλ, nz, nx, ny); I_0: (nx, ny) or (nλ, nx, ny)."""
    return _short_characteristics(k, S_0, I_0, α, atmos, n_sweeps, 0, return_branches)


def short_characteristics_down(k, S_0, I_0, α, atmos, n_sweeps=3, return_branches=False):
    """src/characteristics.jl:110-180."""
    return _short_characteristics(k, S_0, I_0, α, atmos, n_sweeps, 1, return_branches)


# ------------------------------------------------------------------ Λ-iteration engine
def _as_quadrature(quadrature):
    if isinstance(quadrature, (str, bytes)):
        w, t, p, _ = read_quadrature(quadrature)
    else:
        w, t, p = (np.ascontiguousarray(a, dtype=np.float64) for a in quadrature[:3])
    q = _abi.vrt_quadrature(len(w), w.ctypes.data, t.ctypes.data, p.ctypes.data)
    q._keep = (w, t, p)
    return q


class Solver:
    """vrt_solver handle (state of Λ_voronoi).  kind: 'line' or 'continuum'."""

    def __init__(self, sites, quadrature, line=None, α_cont=None, ελ=None, C_rates=None, LTE_pops=None, B_0=None,
                 n_sweeps=3, p=7.0, lam_range=None, lam_chunk=0, prune=1, dir_range=None, cell_shard=None):
        self.sites = sites
        self.line = line
        q = _as_quadrature(quadrature)
        cfg = _abi.vrt_config()
        cfg.n_sweeps, cfg.p, cfg.prune, cfg.lam_chunk = n_sweeps, p, prune, lam_chunk
        if lam_range is not None:
            cfg.lam_begin, cfg.lam_end = lam_range
        if dir_range is not None:
            cfg.dir_begin, cfg.dir_end = dir_range
        if cell_shard is not None:
            cfg.cell_shard_rank, cfg.cell_shard_count = cell_shard
        h = C.c_void_p()
        n = sites.n
        self._keep = []
        if line is not None:
            sd = _abi.vrt_site_data()

            def put(name, a, shape=None):
                if a is None:
                    return
                a = _f(a)
                self._keep.append(a)
                setattr(sd, name, a.ctypes.data)

            put("temperature", sites.temperature)
            put("electron_density", sites.electron_density)
            put("hydrogen_density", sites.hydrogen_populations)
            put("velocity_z", sites.velocity_z)
            put("velocity_x", sites.velocity_x)
            put("velocity_y", sites.velocity_y)
            put("doppler_width", line.ΔD)
            put("alpha_cont", α_cont if α_cont is not None else np.zeros(n))
            put("destruction", ελ)
            put("C", C_rates)
            put("lte_pops", LTE_pops)
            ls = line.as_struct()
            lam = np.ascontiguousarray(line.λ, dtype=np.float64)
            check(lib().vrt_solver_create_line(sites._grid.h, C.byref(ls), _ptr(lam), C.byref(sd), C.byref(q), C.byref(cfg), C.byref(h)))
            self.kind = "line"
        else:
            a, e, b = _f(α_cont), _f(ελ), _f(B_0)
            check(lib().vrt_solver_create_continuum(sites._grid.h, _ptr(a), _ptr(e), _ptr(b), C.byref(q), C.byref(cfg), C.byref(h)))
            self.kind = "continuum"
        self.h = h
        nl = C.c_int64()
        check(lib().vrt_solver_nlam_local(self.h, C.byref(nl)))
        self.nlam = nl.value
        self.n = n
        self._cb = None
        self._ar = None

    def close(self):
        if getattr(self, "h", None):
            lib().vrt_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    FIELDS = {"alpha_cont": 0, "destruction": 1, "C": 2, "lte_pops": 3}

    def set_field(self, name, data):
        a = data if hasattr(data, "data_ptr") else _f(data)
        check(lib().vrt_solver_set_field(self.h, self.FIELDS[name], _ptr(a)))

    def set_allreduce(self, fn):
        """fn(dev_ptr:int, count:int, op:int) -> 0 on success; op 0 = sum over wavelength shards, 1 = max over all,
        2 = sum over direction shards, 3 = reduce-scatter / 4 = all-gather over direction shards (include/vrt.h)"""
        def tramp(ptr, count, op, user):
            try:
                return int(fn(ptr, count, op) or 0)
            except Exception as ex:  # never let an exception cross the C boundary
                print("allreduce hook failed:", ex)
                return 1
        self._ar = _abi.vrt_allreduce_fn(tramp)
        check(lib().vrt_solver_set_allreduce(self.h, self._ar, None))

    def comm_init(self, dir_id=None, dir_rank=0, dir_size=1, lam_id=None, lam_rank=0, lam_size=1):
        """in-library NCCL collectives (include/vrt.h vrt_solver_comm_init); ids are the 128 bytes of nccl_unique_id()"""
        check(lib().vrt_solver_comm_init(self.h, dir_id, int(dir_rank), int(dir_size), lam_id, int(lam_rank), int(lam_size)))

    def set_direction_lambda(self, direction, lam_begin, lam_end):
        """solve direction `direction` (index in this solver's table) on the local wavelengths [lam_begin, lam_end) only"""
        check(lib().vrt_solver_set_direction_lambda(self.h, int(direction), int(lam_begin), int(lam_end)))

    def direction_visits(self):
        """(cell, sweep) visits of each direction this solver holds (θ = 90 rows left out): the cost used to balance direction shards"""
        nd = C.c_int64()
        check(lib().vrt_solver_direction_visits(self.h, C.byref(nd), None, 0))
        out = (C.c_double * max(nd.value, 1))()
        check(lib().vrt_solver_direction_visits(self.h, C.byref(nd), out, nd.value))
        return np.array(out[:nd.value])

    def peer_handle(self):
        """64-byte CUDA IPC handle of this solver's J buffer (vrt_solver_peer_handle)"""
        raw = C.create_string_buffer(64)
        check(lib().vrt_solver_peer_handle(self.h, raw))
        return raw.raw

    def peer_attach(self, handles):
        """handles of the whole direction group in rank order (bytes, 64 per rank): J is then reduced through peer memory"""
        check(lib().vrt_solver_peer_attach(self.h, handles, len(handles) // 64))

    def peer_detach(self):
        """unmap the peers' J buffers; every process must do this (then synchronise) before any of them closes its solver"""
        check(lib().vrt_solver_peer_detach(self.h))

    def cell_slice(self):
        """[first, last) of the cells this process owns, in internal order (site perm_up[c])"""
        a, b = C.c_int64(), C.c_int64()
        check(lib().vrt_solver_cell_slice(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_state_slice(self):
        c0, c1 = self.cell_slice()
        S = np.zeros((self.nlam, c1 - c0), order="F")
        J = np.zeros((self.nlam, c1 - c0), order="F")
        pops = np.zeros((c1 - c0, 3), order="F") if self.kind == "line" else None
        check(lib().vrt_get_state_slice(self.h, _ptr(S), _ptr(J), _ptr(pops)))
        return S, J, pops

    def set_state_slice(self, S=None, populations=None):
        S_ = None if S is None else _f(S)
        P_ = None if populations is None else _f(populations)
        check(lib().vrt_set_state_slice(self.h, _ptr(S_), _ptr(P_)))

    def mean_intensity(self, S, populations=None, J=None, damping=None):
        """J_λ_voronoi on host arrays or device tensors (any of S, populations, J, damping may be a torch CUDA tensor)."""
        if J is None:
            J = np.zeros((self.nlam, self.n), order="F") if self.nlam > 1 or self.kind == "line" else np.zeros(self.n)
        S_ = S if hasattr(S, "data_ptr") else _f(S)
        P_ = populations if (populations is None or hasattr(populations, "data_ptr")) else _f(populations)
        check(lib().vrt_mean_intensity(self.h, _ptr(S_), _ptr(P_), _ptr(J), _ptr(damping)))
        return J

    def calculate_R(self, J, damping=None, R=None):
        if R is None:
            R = np.zeros((3, 3, self.n), order="F")
        J_ = J if hasattr(J, "data_ptr") else _f(J)
        D_ = damping if (damping is None or hasattr(damping, "data_ptr")) else _f(damping)
        check(lib().vrt_calculate_R(self.h, _ptr(J_), _ptr(D_), _ptr(R)))
        return R

    def iterate(self, ϵ, maxiter, callback=None):
        res = _abi.vrt_result()
        history = []

        def tramp(info, user):
            i = info.contents
            rec = {f: getattr(i, f) for f, _ in _abi.vrt_iter_info._fields_ if f != "reserved0"}
            history.append(rec)
            try:
                return int(callback(rec) or 0) if callback else 0
            except Exception as ex:
                print("iteration callback failed:", ex)
                return 1
        self._cb = _abi.vrt_iter_cb(tramp)
        check(lib().vrt_lambda_iterate(self.h, float(ϵ), int(maxiter), self._cb, None, C.byref(res)))
        return {"iterations": res.iterations, "converged": bool(res.converged), "diff": res.diff, "seconds": res.seconds,
                "history": history}

    def checksum(self):
        """device-side fingerprint of the state: ΣS, max|S|, Σ populations, ΣJ (own cells)"""
        out = (C.c_double * 4)()
        check(lib().vrt_state_checksum(self.h, out))
        return {"sum_S": out[0], "max_abs_S": out[1], "sum_populations": out[2], "sum_J_own_cells": out[3]}

    def get_state(self):
        shape = (self.nlam, self.n) if self.kind == "line" else (self.n,)
        S = np.zeros(shape, order="F")
        J = np.zeros(shape, order="F")
        pops = np.zeros((self.n, 3), order="F") if self.kind == "line" else None
        check(lib().vrt_get_state(self.h, _ptr(S), _ptr(J), _ptr(pops)))
        return S, J, pops

    def set_state(self, S=None, populations=None):
        S_ = None if S is None else _f(S)
        P_ = None if populations is None else _f(populations)
        check(lib().vrt_set_state(self.h, _ptr(S_), _ptr(P_)))


def _quad_key(quadrature):
    if isinstance(quadrature, (str, bytes)):
        return quadrature
    return tuple(np.asarray(quadrature[0]).tolist()) + tuple(np.asarray(quadrature[1]).tolist())


def _line_solver(sites, line, quadrature, **kw):
    # id(line) is a safe key: the cached Solver keeps `line` alive (Solver.line), so the id cannot be reused while the entry exists
    key = ("line", id(line), _quad_key(quadrature))
    s = sites._grid.solvers.get(key)
    if s is None:
        s = Solver(sites, quadrature, line=line, **kw)
        sites._grid.solvers[key] = s
    return s


def J_λ_voronoi(S_λ, α_cont, *args):
    """Line form (src/lambda_iteration.jl:60-113): J_λ_voronoi(S_λ, α_cont, populations, sites, line, quadrature) -> (J_λ, damping_λ).
    Continuum form (src/lambda_continuum.jl:27-56): J_λ_voronoi(S_λ, α_cont, sites, quadrature) -> J; the bottom boundary is
    blackbody_λ(500 nm, T) evaluated from sites.temperature."""
    if len(args) == 4:
        populations, sites, line, quadrature = args
        s = _line_solver(sites, line, quadrature, α_cont=α_cont)
        s.set_field("alpha_cont", α_cont)
        damping = np.zeros((s.nlam, s.n), order="F")
        J = s.mean_intensity(S_λ, populations, damping=damping)
        return J, damping
    sites, quadrature = args
    from .atom import B_λ
    B0 = B_λ(500.0, sites.temperature)
    s = Solver(sites, quadrature, α_cont=α_cont, ελ=np.ones(sites.n), B_0=B0)
    try:
        return s.mean_intensity(_f(S_λ).reshape(-1), J=np.zeros(sites.n))
    finally:
        s.close()


def calculate_R(sites, line, J_λ, damping_λ, LTE_pops, quadrature=None):
    """src/rates.jl:154-201 -> R (3, 3, n).  `quadrature` only selects which cached solver is reused."""
    s = None
    for key, cand in sites._grid.solvers.items():
        if key[0] == "line" and key[1] == id(line) and (quadrature is None or key[2] == _quad_key(quadrature)):
            s = cand
            break
    if s is None:
        s = _line_solver(sites, line, quadrature if quadrature is not None else quadrature_path("n1"))
    s.set_field("lte_pops", LTE_pops)
    return s.calculate_R(J_λ, damping_λ)


def get_revised_populations(R, C_rates, atom_density):
    """src/populations.jl:191-221 -> populations (n, 3)."""
    R, C_rates, NH = _f(R), _f(C_rates), _f(atom_density)
    n = NH.shape[0]
    pops = np.zeros((n, 3), order="F")
    check(lib().vrt_get_revised_populations(n, _ptr(R), _ptr(C_rates), _ptr(NH), _ptr(pops)))
    return pops


def Λ_voronoi(ϵ, maxiter, sites, *args, **kw):
    """Line form (src/lambda_iteration.jl:207-297): Λ_voronoi(ϵ, maxiter, sites, line, quadrature, DATA=None;
    α_cont, ελ, C, LTE_pops) -> (J, S, α_cont, populations).  The four keyword arrays are what the reference
    computes with Transparency.jl before its loop (:216-247).
    Continuum form (src/lambda_continuum.jl:109-160): Λ_voronoi(ϵ, maxiter, sites, quadrature; α_cont, ε_λ, B_0) -> (J, S, α_cont)."""
    callback = kw.pop("callback", None)
    if len(args) >= 2 and not isinstance(args[0], (str, bytes, tuple, list)):
        line, quadrature = args[0], args[1]
        s = Solver(sites, quadrature, line=line, α_cont=kw["α_cont"], ελ=kw["ελ"], C_rates=kw["C"], LTE_pops=kw["LTE_pops"],
                   **{k: v for k, v in kw.items() if k in ("n_sweeps", "p", "lam_range", "lam_chunk", "prune", "dir_range")})
        try:
            res = s.iterate(ϵ, maxiter, callback)
            S, J, pops = s.get_state()
        finally:
            s.close()
        Λ_voronoi.last = res
        return J, S, _f(kw["α_cont"]), pops
    quadrature = args[0]
    s = Solver(sites, quadrature, α_cont=kw["α_cont"], ελ=kw["ε_λ"], B_0=kw["B_0"])
    try:
        res = s.iterate(ϵ, maxiter, callback)
        S, J, _ = s.get_state()
    finally:
        s.close()
    Λ_voronoi.last = res
    return J, S, _f(kw["α_cont"])


def J_λ_regular(S_λ, α_cont, *args, I_0=None, I_0_down=None, n_sweeps=3):
    """Line form (src/lambda_iteration.jl:1-58): J_λ_regular(S_λ, α_cont, populations, atmos, line, quadrature) ->
    (J_λ, damping_λ), S_λ (nλ, nz, nx, ny), populations (nz, nx, ny, 3); the bottom boundary is B_λ(λ, T[1,:,:]) (:38).
    Continuum form (src/lambda_continuum.jl:1-24): J_λ_regular(S_λ, α_cont, atmos, quadrature, I_0=…).  S_λ, α_cont:
    (nz, nx, ny) or (nλ, nz, nx, ny); I_0: the bottom boundary the reference builds at :16, blackbody_λ(500 nm,
    T[1,:,:]), (nx, ny) or (nλ, nx, ny); rays with θ < 90 start from zero (:19) unless I_0_down is given."""
    if len(args) == 4:
        populations, atmos, line, quadrature = args
        s = _line_solver(atmos, line, quadrature, α_cont=α_cont)
        s.set_field("alpha_cont", α_cont)
        shape = (s.nlam,) + atmos.shape
        damping = np.zeros(shape, order="F")
        J = s.mean_intensity(S_λ, populations, J=np.zeros(shape, order="F"), damping=damping)
        return J, damping
    atmos, quadrature = args
    nz, nx, ny = len(atmos.z), len(atmos.x), len(atmos.y)
    S_λ, α_cont = _f(S_λ), _f(α_cont)
    nlam = 1 if S_λ.ndim == 3 else S_λ.shape[0]
    want = (nz, nx, ny) if S_λ.ndim == 3 else (nlam, nz, nx, ny)
    if S_λ.shape != want or α_cont.shape != want:
        raise ValueError("S_λ, α_cont must be (nz, nx, ny) or (nλ, nz, nx, ny)")
    bshape = want[:-3] + (nx, ny)
    up = None if I_0 is None else _f(I_0)
    dn = None if I_0_down is None else _f(I_0_down)
    for b in (up, dn):
        if b is not None and b.shape != bshape:
            raise ValueError("I_0 must be %s" % (bshape,))
    if up is None:
        up = np.zeros(bshape, order="F")
    q = _as_quadrature(quadrature)
    J = np.zeros_like(S_λ, order="F")
    check(lib().vrt_regular_mean_intensity(nz, nx, ny, _ptr(atmos.z), _ptr(atmos.x), _ptr(atmos.y), C.byref(q), int(n_sweeps), nlam,
                                           _ptr(S_λ), _ptr(α_cont), _ptr(up), _ptr(dn), _ptr(J)))
    return J


def Λ_regular(ϵ, maxiter, atmos, *args, n_sweeps=3, callback=None, **kw):
    """Line form (src/lambda_iteration.jl:116-205): Λ_regular(ϵ, maxiter, atmos, line, quadrature, DATA=None; α_cont, ελ, C,
    LTE_pops) -> (J, S, α_cont, populations) with J, S (nλ, nz, nx, ny) and populations (nz, nx, ny, 3); the keyword arrays
    are what the reference computes with Transparency.jl before its loop (:123-155).
    Continuum form (src/lambda_continuum.jl:58-107): Λ_regular(ϵ, maxiter, atmos, quadrature, α_cont, ε_λ, B_0) ->
    (J, S, α_cont); α_cont, ε_λ, B_0 (nz, nx, ny) as computed at :66-85."""
    if len(args) >= 2 and not isinstance(args[0], (str, bytes, tuple, list)):
        line, quadrature = args[0], args[1]
        s = Solver(atmos, quadrature, line=line, α_cont=kw["α_cont"], ελ=kw["ελ"], C_rates=kw["C"], LTE_pops=kw["LTE_pops"],
                   n_sweeps=n_sweeps, **{k: v for k, v in kw.items() if k in ("lam_chunk",)})
        try:
            res = s.iterate(ϵ, maxiter, callback)
            S, J, pops = s.get_state()
        finally:
            s.close()
        Λ_regular.last = res
        shape = (s.nlam,) + atmos.shape
        return (J.reshape(shape, order="F"), S.reshape(shape, order="F"), _f(kw["α_cont"]),
                pops.reshape(atmos.shape + (3,), order="F"))
    quadrature = args[0]
    α_cont = args[1] if len(args) > 1 else kw["α_cont"]
    ε_λ = args[2] if len(args) > 2 else kw["ε_λ"]
    B_0 = args[3] if len(args) > 3 else kw["B_0"]
    nz, nx, ny = len(atmos.z), len(atmos.x), len(atmos.y)
    α_cont, ε_λ, B_0 = _f(α_cont), _f(ε_λ), _f(B_0)
    for a in (α_cont, ε_λ, B_0):
        if a.shape != (nz, nx, ny):
            raise ValueError("α_cont, ε_λ, B_0 must be (nz, nx, ny)")
    q = _as_quadrature(quadrature)
    S = np.zeros((nz, nx, ny), order="F")
    J = np.zeros((nz, nx, ny), order="F")
    res = _abi.vrt_result()
    history = []

    def tramp(info, user):
        i = info.contents
        rec = {f: getattr(i, f) for f, _ in _abi.vrt_iter_info._fields_ if f != "reserved0"}
        history.append(rec)
        try:
            return int(callback(rec) or 0) if callback else 0
        except Exception as ex:
            print("iteration callback failed:", ex)
            return 1
    cb = _abi.vrt_iter_cb(tramp)
    check(lib().vrt_regular_lambda_iterate(nz, nx, ny, _ptr(atmos.z), _ptr(atmos.x), _ptr(atmos.y), C.byref(q), int(n_sweeps),
                                           _ptr(α_cont), _ptr(ε_λ), _ptr(B_0), float(ϵ), int(maxiter), cb, None, _ptr(S), _ptr(J),
                                           C.byref(res)))
    Λ_regular.last = {"iterations": res.iterations, "converged": bool(res.converged), "diff": res.diff, "seconds": res.seconds,
                      "history": history}
    return J, S, α_cont


# ASCII aliases
J_lambda_voronoi = J_λ_voronoi
Lambda_voronoi = Λ_voronoi
J_lambda_regular = J_λ_regular
Lambda_regular = Λ_regular
