"""ctypes front end of oracle/libvrt_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.  The
product package (voronoirt_b200/) must never do so.  Function names follow the reference (see
vrt_oracle.c for the file:line citations).  PARITY UNPINNED — see the header of vrt_oracle.c.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from voronoirt_b200._abi import vrt_line, vrt_quadrature, vrt_site_data  # noqa: E402  (struct layouts only)

_LIB = None


def _cpu_key():
    """the oracle is compiled -march=native, so its file name carries the CPU it was built for"""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        model = next((ln for ln in txt.splitlines() if ln.startswith("model name")), "")
        flags = next((ln for ln in txt.splitlines() if ln.startswith("flags")), "")
        return hashlib.sha1((model + flags).encode()).hexdigest()[:10]
    except OSError:
        return "generic"


def lib_path():
    return os.path.join(_HERE, f"libvrt_oracle_{_cpu_key()}.so")


def build():
    subprocess.run(["make", "-s", "-C", _HERE, f"LIB={os.path.basename(lib_path())}"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = lib_path()
        src = [os.path.join(_HERE, f) for f in ("vrt_oracle.c", "vrt_oracle_regular.c")]
        if not os.path.exists(path) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in src):
            build()
        L = C.CDLL(path)
        L.orc_read_neighbours.restype = C.c_int64
        L.orc_sites_create.restype = C.c_void_p
        L.orc_sites_num_layers.restype = C.c_int64
        L.orc_sites_max_nb.restype = C.c_int64
        L.orc_humlicek_re.restype = C.c_double
        L.orc_voigt_profile.restype = C.c_double
        L.orc_B_lambda.restype = C.c_double
        L.orc_criterion.restype = C.c_double
        L.orc_humlicek_re.argtypes = [C.c_double, C.c_double]
        L.orc_voigt_profile.argtypes = [C.c_double] * 3
        L.orc_B_lambda.argtypes = [C.c_double] * 2
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def read_neighbours(fname, n):
    """voronoi_utils.jl:42-70 -> NeighbourMatrix as an (ld, n) C array == Julia n x ld column-major."""
    L = lib()
    ld = L.orc_read_neighbours(fname.encode(), C.c_int64(n), None, C.c_int64(0))
    if ld < 0:
        raise IOError(fname)
    nbr = np.zeros((ld, n), dtype=np.int64)
    L.orc_read_neighbours(fname.encode(), C.c_int64(n), _p(nbr), C.c_int64(ld))
    return nbr


class Sites:
    """VoronoiSites' grid part built by the oracle's read_cell restatement.

    positions: (n, 3) array with columns (z, x, y) == Julia 3 x n column-major.
    nbr: (ld, n) int64 == Julia n x ld column-major.  bounds: (z_min, z_max, x_min, x_max, y_min, y_max).
    """

    def __init__(self, positions, nbr, bounds):
        self.positions = f64(positions)
        self.nbr = i64(nbr)
        self.bounds = f64(bounds)
        self.n = self.positions.shape[0]
        self.ld = self.nbr.shape[0]
        L = lib()
        self.h = L.orc_sites_create(C.c_int64(self.n), _p(self.positions), _p(self.nbr), C.c_int64(self.ld), _p(self.bounds))
        if not self.h:
            raise RuntimeError("oracle: neighbour graph not connected to a wall")
        self.h = C.c_void_p(self.h)
        self.max_nb = L.orc_sites_max_nb(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_sites_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def layers(self, down):
        L = lib()
        nl = L.orc_sites_num_layers(self.h, C.c_int(down))
        perm = np.zeros(self.n, dtype=np.int64)
        off = np.zeros(nl + 1, dtype=np.int64)
        L.orc_sites_get_layers(self.h, C.c_int(down), _p(perm), _p(off))
        return perm, off

    def delaunay_lines(self):
        out = np.zeros((self.n, self.max_nb, 3))
        lib().orc_sites_get_lines(self.h, _p(out))
        return out

    def stencil(self, k, p=7.0):
        k = f64(k)
        up = np.zeros((self.n, 2), dtype=np.int64)
        dots = np.zeros((self.n, 2))
        w = np.zeros((self.n, 2))
        r = np.zeros((self.n, 2))
        lib().orc_stencil(self.h, _p(k), C.c_double(p), _p(up), _p(dots), _p(w), _p(r))
        return up, dots, w, r

    def formal_solve(self, k, down, S, alpha, I0, n_sweeps=3, p=7.0, hoist=1):
        """Delaunay_upII/downII batched over wavelengths. S, alpha: (n, nlam); I0: (n1, nlam) -> I (n, nlam)."""
        k = f64(k)
        S = f64(S).reshape(self.n, -1)
        nlam = S.shape[1]
        alpha = f64(alpha).reshape(self.n, nlam)
        I0 = f64(I0).reshape(-1, nlam)
        out = np.zeros((self.n, nlam))
        lib().orc_formal_solve(self.h, _p(k), C.c_int(down), C.c_double(p), C.c_int(n_sweeps), C.c_int64(nlam),
                               _p(S), _p(alpha), _p(I0), _p(out), C.c_int(hoist))
        return out


def make_quadrature(weights, theta, phi):
    w, t, ph = f64(weights), f64(theta), f64(phi)
    q = vrt_quadrature(len(w), w.ctypes.data, t.ctypes.data, ph.ctypes.data)
    q._keep = (w, t, ph)
    return q


def make_site_data(**kw):
    sd = vrt_site_data()
    keep = []
    for name, _ in vrt_site_data._fields_:
        a = kw.get(name)
        if a is not None:
            a = f64(a)
            keep.append(a)
            setattr(sd, name, a.ctypes.data)
    sd._keep = keep
    return sd


def J_lambda_voronoi(sites, line, lam, sd, quad, S, pops, n_sweeps=3, p=7.0, l0=0, l1=0, hoist=1):
    """lambda_iteration.jl:60-113.  S (n, nlam), pops (3, n) == Julia n x 3 -> (J, damping) both (n, nlam)."""
    lam = f64(lam)
    S = f64(S)
    pops = f64(pops)
    J = np.zeros_like(S)
    damping = np.zeros_like(S)
    lib().orc_J_lambda_voronoi(sites.h, C.byref(line), _p(lam), C.byref(sd), C.byref(quad), C.c_int(n_sweeps), C.c_double(p),
                               _p(S), _p(pops), _p(J), _p(damping), C.c_int64(l0), C.c_int64(l1), C.c_int(hoist))
    return J, damping


def J_continuum(sites, quad, S, alpha, B0, n_sweeps=3, p=7.0, hoist=1):
    S, alpha, B0 = f64(S), f64(alpha), f64(B0)
    J = np.zeros_like(S)
    lib().orc_J_continuum(sites.h, C.byref(quad), C.c_int(n_sweeps), C.c_double(p), _p(S), _p(alpha), _p(B0), _p(J), C.c_int(hoist))
    return J


def calculate_R(line, lam, T, dD, J, damping, lte):
    """rates.jl:154-201.  J, damping (n, nlam); lte (3, n) -> R (n, 3, 3) with R[i, b, a] = R_julia[a+1, b+1, i+1]."""
    lam, T, dD, J, damping, lte = map(f64, (lam, T, dD, J, damping, lte))
    n = T.shape[0]
    R = np.zeros((n, 3, 3))
    lib().orc_calculate_R(C.byref(line), _p(lam), C.c_int64(n), _p(T), _p(dD), _p(J), _p(damping), _p(lte), _p(R))
    return R


def get_revised_populations(R, Cm, NH):
    R, Cm, NH = map(f64, (R, Cm, NH))
    n = NH.shape[0]
    pops = np.zeros((3, n))
    lib().orc_get_revised_populations(C.c_int64(n), _p(R), _p(Cm), _p(NH), _p(pops))
    return pops


def LTE_populations(line, T, ne, NH):
    T, ne, NH = map(f64, (T, ne, NH))
    n = T.shape[0]
    pops = np.zeros((3, n))
    lib().orc_LTE_populations(C.byref(line), C.c_int64(n), _p(T), _p(ne), _p(NH), _p(pops))
    return pops


def lambda_voronoi(sites, line, lam, sd, quad, S0, pops0, eps=1e-3, maxiter=150, n_sweeps=3, p=7.0, hoist=1):
    """lambda_iteration.jl:207-297 -> (J, S, pops, convergence, iterations)."""
    lam = f64(lam)
    S = f64(S0).copy()
    pops = f64(pops0).copy()
    J = np.zeros_like(S)
    conv = np.zeros(maxiter + 1)
    it = lib().orc_lambda_voronoi(sites.h, C.byref(line), _p(lam), C.byref(sd), C.byref(quad), C.c_int(n_sweeps), C.c_double(p),
                                  C.c_double(eps), C.c_int(maxiter), _p(S), _p(J), _p(pops), _p(conv), C.c_int(hoist))
    return J, S, pops, conv, it


def lambda_continuum(sites, quad, alpha, eps_l, B0, eps=1e-3, maxiter=150, n_sweeps=3, p=7.0, hoist=1):
    """lambda_continuum.jl:109-160 -> (J, S, convergence, iterations)."""
    alpha, eps_l, B0 = map(f64, (alpha, eps_l, B0))
    S = B0.copy()
    J = np.zeros_like(S)
    conv = np.zeros(maxiter + 1)
    it = lib().orc_lambda_continuum(sites.h, C.byref(quad), C.c_int(n_sweeps), C.c_double(p), C.c_double(eps), C.c_int(maxiter),
                                    _p(alpha), _p(eps_l), _p(B0), _p(S), _p(J), _p(conv), C.c_int(hoist))
    return J, S, conv, it


def voigt_profile(a, v, dD):
    return lib().orc_voigt_profile(a, v, dD)


def humlicek_re(a, v):
    return lib().orc_humlicek_re(a, v)


def B_lambda(lam_nm, T):
    return lib().orc_B_lambda(lam_nm, T)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """explicit OpenMP thread count (torchrun exports OMP_NUM_THREADS=1)"""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def short_characteristics(z, x, y, k, down, S, I_0, alpha, n_sweeps=3):
    """characteristics.jl:19-180.  S, alpha: (nz, nx, ny) Fortran-ordered (Julia layout, ghost columns included);
    I_0: (nx, ny) -> (I (nz, nx, ny) Fortran-ordered, plane branch per z (1 xy, 2 yz, 3 xz))"""
    z, x, y, k = map(f64, (z, x, y, k))
    S = np.asfortranarray(S, dtype=np.float64)
    alpha = np.asfortranarray(alpha, dtype=np.float64)
    I_0 = np.asfortranarray(I_0, dtype=np.float64)
    nz, nx, ny = S.shape
    out = np.zeros((nz, nx, ny), order="F")
    planes = np.zeros(nz, dtype=np.int32)
    lib().orc_short_characteristics(C.c_int64(nz), C.c_int64(nx), C.c_int64(ny), _p(z), _p(x), _p(y), _p(k), C.c_int(down),
                                    _p(S), _p(I_0), _p(alpha), C.c_int(n_sweeps), _p(out), _p(planes))
    return out, planes


def J_regular(z, x, y, weights, theta, phi, S, alpha, I0_up, n_sweeps=3):
    """lambda_continuum.jl:1-24 for one wavelength.  S, alpha (nz, nx, ny) Fortran-ordered; I0_up (nx, ny)."""
    z, x, y, weights, theta, phi = map(f64, (z, x, y, weights, theta, phi))
    S = np.asfortranarray(S, dtype=np.float64)
    alpha = np.asfortranarray(alpha, dtype=np.float64)
    I0_up = np.asfortranarray(I0_up, dtype=np.float64)
    nz, nx, ny = S.shape
    J = np.zeros((nz, nx, ny), order="F")
    lib().orc_J_regular(C.c_int64(nz), C.c_int64(nx), C.c_int64(ny), _p(z), _p(x), _p(y), C.c_int64(len(weights)), _p(weights),
                        _p(theta), _p(phi), C.c_int(n_sweeps), _p(S), _p(alpha), _p(I0_up), _p(J))
    return J


def lambda_regular(z, x, y, weights, theta, phi, alpha, eps_l, B0, eps=1e-3, maxiter=150, n_sweeps=3):
    """lambda_continuum.jl:58-107 -> (J, S, convergence, iterations)"""
    z, x, y, weights, theta, phi = map(f64, (z, x, y, weights, theta, phi))
    alpha, eps_l, B0 = (np.asfortranarray(a, dtype=np.float64) for a in (alpha, eps_l, B0))
    nz, nx, ny = B0.shape
    S = np.zeros((nz, nx, ny), order="F")
    J = np.zeros((nz, nx, ny), order="F")
    conv = np.zeros(maxiter + 1)
    it = lib().orc_lambda_regular(C.c_int64(nz), C.c_int64(nx), C.c_int64(ny), _p(z), _p(x), _p(y), C.c_int64(len(weights)),
                                  _p(weights), _p(theta), _p(phi), C.c_int(n_sweeps), C.c_double(eps), C.c_int(maxiter),
                                  _p(alpha), _p(eps_l), _p(B0), _p(S), _p(J), _p(conv))
    return J, S, conv, it


def J_lambda_regular(z, x, y, line, lam, sd, quad, S, pops, n_sweeps=3):
    """lambda_iteration.jl:1-58.  S (n, nlam) with n = nz*nx*ny cells in column-major (nz, nx, ny) order; pops (3, n)."""
    z, x, y, lam = map(f64, (z, x, y, lam))
    S = f64(S)
    pops = f64(pops)
    J = np.zeros_like(S)
    damping = np.zeros_like(S)
    lib().orc_J_lambda_regular(C.c_int64(len(z)), C.c_int64(len(x)), C.c_int64(len(y)), _p(z), _p(x), _p(y), C.byref(line), _p(lam),
                               C.byref(sd), C.byref(quad), C.c_int(n_sweeps), _p(S), _p(pops), _p(J), _p(damping))
    return J, damping


def lambda_regular_line(z, x, y, line, lam, sd, quad, S0, pops0, eps=1e-3, maxiter=150, n_sweeps=3):
    """lambda_iteration.jl:116-205 -> (J, S, pops, convergence, iterations)"""
    z, x, y, lam = map(f64, (z, x, y, lam))
    S = f64(S0).copy()
    pops = f64(pops0).copy()
    J = np.zeros_like(S)
    conv = np.zeros(maxiter + 1)
    it = lib().orc_lambda_regular_line(C.c_int64(len(z)), C.c_int64(len(x)), C.c_int64(len(y)), _p(z), _p(x), _p(y), C.byref(line),
                                       _p(lam), C.byref(sd), C.byref(quad), C.c_int(n_sweeps), C.c_double(eps), C.c_int(maxiter),
                                       _p(S), _p(J), _p(pops), _p(conv))
    return J, S, pops, conv, it


def trilinear(z, x, y, vals, positions):
    """functions.jl:207-248 over the columns of positions (3, n) rows (z, x, y); vals (nz, nx, ny) -> (out (n,), n_outside)"""
    z, x, y = map(f64, (z, x, y))
    vals = np.asfortranarray(vals, dtype=np.float64)
    pos = np.asfortranarray(positions, dtype=np.float64)
    n = pos.shape[1]
    out = np.zeros(n)
    L = lib()
    L.orc_trilinear.restype = C.c_int64
    bad = L.orc_trilinear(C.c_int64(len(z)), C.c_int64(len(x)), C.c_int64(len(y)), _p(z), _p(x), _p(y), _p(vals), C.c_int64(n), _p(pos), _p(out))
    return out, int(bad)
