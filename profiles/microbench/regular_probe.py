"""First measurement of the regular-grid solver (BASELINE config 4 shape: 400 x 258 x 258, ghost columns
included) on one B200: one direction per branch, nlam wavelengths resident in HBM, against the CPU oracle on wavelength 0
of the same arrays (parity at full size + single-thread CPU time).  Usage: python profiles/microbench/regular_probe.py [nlam]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402  (checker + CPU baseline only)
import voronoirt_b200 as V  # noqa: E402
from voronoirt_b200 import _lib  # noqa: E402
from voronoirt_b200.api import _ptr  # noqa: E402

nlam = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nz, nx, ny = 400, 258, 258
dz = np.concatenate([np.full(300, 12e3), np.linspace(12e3, 60e3, nz - 301)])
z = np.concatenate([[-0.5e6], -0.5e6 + np.cumsum(dz)])
x = (np.arange(nx) - 1) * 23437.5
y = (np.arange(ny) - 1) * 23437.5
g = torch.Generator(device="cuda").manual_seed(2022)
shape = (ny, nx, nz, nlam)                      # C order == Julia (nlam, nz, nx, ny) column-major
S = torch.rand(shape, generator=g, device="cuda", dtype=torch.float64) + 0.1
zz = torch.tensor(z, device="cuda").view(1, 1, nz, 1)
alpha = torch.rand(shape, generator=g, device="cuda", dtype=torch.float64)
alpha += 0.5
alpha *= 1e-3 * torch.exp(-(zz + 0.5e6) / 4e5)
for a in (S, alpha):                            # periodic ghost columns
    a[0] = a[-2]; a[-1] = a[1]; a[:, 0] = a[:, -2]; a[:, -1] = a[:, 1]
I0 = torch.rand((ny, nx, nlam), generator=g, device="cuda", dtype=torch.float64)
I0[0] = I0[-2]; I0[-1] = I0[1]; I0[:, 0] = I0[:, -2]; I0[:, -1] = I0[:, 1]
out = torch.empty_like(S)
branch = np.zeros(nz, dtype=np.int32)
L = _lib.lib()
rows = []
DIRS = ((152.7, 315.5), (70.3, 346.4), (78.2, 55.4), (27.3, 135.5))
if os.environ.get('VRT_PROBE_DIRS'):
    DIRS = [DIRS[int(i)] for i in os.environ['VRT_PROBE_DIRS'].split(',')]
for theta, phi in DIRS:
    k = np.ascontiguousarray(V.direction(theta, phi))
    down = int(theta < 90)
    ms = []
    for rep in range(int(os.environ.get('VRT_PROBE_REPS', '3'))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.vrt_regular_formal_solve(nz, nx, ny, _ptr(z), _ptr(x), _ptr(y), _ptr(k), down, 3, nlam, _ptr(S), _ptr(alpha),
                                              _ptr(I0), _ptr(out), _ptr(branch)))
        torch.cuda.synchronize()
        st = _lib.last_stats()
        ms.append(((time.perf_counter() - t0) * 1e3, st["sweep_ms"], st["kernels"]))
    wall, plane_ms, launches = min(ms)
    if os.environ.get('VRT_PROBE_NO_CPU'):
        print(json.dumps({'theta': theta, 'gpu_wall_ms': wall, 'gpu_plane_loop_ms': plane_ms}), flush=True)
        continue
    # CPU oracle, wavelength 0
    S0 = np.asfortranarray(S[..., 0].cpu().numpy().transpose(2, 1, 0))
    a0 = np.asfortranarray(alpha[..., 0].cpu().numpy().transpose(2, 1, 0))
    i0 = np.asfortranarray(I0[..., 0].cpu().numpy().transpose(1, 0))
    t0 = time.perf_counter()
    ref, rb = O.short_characteristics(z, x, y, k, down, S0, i0, a0)
    cpu_s = time.perf_counter() - t0
    got = out[..., 0].cpu().numpy().transpose(2, 1, 0)
    err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
    upd = (nz - 1) * (nx - 2) * (ny - 2) * nlam
    rows.append({"theta": theta, "phi": phi, "planes_xy_yz_xz": [int((branch == b).sum()) for b in (1, 2, 3)],
                 "branches_equal_oracle": bool(np.array_equal(branch, rb)), "max_rel_err_vs_oracle_lam0": err,
                 "gpu_wall_ms": wall, "gpu_plane_loop_ms": plane_ms, "launches": launches, "nlam": nlam,
                 "gpu_updates_per_s": upd / (wall * 1e-3), "cpu_1thread_s_per_lam": cpu_s,
                 "cpu_updates_per_s": upd / nlam / cpu_s})
    print(json.dumps(rows[-1]), flush=True)
if os.environ.get('VRT_PROBE_J'):
    # J_λ_regular over all ul7n12 directions (BASELINE metric on the config-4 shape), everything resident in HBM
    quad = np.loadtxt(os.path.join(ROOT, "voronoirt_b200", "quadratures", "ul7n12.dat"))
    w, t, p = (np.ascontiguousarray(quad[:, i]) for i in range(3))
    from voronoirt_b200 import _abi
    q = _abi.vrt_quadrature(len(w), w.ctypes.data, t.ctypes.data, p.ctypes.data)
    Jt = torch.empty_like(S)
    best = None
    import ctypes
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.vrt_regular_mean_intensity(nz, nx, ny, _ptr(z), _ptr(x), _ptr(y), ctypes.byref(q), 3, nlam, _ptr(S), _ptr(alpha),
                                                _ptr(I0), None, _ptr(Jt)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = _lib.last_stats()
        best = (dt, st) if best is None or dt < best[0] else best
    upd = (nz - 1) * (nx - 2) * (ny - 2) * nlam * len(w)
    rec = {"what": "J_lambda_regular, ul7n12, %d wavelengths, 400x258x258" % nlam, "wall_ms": best[0] * 1e3, "plane_loops_ms": best[1]["sweep_ms"],
           "launches": best[1]["kernels"], "updates_per_s": upd / best[0], "J_mean": float(Jt.mean())}
    print(json.dumps(rec), flush=True)
    rows.append(rec)
if rows:
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "regular_probe_%d.json" % nlam), "w"), indent=1)
V.regular_release_workspace()
