"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol include/vrt.h declares,
struct layouts agree between the header and the ctypes mirror, compute entry points fail loudly without a device,
and the host-side logic (wavelength sharding, trapezoid regrouping, quadrature tables, grid tiling) is right."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|void|const char\*)\s+(vrt_[A-Za-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from voronoirt_b200 import _abi, _lib
    names = header_functions()
    assert len(names) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libvrt.so does not export {n}"
    assert sorted(_abi.PROTOTYPES) == names, "voronoirt_b200/_abi.py and include/vrt.h declare different entry points"
    assert _lib.lib().vrt_abi_version() == 4


def test_struct_layouts_match_the_header():
    """compile a tiny C program that prints sizeof/offsetof and compare with the ctypes mirror"""
    from voronoirt_b200 import _abi
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "vrt.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu\n", sizeof(vrt_line), sizeof(vrt_site_data), sizeof(vrt_quadrature), sizeof(vrt_config), sizeof(vrt_iter_info), sizeof(vrt_result));
 printf("%zu %zu %zu %zu %zu\n", offsetof(vrt_line, lambda0), offsetof(vrt_line, atom_weight), offsetof(vrt_line, c_quadratic_stark), offsetof(vrt_config, lam_chunk), offsetof(vrt_config, prune));
 return 0; }'''
    import tempfile
    d = tempfile.mkdtemp()
    open(os.path.join(d, "t.c"), "w").write(prog)
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
    out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    got = [int(v) for v in out]
    exp = [C.sizeof(_abi.vrt_line), C.sizeof(_abi.vrt_site_data), C.sizeof(_abi.vrt_quadrature), C.sizeof(_abi.vrt_config),
           C.sizeof(_abi.vrt_iter_info), C.sizeof(_abi.vrt_result), _abi.vrt_line.lambda0.offset, _abi.vrt_line.atom_weight.offset,
           _abi.vrt_line.c_quadratic_stark.offset, _abi.vrt_config.lam_chunk.offset, _abi.vrt_config.prune.offset]
    assert got == exp


def test_no_cpu_fallback_compute_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from voronoirt_b200 import _lib
    L = _lib.lib()
    pos = np.asfortranarray(np.random.default_rng(0).random((3, 4)))
    nbr = np.asfortranarray(np.array([[2, 2, -5], [2, 1, -6], [2, 1, -6], [2, 1, -5]], dtype=np.int64))
    b = np.array([0, 1, 0, 1, 0, 1.0])
    h = C.c_void_p()
    rc = L.vrt_grid_create(4, C.c_void_p(pos.ctypes.data), C.c_void_p(nbr.ctypes.data), 3, C.c_void_p(b.ctypes.data), C.byref(h))
    assert rc == -2 and b"CUDA" in L.vrt_last_error()      # VRT_E_CUDA, never a silent CPU path
    with pytest.raises(_lib.VRTError):
        _lib.check(rc)
    # the regular-grid entry as well
    import voronoirt_b200 as V
    ax = np.linspace(0, 1, 6)
    with pytest.raises(_lib.VRTError, match="CUDA"):
        V.short_characteristics_up(V.direction(160, 45), np.zeros((6, 6, 6)), np.zeros((6, 6)), np.zeros((6, 6, 6)), V.Atmosphere(ax, ax, ax))


def test_product_never_touches_the_oracle():
    """the product path must not import, link or execute anything under oracle/ (only the voro++ driver the reference ships)"""
    pkg = os.path.join(ROOT, "voronoirt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), encoding="utf-8").read()
                assert "libvrt_oracle" not in txt and "import oracle" not in txt and "vrt_oracle" not in txt, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libvrt.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_read_neighbours_parser(tmp_path):
    from voronoirt_b200 import read_neighbours
    f = tmp_path / "nb.txt"
    f.write_text("3 1 2 -5\n1 2 3 -6 -5\n2 1 3\n")           # voro++ prints lines in arbitrary id order
    nbr = read_neighbours(str(f), 3)
    assert nbr.shape == (3, 5)
    assert list(nbr[0]) == [4, 2, 3, -6, -5] and list(nbr[1]) == [2, 1, 3, 0, 0] and list(nbr[2]) == [3, 1, 2, -5, 0]


def test_quadrature_tables():
    from voronoirt_b200 import quadrature_path, read_quadrature
    for name, npts in (("n1", 1), ("ul2n3", 3), ("ul7n12", 12), ("ul9n20", 20)):
        w, th, ph, n = read_quadrature(quadrature_path(name))
        assert n == len(w) and abs(w.sum() - 1) < 1e-12
        if name != "n1":
            assert n == npts and abs((w * np.cos(np.deg2rad(th))).sum()) < 1e-12       # half up, half down


def test_shard_ranges_and_trapezoid_regrouping():
    import bench
    for nlam, world in ((91, 8), (91, 4), (91, 2), (91, 1), (19, 2), (7, 8)):
        r = [bench.shard_range(nlam, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nlam and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    # Σ_l (f_l + f_{l+1})(λ_{l+1}-λ_l) == Σ_l f_l W_l with the per-wavelength weights the rates kernel uses
    rng = np.random.default_rng(1)
    lam = np.sort(rng.random(20)) + 1
    f = rng.random(20)
    pair = sum((f[l] + f[l + 1]) * (lam[l + 1] - lam[l]) for l in range(19))
    W = np.zeros(20)
    W[:-1] += np.diff(lam)
    W[1:] += np.diff(lam)
    assert abs((f * W).sum() - pair) < 1e-14 * abs(pair)
    # and the shards' partial sums add up
    assert abs(sum((f[a:b] * W[a:b]).sum() for a, b in [(0, 7), (7, 13), (13, 20)]) - pair) < 1e-14 * abs(pair)


def test_tiled_grid_matches_a_direct_tessellation():
    from voronoirt_b200 import api, synth
    from conftest import load_grid
    if api.default_voro_exec() is None:
        pytest.skip("voro++ driver not available")
    pos, nbr, b = load_grid("grid_unit300")
    unit = dict(z_min=0.0, z_max=1.0, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0)
    P, N, B2 = synth.tile_grid(pos, nbr, 2, 3, unit)
    N2 = synth.voronoi_neighbours(P, bounds=B2)
    assert P.shape[1] == 1800 and B2["x_max"] == 2.0 and B2["y_max"] == 3.0
    for i in range(P.shape[1]):
        assert set(N[i, 1:N[i, 0] + 1]) == set(N2[i, 1:N2[i, 0] + 1])


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import oracle as O
    from conftest import load_grid, oracle_sites
    from voronoirt_b200 import atom, synth
    pos, nbr, b = load_grid("grid_unit300")
    n = pos.shape[1]
    a = synth.atmosphere(pos[0] * 1e7, pos[1] * 6e6, pos[2] * 6e6)
    line, lte, α_cont, ελ, Cr = synth.line_inputs(a["temperature"], a["electron_density"], a["hydrogen_density"], 10, 4)
    nlam = len(line.λ)
    lo, hi = bench.shard_range(nlam, world, rank)
    s = oracle_sites(O, np.asfortranarray(pos * 2e5), nbr, b * 2e5)
    sd = O.make_site_data(temperature=a["temperature"], electron_density=a["electron_density"], hydrogen_density=a["hydrogen_density"],
                          velocity_z=a["velocity_z"], velocity_x=a["velocity_x"], velocity_y=a["velocity_y"], doppler_width=line.ΔD,
                          alpha_cont=α_cont, destruction=ελ, C=np.ascontiguousarray(Cr.T), lte_pops=np.ascontiguousarray(lte.T))
    w, th, ph = np.array([0.5, 0.5]), np.array([160.0, 20.0]), np.array([45.0, 195.0])
    q_ = O.make_quadrature(w, th, ph)
    S = np.ascontiguousarray(atom.B_λ(line.λ[None, :], a["temperature"][:, None]))
    # every rank solves only its wavelength shard (what each GPU does), then the rates are all-reduced
    J, damping = O.J_lambda_voronoi(s, line.as_struct(), line.λ, sd, q_, S, lte.T, l0=lo, l1=hi)
    # per-wavelength trapezoid weights restricted to the shard == the kernel's partial sums
    Jm = np.zeros_like(J)
    Jm[:, lo:hi] = J[:, lo:hi]
    tJ = torch.from_numpy(Jm)
    dist.all_reduce(tJ)                               # emulates "J is owner-complete per wavelength"
    tD = torch.from_numpy(damping)                    # a shard only fills its own columns of the damping parameter, too
    dist.all_reduce(tD)
    R = O.calculate_R(line.as_struct(), line.λ, a["temperature"], line.ΔD, tJ.numpy(), tD.numpy(), lte.T)
    diff_local = torch.tensor([float(np.abs(J[:, lo:hi]).max())], dtype=torch.float64)
    dist.all_reduce(diff_local, op=dist.ReduceOp.MAX)
    if rank == 0:
        Jfull, dfull = O.J_lambda_voronoi(s, line.as_struct(), line.λ, sd, q_, S, lte.T)
        Rfull = O.calculate_R(line.as_struct(), line.λ, a["temperature"], line.ΔD, Jfull, dfull, lte.T)
        q.put((bool(np.array_equal(tJ.numpy(), Jfull)), float(np.abs(R - Rfull).max()), float(diff_local), float(np.abs(Jfull).max())))
    dist.destroy_process_group()


def test_wavelength_sharding_with_two_gloo_ranks():
    """world_size-2 gloo run of the N>1 host logic: shards by bench.shard_range, all-reduce of the shard results"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    same, rdiff, dmax, jmax = res
    assert same and rdiff == 0.0 and dmax == jmax


def test_c_host_example_compiles_and_links_against_the_library(tmp_path):
    """examples/lambda_checkpoint.c (a C host with the per-iteration checkpoint callback) is valid C against include/vrt.h and
    resolves every symbol it uses in libvrt.so (compile + link only: running it needs a device)"""
    from voronoirt_b200 import _lib
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    obj = tmp_path / "ex.o"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", os.path.join(ROOT, "examples", "lambda_checkpoint.c"),
                    "-o", str(obj)], check=True)
    main = tmp_path / "main.c"
    main.write_text('#include "vrt.h"\nint run_lambda(const char*, int64_t, const double*, const double*, const vrt_line*, const double*, const vrt_site_data*, '
                    'const vrt_quadrature*, const char*);\nint main(void){ return vrt_abi_version() == VRT_ABI_VERSION && (void*)run_lambda != 0 ? 0 : 1; }\n')
    exe = tmp_path / "ex"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(main), str(obj), "-L", libdir, "-lvrt", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0      # loads libvrt.so, the ABI version matches the header
