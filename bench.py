#!/usr/bin/env python
"""bench.py — headline benchmark of the irregular-grid hot path: NLTE line Λ-iterations on a synthetic
Bifrost-shaped Voronoi grid (BASELINE.json metric: cell·angle·freq updates/s per formal solution; s per NLTE
Λ-iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload nlte_16m|nlte_4m|nlte_1m|small] [--impl reference]

Default workload: `nlte_16m`, the configuration BASELINE.json's target is quoted on (>=16 M-site Voronoi NLTE line solve, ul9n20,
91 wavelengths); it fits one B200.  `nlte_1m` is configs[2] (1 M sites, ul7n12) and the workload of the ncu captures.

A "step" is one full Λ-iteration (opacity + formal solution over all directions and wavelengths + source update
+ radiative rates + statistical equilibrium + criterion).  `value` = n_sites·n_dirs·n_λ / (time per step), inputs
resident in HBM; `e2e` = the same through the C ABI with pinned HOST buffers (S and populations in and out, every
step).  N > 1 (torchrun): the quadrature directions (and, beyond the direction count, the wavelengths) are sharded over
the ranks, each holding the full grid; one NCCL all-reduce of J (plus the rates when wavelengths are sharded, plus the
scalar criterion) per iteration; the total problem is fixed ("strong" scaling).
--impl reference times the CPU oracle (a port of the reference's Julia algorithm; Julia itself is not installed)
on a bounded sample of the same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (base sites, tiles in x, tiles in y, quadrature, nλ_bb, nλ_bf)
    # *_native: sites sampled from the synthetic atmosphere cube and tessellated ON THE GPU (vrt_rejection_sampling,
    # vrt_voronoi_neighbours, vrt_trilinear: the reference's set-up pipeline without voro++), no tiling.
    "nlte_16m_native": (16000000, 1, 1, "ul9n20", 50, 20),   # DEFAULT: BASELINE configs[4] / the north-star target
    "nlte_4m_native": (4000000, 1, 1, "ul9n20", 50, 20),
    "nlte_1m_native": (1000000, 1, 1, "ul7n12", 50, 20),     # BASELINE configs[2] (λ-sharded at N > 1 with --shard lambda)
    # voro++ tessellation of `base` sites, tiled periodically in x and y (round-1 workloads, kept for comparison)
    "small": (20000, 1, 1, "ul7n12", 50, 20),
    "nlte_1m": (250000, 2, 2, "ul7n12", 50, 20),
    "nlte_1m_direct": (1000000, 1, 1, "ul7n12", 50, 20),
    "nlte_4m": (250000, 4, 4, "ul9n20", 50, 20),
    "nlte_16m": (250000, 8, 8, "ul9n20", 50, 20),
    # BASELINE configs[1]: 500 nm continuum Λ-iteration, 1 M sites, ul7n12, one wavelength
    "continuum_1m": (1000000, 1, 1, "ul7n12", 0, 0),
    # BASELINE configs[0]: searchlight beam test, 51^3 uniform sites, the ul7n12 directions one at a time (p = 7)
    "searchlight": (132651, 1, 1, "ul7n12", 0, 0),
    # BASELINE configs[3]: the regular-grid comparison solver, 256 x 256 x 400 (+ ghost columns), ul7n12, 91 wavelengths (N = 1 only)
    "regular_400": (0, 1, 1, "ul7n12", 50, 20),
}
NATIVE = ("nlte_16m_native", "nlte_4m_native", "nlte_1m_native", "continuum_1m")
CPU_ARM_SITES = 500000   # --impl reference on a *_native workload: sites of its voro++-tessellated sample of the same atmosphere


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    return O


def build_problem(workload, cpu_arm=False):
    """-> dict(pos, nbr, bounds, atm, n, quadrature path, ...).
    cpu_arm: the problem of `--impl reference`.  Its process must never load libvrt.so, so neighbour files are parsed by
    the oracle, and a *_native workload (whose 16 M sites only the GPU tessellates in seconds; voro++ needs ~0.5 h) is
    represented by CPU_ARM_SITES sites drawn from the same atmosphere cube by the same rule and tessellated by the
    reference's own voro++ driver."""
    from voronoirt_b200 import api, synth
    base, kx, ky, qname, nbb, nbf = WORKLOADS[workload]
    cache = os.environ.get("VRT_CACHE", os.path.join(ROOT, ".vrt_cache"))
    os.makedirs(cache, exist_ok=True)
    rank = int(os.environ.get("RANK", "0"))
    parse = None
    if cpu_arm:
        O = _oracle()
        parse = lambda fname, n: np.asfortranarray(O.read_neighbours(fname, n).T)   # noqa: E731
    common = dict(qpath=api.quadrature_path(qname), qname=qname, nbb=nbb, nbf=nbf)
    if workload == "searchlight":
        # compare_searchlight.jl:10-40: n_sites = 51^3 uniform in the unit box, seed 2022
        f = os.path.join(cache, "searchlight_132651.npz")
        if not os.path.exists(f):
            rng = np.random.default_rng(2022)
            pos = np.asfortranarray(rng.random((3, base)))
            B1 = dict(z_min=0.0, z_max=1.0, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0)
            if api.default_voro_exec() is not None:
                nbr = synth.voronoi_neighbours(pos, bounds=B1, parse=parse)
            elif not cpu_arm:
                nbr = api.voronoi_neighbours(pos, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
            else:
                raise FileNotFoundError("voro++ driver not staged (baseline/_ref/output_sites)")
            np.savez(f + ".tmp.npz", pos=pos, nbr=nbr.astype(np.int32))
            os.replace(f + ".tmp.npz", f)
        d = np.load(f)
        pos, nbr = np.asfortranarray(d["pos"]), np.asfortranarray(d["nbr"].astype(np.int64))
        return dict(pos=pos, nbr=nbr, bounds=dict(z_min=0.0, z_max=1.0, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0), atm=None,
                    n=pos.shape[1], tiles=(1, 1), base=base, **common)
    if workload in NATIVE and not cpu_arm:
        t = time.time()
        pos, atm = synth.native_sites(base, seed=2022)
        t1 = time.time()
        B = synth.BOX
        nbr = api.voronoi_neighbours(pos, B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"])
        log(f"sampled and initialised {base} sites on the GPU in {t1 - t:.1f}s (cube included), tessellated them in {time.time() - t1:.2f}s "
            f"(max {int(nbr[:, 0].max())} faces)")
        return dict(pos=pos, nbr=nbr, bounds=dict(synth.BOX), atm=atm, n=pos.shape[1], tiles=(1, 1), base=base, native=True, **common)
    if workload in NATIVE:   # CPU arm of a native workload
        nref = min(base, CPU_ARM_SITES)
        f = os.path.join(cache, f"cpu_arm_{nref}_seed2022.npz")
        if not os.path.exists(f):
            t = time.time()
            pos, atm = synth.native_sites_cpu(nref, seed=2022)
            nbr = synth.voronoi_neighbours(pos, parse=parse)
            np.savez(f + ".tmp.npz", pos=pos, nbr=nbr.astype(np.int32), **atm)
            os.replace(f + ".tmp.npz", f)
            log(f"CPU arm: sampled {nref} sites (numpy) and tessellated them with voro++ in {time.time() - t:.1f}s")
        d = np.load(f)
        pos, nbr = np.asfortranarray(d["pos"]), np.asfortranarray(d["nbr"].astype(np.int64))
        atm = {k: d[k] for k in synth.FIELDS}
        return dict(pos=pos, nbr=nbr, bounds=dict(synth.BOX), atm=atm, n=pos.shape[1], tiles=(1, 1), base=base, native=True,
                    spatial_sample=f"{nref} of {base} sites", **common)
    f = os.path.join(cache, f"base_{base}_seed2022.npz")
    if not os.path.exists(f):
        if rank == 0:
            t = time.time()
            pos = synth.sample_sites(base, seed=2022)
            if api.default_voro_exec() is not None:
                nbr = synth.voronoi_neighbours(pos, parse=parse)
                how = "voro++"
            elif not cpu_arm:   # the reference's driver was not staged (baseline/_ref/output_sites): same neighbour sets from the GPU
                B = synth.BOX
                nbr = api.voronoi_neighbours(pos, B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"])
                how = "vrt_voronoi_neighbours (voro++ driver not found)"
            else:
                raise FileNotFoundError("voro++ driver not staged (baseline/_ref/output_sites)")
            np.savez(f + ".tmp.npz", pos=pos, nbr=nbr.astype(np.int32))
            os.replace(f + ".tmp.npz", f)
            log(f"tessellated {base} base sites with {how} in {time.time() - t:.1f}s")
        else:
            while not os.path.exists(f):
                time.sleep(1.0)
            time.sleep(1.0)
    d = np.load(f)
    pos, nbr = np.asfortranarray(d["pos"]), np.asfortranarray(d["nbr"].astype(np.int64))
    atm = synth.atmosphere(pos[0], pos[1], pos[2])
    bounds = dict(synth.BOX)
    if kx * ky > 1:
        pos, nbr, bounds = synth.tile_grid(pos, nbr, kx, ky)
        atm = {k: np.tile(v, kx * ky) for k, v in atm.items()}
    return dict(pos=pos, nbr=nbr, bounds=bounds, atm=atm, n=pos.shape[1], tiles=(kx, ky), base=base, **common)


def shard_grid(world, ndirs):
    """world = D direction shards x G wavelength shards, D as large as the direction count allows"""
    D = world
    while D > 1 and (D > ndirs or world % D):
        D -= 1
    return D, world // D


def shard_range(nlam, world, rank):
    """contiguous balanced wavelength shards: the first nlam % world ranks get one more"""
    q, r = divmod(nlam, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class ClockSampler:
    """samples nvidia-smi clocks and throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    for p in (os.path.join(ROOT, "MEASURED_PEAKS.json"), "/root/repo/MEASURED_PEAKS.json"):
        if os.path.exists(p):
            try:
                return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            except Exception:
                pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_sample(P, line, inputs, threads, target_updates=2.0e8, hoist=0, keep_J=False):
    """times the CPU oracle (port of the reference algorithm, threads over wavelengths like lambda_iteration.jl:91)
    on a bounded sample: the first directions of the quadrature x a contiguous block of wavelengths around the line centre x
    all sites.  The OpenMP thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    O = _oracle()
    from voronoirt_b200 import api, atom
    threads = O.set_num_threads(max(1, threads))
    lte, α_cont, ελ, Cr = inputs
    ctx = P.get("_oracle_ctx")
    if ctx is None:      # the oracle's grid (read_cell restated: layers, permutations, Delaunay lines) is built once per problem
        atm = P["atm"]
        b = P["bounds"]
        bounds = [b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"]]
        t0 = time.time()
        osites = O.Sites(np.ascontiguousarray(P["pos"].T), np.ascontiguousarray(P["nbr"].T), bounds)
        sd = O.make_site_data(temperature=atm["temperature"], electron_density=atm["electron_density"], hydrogen_density=atm["hydrogen_density"],
                              velocity_z=atm["velocity_z"], velocity_x=atm["velocity_x"], velocity_y=atm["velocity_y"], doppler_width=line.ΔD,
                              alpha_cont=α_cont, destruction=ελ, C=np.ascontiguousarray(Cr.T), lte_pops=np.ascontiguousarray(lte.T))
        S = np.ascontiguousarray(atom.B_λ(line.λ[None, :], atm["temperature"][:, None]))
        ctx = P["_oracle_ctx"] = (osites, sd, S)
        log(f"CPU oracle: grid of {P['n']} sites built in {time.time() - t0:.1f}s")
    osites, sd, S = ctx
    w, th, ph, nq = api.read_quadrature(P["qpath"])
    nlam = len(line.λ)
    nl = min(nlam, threads)
    kd = int(min(nq, max(1, round(target_updates / (P["n"] * nl)))))      # bounded sample: first kd directions of the table
    w, th, ph, nq = w[:kd], th[:kd], ph[:kd], kd
    oq = O.make_quadrature(w, th, ph)
    ls = line.as_struct()
    l0 = max(0, line.λidx[1] // 2 - nl // 2)
    box = {}

    def run():
        t = time.perf_counter()
        J, _ = O.J_lambda_voronoi(osites, ls, line.λ, sd, oq, S, lte.T, l0=l0, l1=l0 + nl, hoist=hoist)
        dt = time.perf_counter() - t
        if keep_J:
            box["J"] = np.ascontiguousarray(J[:, l0:l0 + nl])     # (n, nl)
        return dt
    return run, dict(nlam_sample=nl, l0=l0, ndirs=int(nq), n=P["n"], threads=threads, quad=(w, th, ph), S=S, box=box)


def gpu_parity_on_sample(V, sites, P, line, inputs, info):
    """the CUDA path on exactly the shard the CPU baseline just solved (same directions, same wavelengths, same S and
    populations, all sites), through the C ABI with host buffers -> error measures against the oracle's J"""
    lte, α_cont, ελ, Cr = inputs
    Jo = info["box"].get("J")
    if Jo is None:
        return None
    l0, nl = info["l0"], info["nlam_sample"]
    solver = V.Solver(sites, info["quad"], line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte, lam_range=(l0, l0 + nl))
    try:
        S = np.asfortranarray(info["S"][:, l0:l0 + nl].T)          # (nl, n)
        Jg = solver.mean_intensity(S, lte).T                        # (n, nl)
    finally:
        solver.close()
    d = np.abs(Jg - Jo)
    scale = float(np.abs(Jo).max())
    pos = Jo > 1e-300
    return {"what": f"J of {info['ndirs']} direction(s) x {nl} wavelengths x {info['n']} sites: CUDA path (vrt_mean_intensity, host buffers) against the CPU oracle, same inputs",
            "max_rel_err": float(d.max() / scale), "max_pointwise_rel_err": float((d[pos] / Jo[pos]).max()) if pos.any() else 0.0,
            "tolerance": 1e-9, "ok": bool(d.max() / scale <= 1e-9), "sites": int(info["n"]), "n_values": int(Jo.size)}


METRIC = "cell*angle*freq updates/s per formal solution (NLTE Lambda-iteration)"


def reference_arm(args, W, K):
    """--impl reference: the CPU port of the reference's J_λ_voronoi on the host cores, bounded sample per step.  This process
    never loads libvrt.so (neighbour files are parsed by the oracle)."""
    from voronoirt_b200 import synth
    P = build_problem(args.workload, cpu_arm=True)
    atm = P["atm"]
    line, lte, α_cont, ελ, Cr = synth.line_inputs(atm["temperature"], atm["electron_density"], atm["hydrogen_density"], P["nbb"], P["nbf"])
    threads = os.cpu_count() or 1
    # size the sample for roughly 15 s per step: calibrate on one direction, then take as many directions as fit
    run1, info1 = cpu_reference_sample(P, line, (lte, α_cont, ελ, Cr), threads, target_updates=1.0)
    t1 = run1()
    per_update = t1 / (info1["n"] * info1["ndirs"] * info1["nlam_sample"])
    target = max(1.0, float(os.environ.get("VRT_CPU_STEP_SECONDS", "15")) / per_update)
    run, info = cpu_reference_sample(P, line, (lte, α_cont, ελ, Cr), threads, target_updates=target)
    for _ in range(min(W, 1)):
        run()
    ts = [run() for _ in range(K)]
    t = float(np.mean(ts))
    updates = info["n"] * info["ndirs"] * info["nlam_sample"]
    val = updates / t
    sample = (f"{info['nlam_sample']} of {len(line.λ)} wavelengths (index {info['l0']}..) x the first {info['ndirs']} of the quadrature's directions x {info['n']} sites"
              + (f" (spatial sample: {P['spatial_sample']} of the workload, drawn from the same atmosphere by the same rule and tessellated by the reference's voro++ driver)" if P.get("spatial_sample") else "")
              + ", stencil recomputed at every visit like irregular_ray_tracing.jl:50; C/OpenMP port of the Julia reference (Julia not installed), -O3 -march=native")
    out = {"impl": "reference", "metric": METRIC, "value": val,
           "unit": "updates/s", "n_gpus": args.gpus, "steps": K, "warmup": min(W, 1), "ms_per_step": 1e3 * t, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.workload, P, len(line.λ)), "sites": WORKLOADS[args.workload][0] * WORKLOADS[args.workload][1] * WORKLOADS[args.workload][2],
                      "quadrature": P["qname"], "n_lambda": len(line.λ)},
           "cpu_baseline": {"value": val, "unit": "updates/s", "cores": info["threads"], "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return _line(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("VRT_WORKLOAD", "nlte_16m_native"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shard", default=os.environ.get("VRT_SHARD", "auto"), choices=["auto", "lambda"],
                    help="N > 1: auto = direction shards first (then wavelengths); lambda = wavelength shards only (BASELINE configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    K = max(args.steps, 1)

    if args.workload.startswith("regular"):
        return main_regular(args, W, K)
    if args.workload == "searchlight":
        return main_searchlight(args, W, K)
    if args.workload.startswith("continuum"):
        return main_continuum(args, W, K)
    if args.impl == "reference":
        return reference_arm(args, W, K) if rank == 0 else 0

    import torch
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib, atom, synth
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().vrt_set_device(local_rank))
    dist = None
    nccl_log = None
    if world > 1:
        import torch.distributed as dist
        # NCCL's INFO log (rings / NVLS, which tells whether the switch reduces) goes to a file per rank, not to stdout:
        # rank 0 prints exactly one JSON line
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        nccl_log = os.path.join(ROOT, "gpurun_out", f"nccl_n{world}_rank%r.log".replace("%r", str(rank)))
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,GRAPH,ENV")
        os.environ.setdefault("NCCL_DEBUG_FILE", nccl_log)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t_setup = time.time()
    P = build_problem(args.workload)
    atm, b, n = P["atm"], P["bounds"], P["n"]
    line, lte, α_cont, ελ, Cr = synth.line_inputs(atm["temperature"], atm["electron_density"], atm["hydrogen_density"], P["nbb"], P["nbf"])
    nlam = len(line.λ)
    w, th, ph, nq = V.read_quadrature(P["qpath"])
    # multi-GPU decomposition: D direction shards x G wavelength shards (D*G = world).  Directions first: the sweep's cost
    # is per (cell, direction) visit, so fewer directions per GPU scales it, narrower wavelength rows barely do.
    D, G = (1, world) if args.shard == "lambda" else shard_grid(world, int(nq))
    di, gi = rank % D, rank // D
    lo, hi = shard_range(nlam, G, gi)
    cell = V.read_cell(P["nbr"], n, P["pos"], b["x_min"], b["x_max"], b["y_min"], b["y_max"])
    sites = V.VoronoiSites(*cell, atm["temperature"], atm["electron_density"], atm["hydrogen_density"], atm["velocity_z"],
                           atm["velocity_x"], atm["velocity_y"], b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"], n)
    ndirs = int(np.sum(th != 90))
    def make_solver(mine, pieces=()):
        """solver of this rank: the whole directions `mine` plus the wavelength parts `pieces` = [(direction, lo, hi)]"""
        rows = sorted(set(int(i) for i in mine) | set(int(p[0]) for p in pieces))
        my_quad = (w[rows], th[rows], ph[rows]) if D > 1 else P["qpath"]
        sol = V.Solver(sites, my_quad, line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte,
                       lam_range=(lo, hi) if G > 1 else None, dir_range=(0, len(rows)) if D > 1 else None,
                       cell_shard=(di, D) if D > 1 and not os.environ.get("VRT_NO_CELL_SHARD") else None)
        for (i, plo, phi) in pieces:
            sol.set_direction_lambda(rows.index(int(i)), plo, phi)
        return sol
    mine = round_robin(int(nq), D, di)
    pieces = []
    solver = make_solver(mine)
    balance = None
    if D > 1 and not os.environ.get("VRT_ROUND_ROBIN"):
        # balance the direction shards by the measured work of each direction (visits of its sweep program): every rank
        # reports the directions it built, the table is all-reduced, and the shards are re-dealt: whole directions longest
        # first, and — when the direction count does not divide by the shard count — the most expensive ones split by wavelength
        cost = torch.zeros(int(nq), dtype=torch.float64, device="cuda")
        cost[torch.as_tensor(mine, device="cuda")] = torch.as_tensor(solver.direction_visits(), device="cuda")
        dist.all_reduce(cost, op=dist.ReduceOp.MAX)
        costs = cost.cpu().numpy()
        rr_load = [float(costs[round_robin(int(nq), D, r)].sum()) for r in range(D)]
        sp = None if (G > 1 or os.environ.get("VRT_NO_SPLIT")) else split_assign(list(costs), D, nlam)
        if sp is not None:
            plan, load = sp
            new_mine, new_pieces = plan[di]
            how = f"{len(new_mine)} whole directions per shard (longest processing time first) + the {int(nq) % D} most expensive directions split by wavelength into {D // (int(nq) % D)} parts (vrt_solver_set_direction_lambda)"
        else:
            shards, load = lpt_assign(list(costs), D)
            new_mine, new_pieces = shards[di], []
            how = "longest processing time first on the visits of each direction's sweep program"
        balance = {"how": how, "max_over_mean_load": max(load) / (sum(load) / D), "round_robin_max_over_mean_load": max(rr_load) / (sum(rr_load) / D)}
        if not np.array_equal(new_mine, mine) or new_pieces:
            solver.close()
            _lib.check(_lib.lib().vrt_grid_release_schedules(sites._grid.h))
            mine, pieces = new_mine, new_pieces
            solver = make_solver(mine, pieces)
    coll = {"ms": 0.0, "bytes": 0}
    comm_how = None
    if world > 1:
        comm_how = attach_collectives(solver, dist, torch, local_rank, rank, world, D, G, di, gi, coll)
    log(f"setup {time.time() - t_setup:.1f}s: n={n} dirs={ndirs} nlam={nlam} shards: {D} direction x {G} wavelength; rank 0 has directions {[int(i) for i in mine]} + parts {[(int(a), int(b), int(c)) for a, b, c in pieces]} wavelengths [{lo},{hi}) layers up/down={len(sites.layers_up) - 1}/{len(sites.layers_down) - 1}")

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: state lives in HBM, K full Λ-iterations
    solver.iterate(-1.0, W)
    sync()
    coll["ms"], coll["bytes"] = 0.0, 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = solver.iterate(-1.0, K)
    e1.record()
    sync()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    stats = _lib.last_stats()
    hist = res["history"]
    checksum = solver.checksum()          # after W + K iterations from S = B, populations = LTE: must agree for every N
    checksum["iterations"] = W + K
    coll_ms, coll_bytes = coll["ms"] / K, coll["bytes"] / K
    tt = torch.tensor([ms, stats["sweep_ms"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max, sweep_ms_max = float(tt[0]), float(tt[1])
    updates_total = float(n) * ndirs * nlam          # all ranks together, per step
    value = updates_total / (ms_max / K / 1e3)

    # ---- roofline of the dominant kernel (k_sweep): algorithmic bytes / measured kernel time (CUDA events in the library)
    B_alg = 40.0 + 104.0 / nlam
    local_updates = float(n) * (int(np.sum(th[mine] != 90)) * (hi - lo) + sum(p[2] - p[1] for p in pieces))
    launches = max(stats["kernels"], 1)
    achieved = B_alg * local_updates * K / (stats["sweep_ms"] / 1e3) / 1e9 if stats["sweep_ms"] > 0 else 0.0
    peak, peak_src = measured_peak()
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_sweep_tma", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_update": B_alg,
                "sweep_ms_per_step": sweep_ms_max / K, "sweep_share_of_step": sweep_ms_max / ms_max,
                "whole_step_frac": B_alg * updates_total / (ms_max / K / 1e3) / 1e9 / peak / world,
                "cell_visits_per_step": stats["visits"] / K * (1.0), "dependent_steps_per_step": stats["steps"] / K}

    # ---- end to end through the C ABI with pinned host buffers: state in, one Λ-iteration, S/J/populations out
    # Per step: S and populations go host -> device (restart state, like recover_simulation.jl), one Λ-iteration runs, and S and
    # populations come back device -> host (what the reference writes to its HDF5 file after every iteration,
    # lambda_iteration.jl:280-281).  With cell shards (N > 1) every rank moves only its own cell slice of S and of the
    # populations: the slices of the other ranks arrive through the all-gather the iteration does anyway.
    e2e = None
    dt_local = float("inf")
    nl = hi - lo
    sliced = D > 1 and not os.environ.get("VRT_NO_CELL_SHARD") and not os.environ.get("VRT_E2E_FULL")
    c0, c1 = solver.cell_slice() if sliced else (0, n)
    ncell = c1 - c0
    try:
        if args.no_e2e:
            raise StopIteration
        import ctypes as C
        L = _lib.lib()
        hS = torch.empty((ncell, nl), dtype=torch.float64).pin_memory()
        hP = torch.empty((3, ncell), dtype=torch.float64).pin_memory()
        if sliced:
            _lib.check(L.vrt_get_state_slice(solver.h, C.c_void_p(hS.data_ptr()), None, C.c_void_p(hP.data_ptr())))
        else:
            _lib.check(L.vrt_get_state(solver.h, C.c_void_p(hS.data_ptr()), None, C.c_void_p(hP.data_ptr())))

        null_cb = _abi_null_cb()
        dbg = os.environ.get("VRT_DEBUG")

        def step():
            ta = time.perf_counter()
            if sliced:
                _lib.check(L.vrt_set_state_slice(solver.h, C.c_void_p(hS.data_ptr()), C.c_void_p(hP.data_ptr())))
            else:
                _lib.check(L.vrt_set_state(solver.h, C.c_void_p(hS.data_ptr()), C.c_void_p(hP.data_ptr())))
            tb = time.perf_counter()
            _lib.check(L.vrt_lambda_iterate(solver.h, -1.0, 1, null_cb, None, None))
            tc = time.perf_counter()
            if sliced:
                _lib.check(L.vrt_get_state_slice(solver.h, C.c_void_p(hS.data_ptr()), None, C.c_void_p(hP.data_ptr())))
            else:
                _lib.check(L.vrt_get_state(solver.h, C.c_void_p(hS.data_ptr()), None, C.c_void_p(hP.data_ptr())))
            if dbg:
                log(f"e2e step: set_state {1e3 * (tb - ta):.1f} ms, iterate {1e3 * (tc - tb):.1f} ms, get_state {1e3 * (time.perf_counter() - tc):.1f} ms")
        for _ in range(2):
            step()
        for _rep in range(2):      # best of two K-step repetitions (host-side jitter of the pinned copies is large)
            sync()
            t0 = time.perf_counter()
            for _ in range(K):
                step()
            sync()
            dt_local = min(dt_local, time.perf_counter() - t0)
    except StopIteration:
        pass
    except Exception as ex:   # e.g. not enough HBM left for the staging buffer: report no e2e rather than no bench line
        log(f"e2e leg failed: {ex}")
    if not args.no_e2e:
        tt = torch.tensor([dt_local], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        if np.isfinite(dt):
            per_rank = int(8 * (ncell * nl + 3 * ncell))
            e2e = {"value": updates_total / (dt / K), "unit": "updates/s", "h2d_bytes_per_step": per_rank * (world if sliced else 1),
                   "d2h_bytes_per_step": per_rank * (world if sliced else 1), "ms_per_step": 1e3 * dt / K,
                   "api": ("vrt_set_state_slice + vrt_lambda_iterate(1 iteration) + vrt_get_state_slice: every rank moves its own cell slice of S and the populations "
                           "through pinned host buffers (bytes are the sum over ranks)" if sliced else
                           "vrt_set_state(S, populations) + vrt_lambda_iterate(1 iteration) + vrt_get_state(S, populations) with pinned host buffers")
                          + "; best of 2 repetitions of K steps"}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload, and the CUDA path checked
    # against it on exactly that sample (parity at the benchmark's own size)
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        solver.close()      # free the work buffers: the parity solve below needs its own
        solver = None
        threads = os.cpu_count() or 1
        inputs = (lte, α_cont, ελ, Cr)
        run, info = cpu_reference_sample(P, line, inputs, threads, keep_J=True)
        t = run()
        upd = info["n"] * info["ndirs"] * info["nlam_sample"]
        cpu = {"value": upd / t, "unit": "updates/s", "cores": info["threads"], "kind": "port",
               "sample": f"{info['nlam_sample']} of {nlam} wavelengths x the first {info['ndirs']} of {int(nq)} directions x {info['n']} sites, one formal solution "
                         f"({t:.1f} s), stencil recomputed per visit like the reference (faithful); C/OpenMP port (Julia not installed), -O3 -march=native"}
        try:
            parity = gpu_parity_on_sample(V, sites, P, line, inputs, info)
        except Exception as ex:  # noqa: BLE001
            parity = {"error": str(ex)}
        # second flavour: stencil hoisted out of the wavelength loop (the fairer CPU bar)
        run_h, _ = cpu_reference_sample(P, line, inputs, threads, hoist=1)
        th_ = run_h()
        cpu["hoisted_value"] = upd / th_
        cpu["hoisted_note"] = f"same sample with the upwind stencil computed once per direction ({th_:.1f} s)"

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "updates/s",
               "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(args.workload, P, nlam), "sites": n, "quadrature": P["qname"], "n_dirs": ndirs, "n_lambda": nlam,
                          "parallelism": f"{D} direction shards x {G} wavelength shards; {comm_how}" if world > 1 else "single GPU", "direction_balance": balance, "l2_policy": "inputs larger than L2 "
                          f"(S+J+I+alpha = {8 * n * nlam * (2 + 2 * ndirs) / 1e9:.1f} GB)", "n_sweeps": 3, "p": 7.0,
                          "visit_order": os.environ.get("VRT_BLOCKS", "default") + "/" + os.environ.get("VRT_SLAB", "default")},
               "s_per_lambda_iteration": ms_max / K / 1e3, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "checksum": checksum,
               "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "collectives": {"ms_per_step": coll_ms, "bytes_per_step": coll_bytes, "how": comm_how, "nccl_log": nccl_log} if world > 1 else None,
               "stage_ms": {k: float(np.mean([h[k] for h in hist])) for k in ("t_opacity_ms", "t_sweep_ms", "t_source_ms", "t_rates_ms", "t_stateq_ms", "t_total_ms")} if hist else None}
        print(json.dumps(out))
    if solver is not None:
        if dist is not None:      # importers unmap the peers' J buffers before any exporter frees its own
            solver.peer_detach()
            dist.barrier()
        solver.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def lpt_assign(costs, D):
    """longest processing time first: directions in descending cost to the least loaded of D shards -> list of index arrays"""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * D
    out = [[] for _ in range(D)]
    for i in order:
        r = min(range(D), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += costs[i]
    return [np.array(sorted(v), dtype=np.int64) for v in out], load


def split_assign(costs, D, nlam, min_width=16):
    """Directions that do not divide evenly over D shards: the nq % D most expensive ones are split by wavelength into D / (nq % D)
    parts each (vrt_solver_set_direction_lambda), so that every shard gets nq // D whole directions plus one part — 20 directions on
    8 shards: two whole ones and one half each.  -> per shard (whole direction indices, [(direction, lam_lo, lam_hi)]), loads;
    None when the counts do not allow it (then lpt_assign)."""
    nq = len(costs)
    q, r = divmod(nq, D)
    if r == 0 or D % r != 0:
        return None
    parts = D // r
    if nlam // parts < min_width:
        return None
    order = sorted(range(nq), key=lambda i: (-costs[i], i))
    split, whole = order[:r], order[r:]
    shards, load = lpt_assign([costs[i] for i in whole], D)
    shards = [np.array(sorted(whole[j] for j in sh), dtype=np.int64) for sh in shards]
    # a pass over a fraction f of the wavelengths costs about 0.3 + 0.7 f of a full one (measured: half rows 0.65)
    frac = 1.0 / parts
    pieces = []
    for i in split:
        for k in range(parts):
            lo = (nlam * k) // parts
            hi = (nlam * (k + 1)) // parts
            pieces.append((costs[i] * (0.3 + 0.7 * frac), i, lo, hi))
    pieces.sort(key=lambda t: (-t[0], t[1], t[2]))
    extra = [[] for _ in range(D)]
    free = set(range(D))
    for c, i, lo, hi in pieces:
        rnk = min(free, key=lambda j: (load[j], j))
        free.discard(rnk)
        extra[rnk].append((i, lo, hi))
        load[rnk] += c
    return [(shards[j], extra[j]) for j in range(D)], load


def round_robin(nq, D, di):
    """rank di takes di, di+D, ...: neighbouring lines of the quadrature files are up/down pairs of similar inclination"""
    return np.arange(di, nq, D) if D > 1 else np.arange(nq)


def attach_collectives(solver, dist, torch, local_rank, rank, world, D, G, di, gi, coll):
    """N > 1: the collectives of the Λ-iteration.  Preferred: inside the library (vrt_solver_comm_init: ncclCommInitRank from a
    unique id broadcast here, reduce-scatter / all-gather / all-reduce issued by libvrt.so on its own stream).  Fallback
    (VRT_HOST_COLLECTIVES=1 or a library built without NCCL): the round-1 host hook through torch.distributed."""
    import ctypes as C
    from voronoirt_b200 import _lib
    L = _lib.lib()
    if not os.environ.get("VRT_HOST_COLLECTIVES") and L.vrt_nccl_available() == 1:
        from voronoirt_b200 import api

        def group_id(members):
            """a unique id made by the first member, broadcast to the others (torch.distributed only ferries 128 bytes)"""
            grp = dist.new_group(members)          # every rank must take part in every new_group call
            buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == members[0]:
                buf = torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8).cuda()
            if rank in members:
                dist.broadcast(buf, src=members[0], group=grp)
                return buf.cpu().numpy().tobytes(), grp
            return None, None
        dir_id = lam_id = my_dir_group = None
        for g in range(G):
            if D > 1:
                r, grp = group_id([g * D + d for d in range(D)])
                if g == gi:
                    dir_id, my_dir_group = r, grp
        for d in range(D):
            if G > 1:
                r, _ = group_id([g * D + d for g in range(G)])
                lam_id = r if d == di else lam_id
        solver.comm_init(dir_id, di, D, lam_id, gi, G)
        how_J = "reduce-scatter of J"
        if D > 1 and not os.environ.get("VRT_NO_PEER_REDUCE") and not os.environ.get("VRT_NO_CELL_SHARD"):
            # J reduced through peer memory: the CUDA IPC handles of the J buffers travel once (64 bytes per rank)
            mine_h = torch.frombuffer(bytearray(solver.peer_handle()), dtype=torch.uint8).cuda()
            allh = torch.zeros(64 * D, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allh, mine_h, group=my_dir_group)
            try:
                solver.peer_attach(allh.cpu().numpy().tobytes())
                how_J = "J reduced through peer memory (CUDA IPC over NVLink) inside the source-update kernel"
            except Exception as ex:  # noqa: BLE001
                log(f"peer-memory reduction not available ({ex}); using the NCCL reduce-scatter")
        return (f"collectives inside libvrt.so (ncclCommInitRank from a broadcast unique id, own stream): {how_J}, all-gather of S and the "
                "populations, all-reduce of the rates over wavelength shards, max of the criterion; source update, rates and statistical equilibrium sharded over cells")

    class _Dev:
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    # process groups: ranks sharing a direction shard (they differ in wavelength shard) and vice versa
    lam_groups = [dist.new_group([g * D + d for g in range(G)]) for d in range(D)]
    dir_groups = [dist.new_group([g * D + d for d in range(D)]) for g in range(G)]

    def allreduce(ptr, count, op):
        if op == 0 and G == 1:
            return 0        # the rates are already complete: this rank holds every wavelength
        if op == 2 and D == 1:
            return 0
        t = torch.as_tensor(_Dev(ptr, count), device=torch.device("cuda", local_rank))
        t0c = time.perf_counter()
        if op in (3, 4):      # cell slices over the direction group: slice `di` of D equal slices is this rank's
            sl = t[di * (count // D):(di + 1) * (count // D)]
            if op == 3:
                dist.reduce_scatter_tensor(sl, t, op=dist.ReduceOp.SUM, group=dir_groups[gi])
            else:
                dist.all_gather_into_tensor(t, sl, group=dir_groups[gi])
        elif op == 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elif op == 0:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=lam_groups[di])
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=dir_groups[gi])
        torch.cuda.synchronize()
        coll["ms"] += 1e3 * (time.perf_counter() - t0c)
        coll["bytes"] += 8 * count
        return 0
    solver.set_allreduce(allreduce)
    return "host hook: torch.distributed NCCL reduce-scatter of J + all-gather of S and the populations per iteration (vrt_solver_set_allreduce)"


def _libvrt_mapped():
    try:
        return "libvrt.so" in open("/proc/self/maps").read()
    except OSError:
        return None


def _line(out):
    if out.get("impl") == "reference":
        out["libvrt_mapped"] = _libvrt_mapped()     # the CPU arm must not touch the product library
    print(json.dumps(out))
    return 0


def main_continuum(args, W, K):
    """BASELINE configs[1]: continuum Λ-iteration (Λ_voronoi of lambda_continuum.jl:109-160) on 1 M sites, ul7n12, one
    wavelength.  A step is one Λ-iteration: 12 formal solutions, S = (1-ε)J + εB, criterion.  N > 1: replicas only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    from voronoirt_b200 import api, synth
    metric = "cell*angle*freq updates/s per formal solution (continuum Lambda-iteration)"
    B_alg = 144.0     # 40 + 104/nλ with nλ = 1
    cpu_arm = args.impl == "reference"
    P = build_problem(args.workload, cpu_arm=cpu_arm)
    atm, b, n = P["atm"], P["bounds"], P["n"]
    α, ε, B0 = synth.continuum_inputs(atm["temperature"], atm["electron_density"], atm["hydrogen_density"])
    w, th, ph, nq = api.read_quadrature(P["qpath"])
    ndirs = int(np.sum(th != 90))
    updates = float(n) * ndirs
    wname = (f"{args.workload}: continuum Lambda-iteration (500 nm), {WORKLOADS[args.workload][0]} Voronoi sites sampled from the synthetic Bifrost-shaped atmosphere, "
             f"{P['qname']}, one wavelength")

    def cpu_leg():
        O = _oracle()
        threads = O.set_num_threads(os.cpu_count() or 1)
        bounds = [b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"]]
        osites = O.Sites(np.ascontiguousarray(P["pos"].T), np.ascontiguousarray(P["nbr"].T), bounds)
        oq = O.make_quadrature(w, th, ph)
        box = {}

        def run():
            t = time.perf_counter()
            box["J"] = O.J_continuum(osites, oq, B0, α, B0, hoist=0)
            return time.perf_counter() - t
        return run, threads, box

    if cpu_arm:
        run, threads, _ = cpu_leg()
        for _ in range(min(W, 1)):
            run()
        t = float(np.mean([run() for _ in range(K)]))
        val = updates / t
        sample = (f"one J_λ_voronoi (all {ndirs} directions, serial over directions like lambda_continuum.jl:40; one wavelength so one thread does the work) on {n} sites"
                  + (f" (spatial sample: {P['spatial_sample']}, voro++ tessellation)" if P.get("spatial_sample") else "") + "; C port of the Julia reference, -O3 -march=native")
        return _line({"impl": "reference", "metric": metric, "value": val, "unit": "updates/s", "n_gpus": args.gpus, "steps": K, "warmup": min(W, 1),
                      "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": wname},
                      "cpu_baseline": {"value": val, "unit": "updates/s", "cores": 1, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})

    import torch
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib
    import ctypes as C
    torch.cuda.set_device(0)
    cell = V.read_cell(P["nbr"], n, P["pos"], b["x_min"], b["x_max"], b["y_min"], b["y_max"])
    sites = V.VoronoiSites(*cell, atm["temperature"], atm["electron_density"], atm["hydrogen_density"], atm["velocity_z"],
                           atm["velocity_x"], atm["velocity_y"], b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"], n)
    solver = V.Solver(sites, P["qpath"], α_cont=α, ελ=ε, B_0=B0)
    solver.iterate(-1.0, W)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = solver.iterate(-1.0, K)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    stats = _lib.last_stats()
    checksum = solver.checksum()
    peak, peak_src = measured_peak()
    achieved = B_alg * updates * K / (stats["sweep_ms"] / 1e3) / 1e9 if stats["sweep_ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_sweep (register path, rows of one double)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_update": B_alg,
                "sweep_ms_per_step": stats["sweep_ms"] / K, "sweep_share_of_step": stats["sweep_ms"] / ms,
                "note": "latency-bound: a row is a single double, the working set (1 M sites x a few arrays) sits in L2"}
    # e2e: S in from pinned host memory, one Λ-iteration, S and J back out
    e2e = None
    if not args.no_e2e:
        L = _lib.lib()
        hS = torch.empty(n, dtype=torch.float64).pin_memory()
        hJ = torch.empty(n, dtype=torch.float64).pin_memory()
        _lib.check(L.vrt_get_state(solver.h, C.c_void_p(hS.data_ptr()), None, None))
        cb = _abi_null_cb()

        def step():
            _lib.check(L.vrt_set_state(solver.h, C.c_void_p(hS.data_ptr()), None))
            _lib.check(L.vrt_lambda_iterate(solver.h, -1.0, 1, cb, None, None))
            _lib.check(L.vrt_get_state(solver.h, C.c_void_p(hS.data_ptr()), C.c_void_p(hJ.data_ptr()), None))
        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": updates / (dt / K), "unit": "updates/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 16 * n, "ms_per_step": 1e3 * dt / K,
               "api": "vrt_set_state(S) + vrt_lambda_iterate(1 iteration) + vrt_get_state(S, J) with pinned host buffers"}
    cpu = parity = None
    if not args.no_cpu_baseline:
        run, threads, box = cpu_leg()
        t = run()
        cpu = {"value": updates / t, "unit": "updates/s", "cores": 1, "kind": "port",
               "sample": f"one J_λ_voronoi of the same workload: all {ndirs} directions x {n} sites ({t:.1f} s); serial like lambda_continuum.jl:40 (one wavelength: nothing to thread over)"}
        Jg = solver.mean_intensity(B0, J=np.zeros(n))
        Jo = box["J"]
        parity = {"what": f"J of one J_λ_voronoi call ({ndirs} directions x {n} sites, S = B_0): CUDA path against the CPU oracle",
                  "max_rel_err": float(np.abs(Jg - Jo).max() / np.abs(Jo).max()), "tolerance": 1e-9}
        parity["ok"] = parity["max_rel_err"] <= 1e-9
    solver.close()
    hist = res["history"]
    return _line({"metric": metric, "value": updates / (ms / K / 1e3), "unit": "updates/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
                  "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": wname, "sites": n, "quadrature": P["qname"], "n_dirs": ndirs, "n_lambda": 1,
                             "parallelism": "single GPU" if world == 1 else "replicas only (rank 0 reports)",
                             "l2_policy": "working set smaller than L2 by nature of the workload (8 MB per array); 12 directions x 3 sweeps touch every array between two uses of a row"},
                  "s_per_lambda_iteration": ms / K / 1e3, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "checksum": checksum, "e2e": e2e,
                  "gpu_launches": int(max(stats["kernels"], 1)), "clocks": clocks,
                  "stage_ms": {k: float(np.mean([h[k] for h in hist])) for k in ("t_opacity_ms", "t_sweep_ms", "t_source_ms", "t_total_ms")} if hist else None})


def main_searchlight(args, W, K):
    """BASELINE configs[0]: searchlight beam test (compare_searchlight.jl:10-152): 51^3 uniform sites in the unit box, S = 0,
    α = 0, I_0 = 1 inside a disk of radius 0.1 on the boundary layer, the ul7n12 directions one at a time with p = 7.
    A step is the 12 formal solutions (Delaunay_upII / Delaunay_downII), each a separate call like the reference's loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from voronoirt_b200 import api
    metric = "cell*angle*freq updates/s per formal solution (searchlight)"
    cpu_arm = args.impl == "reference"
    P = build_problem(args.workload, cpu_arm=cpu_arm)
    n, b = P["n"], P["bounds"]
    w, th, ph, nq = api.read_quadrature(P["qpath"])
    dirs = [(t, p) for t, p in zip(th, ph) if t != 90]
    updates = float(n) * len(dirs)
    wname = f"searchlight: {n} uniform Voronoi sites (51^3, unit box), S = 0, alpha = 0, unit beam of radius 0.1, the {len(dirs)} directions of {P['qname']} one at a time, p = 7"
    R0 = 0.1

    def boundary(pos, perm, n1):
        c = perm[:n1] - 1
        return (((pos[1, c] - 0.5) ** 2 + (pos[2, c] - 0.5) ** 2) < R0 ** 2).astype(np.float64)

    def cpu_leg():
        O = _oracle()
        bounds = [b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"]]
        osites = O.Sites(np.ascontiguousarray(P["pos"].T), np.ascontiguousarray(P["nbr"].T), bounds)
        lay = [osites.layers(0), osites.layers(1)]
        box = {}

        def run():
            t = time.perf_counter()
            out = []
            for (t_, p_) in dirs:
                k = api.direction(t_, p_)
                down = int(t_ < 90)
                perm, off = lay[down]
                I0 = boundary(P["pos"], perm, off[1] - 1)
                out.append(osites.formal_solve(k, down, np.zeros(n), np.zeros(n), I0, hoist=0)[:, 0])
            box["I"] = out
            return time.perf_counter() - t
        return run, box

    if cpu_arm:
        run, _ = cpu_leg()
        for _ in range(min(W, 1)):
            run()
        t = float(np.mean([run() for _ in range(K)]))
        val = updates / t
        return _line({"impl": "reference", "metric": metric, "value": val, "unit": "updates/s", "n_gpus": args.gpus, "steps": K, "warmup": min(W, 1),
                      "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": wname},
                      "cpu_baseline": {"value": val, "unit": "updates/s", "cores": 1, "kind": "port",
                                       "sample": "the whole workload (12 directions, one thread: the reference's searchlight loop is serial); C port of Delaunay_upII/downII"},
                      "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})

    import torch
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib
    import ctypes as C
    torch.cuda.set_device(0)
    cell = V.read_cell(P["nbr"], n, P["pos"], b["x_min"], b["x_max"], b["y_min"], b["y_max"])
    z = np.zeros(n)
    sites = V.VoronoiSites(*cell, z, z, z, z, z, z, b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"], n)
    L = _lib.lib()
    gh = sites._grid.h
    work = []
    for (t_, p_) in dirs:
        k = np.ascontiguousarray(api.direction(t_, p_))
        down = int(t_ < 90)
        perm = sites.perm_down if down else sites.perm_up
        off = sites.layers_down if down else sites.layers_up
        I0 = boundary(P["pos"], perm, off[1] - 1)
        work.append((k, down, I0))
    dS, dA = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    dI = torch.empty(n, dtype=torch.float64, device="cuda")
    dI0 = [torch.from_numpy(I0).cuda() for (_, _, I0) in work]
    hS, hA = torch.zeros(n, dtype=torch.float64).pin_memory(), torch.zeros(n, dtype=torch.float64).pin_memory()
    hI = torch.empty(n, dtype=torch.float64).pin_memory()

    def step_dev():
        for (k, down, _), i0 in zip(work, dI0):
            _lib.check(L.vrt_formal_solve(gh, C.c_void_p(k.ctypes.data), down, 7.0, 3, 1, C.c_void_p(dS.data_ptr()), C.c_void_p(dA.data_ptr()),
                                          C.c_void_p(i0.data_ptr()), C.c_void_p(dI.data_ptr())))

    def step_host(keep=None):
        for (k, down, I0) in work:
            _lib.check(L.vrt_formal_solve(gh, C.c_void_p(k.ctypes.data), down, 7.0, 3, 1, C.c_void_p(hS.data_ptr()), C.c_void_p(hA.data_ptr()),
                                          C.c_void_p(I0.ctypes.data), C.c_void_p(hI.data_ptr())))
            if keep is not None:
                keep.append(hI.numpy().copy())
    for _ in range(W):
        step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sweep_ms, launches = 0.0, 0
    e0.record()
    for _ in range(K):
        for (k, down, _), i0 in zip(work, dI0):
            _lib.check(L.vrt_formal_solve(gh, C.c_void_p(k.ctypes.data), down, 7.0, 3, 1, C.c_void_p(dS.data_ptr()), C.c_void_p(dA.data_ptr()),
                                          C.c_void_p(i0.data_ptr()), C.c_void_p(dI.data_ptr())))
            st = _lib.last_stats()
            sweep_ms += st["sweep_ms"]
            launches += int(st["kernels"])
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    step_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        step_host()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    peak, peak_src = measured_peak()
    B_alg = 144.0
    achieved = B_alg * updates * K / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_sweep (register path, rows of one double)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_update": B_alg, "sweep_ms_per_step": sweep_ms / K,
                "sweep_share_of_step": sweep_ms / ms, "note": "133 k sites x one wavelength: launch- and latency-bound, the whole grid sits in L2"}
    e2e = {"value": updates / (dt / K), "unit": "updates/s", "h2d_bytes_per_step": int(len(dirs) * 8 * (2 * n + n // 50)), "d2h_bytes_per_step": int(len(dirs) * 8 * n),
           "ms_per_step": 1e3 * dt / K, "api": "vrt_formal_solve per direction with HOST buffers (S, alpha, I_0 in; I out), like the reference's loop over Delaunay_upII / Delaunay_downII"}
    cpu = parity = None
    if not args.no_cpu_baseline:
        run, box = cpu_leg()
        t = run()
        cpu = {"value": updates / t, "unit": "updates/s", "cores": 1, "kind": "port", "sample": f"the whole workload, serial like the reference's searchlight loop ({t:.1f} s)"}
        got = []
        step_host(got)
        err = max(float(np.abs(g - o).max()) for g, o in zip(got, box["I"]))
        # the beam keeps its unit amplitude scale (0 <= I <= 1): absolute = relative to max I_0
        parity = {"what": f"I of all {len(dirs)} directions x {n} sites: CUDA path (vrt_formal_solve, host buffers) against the CPU oracle", "max_abs_err_over_max_I0": err,
                  "tolerance": 1e-9, "ok": err <= 1e-9}
    return _line({"metric": metric, "value": updates / (ms / K / 1e3), "unit": "updates/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
                  "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": wname, "sites": n, "quadrature": P["qname"], "n_dirs": len(dirs), "n_lambda": 1, "parallelism": "single GPU (replicas only for N > 1)",
                             "l2_policy": "the workload is smaller than L2 by definition (133 k sites); every step rewrites all of I"},
                  "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": launches // max(K, 1), "clocks": clocks})


def regular_problem(nz, nx, ny, nbb, nbf):
    """Synthetic Bifrost-shaped atmosphere on the regular grid of BASELINE configs[3]: x, y uniform with periodic ghost
    columns, z stretched from 12 km to 60 km spacing.  -> axes, (nz, nx, ny) fields (column-major), line inputs."""
    from voronoirt_b200 import synth
    B = synth.BOX
    n_fine = (3 * nz) // 4
    dz = np.concatenate([np.full(n_fine, 12e3), np.linspace(12e3, 60e3, nz - 1 - n_fine)])
    z = np.concatenate([[B["z_min"]], B["z_min"] + np.cumsum(dz)])
    x = (np.arange(nx) - 1) * (B["x_max"] / (nx - 2))
    y = (np.arange(ny) - 1) * (B["y_max"] / (ny - 2))
    fields = {}
    # evaluate plane by plane in y to bound the temporaries; the atmosphere is periodic in x and y
    Zp, Xp = np.meshgrid(z, np.mod(x, B["x_max"]), indexing="ij")
    cols = {k: [] for k in ("temperature", "electron_density", "hydrogen_density", "velocity_z", "velocity_x", "velocity_y")}
    for iy in range(ny):
        a = synth.atmosphere(Zp.ravel(order="F"), Xp.ravel(order="F"), np.full(Zp.size, np.mod(y[iy], B["y_max"])))
        for k in cols:
            cols[k].append(np.asarray(a[k], dtype=np.float64))
    flat = {k: np.concatenate(v) for k, v in cols.items()}           # cell = iz + nz*(ix + nx*iy)
    for k, v in flat.items():
        f = v.reshape((nz, nx, ny), order="F")
        f[:, 0, :] = f[:, -2, :]; f[:, -1, :] = f[:, 1, :]
        f[:, :, 0] = f[:, :, -2]; f[:, :, -1] = f[:, :, 1]
        fields[k] = f
    line, lte, α_cont, ελ, Cr = synth.line_inputs(flat["temperature"], flat["electron_density"], flat["hydrogen_density"], nbb, nbf)
    return dict(z=z, x=x, y=y, shape=(nz, nx, ny), n=nz * nx * ny, fields=fields, flat=flat, line=line, lte=lte, α_cont=α_cont, ελ=ελ, C=Cr)


def main_regular(args, W, K):
    """BASELINE configs[3]: NLTE line Λ-iteration on the regular grid (Λ_regular, lambda_iteration.jl:116-205).  N > 1: the
    directions of the quadrature are sharded over the ranks (contiguous ranges), J is all-reduced inside the library."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    base, kx, ky, qname, nbb, nbf = WORKLOADS[args.workload]
    shape = tuple(int(v) for v in os.environ.get("VRT_REG_SHAPE", "400,258,258").split(","))
    from voronoirt_b200 import api
    qpath = api.quadrature_path(qname)
    w, th, ph, nq = api.read_quadrature(qpath)
    ndirs = int(np.sum(th != 90))
    metric = "cell*angle*freq updates/s per formal solution (NLTE Lambda-iteration)"

    def cpu_sample(threads):
        """the oracle's J_λ_regular (port of lambda_iteration.jl:1-58, threads over wavelengths like :30) on a narrower box of the
        same atmosphere: all wavelengths x one direction per ray routine (yz, xy, xz)"""
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        from voronoirt_b200 import atom
        nzs, nxs, nys = shape[0], min(shape[1], 66), min(shape[2], 66)
        Q = regular_problem(nzs, nxs, nys, nbb, nbf)
        line = Q["line"]
        sd = O.make_site_data(temperature=Q["flat"]["temperature"], electron_density=Q["flat"]["electron_density"],
                              hydrogen_density=Q["flat"]["hydrogen_density"], velocity_z=Q["flat"]["velocity_z"],
                              velocity_x=Q["flat"]["velocity_x"], velocity_y=Q["flat"]["velocity_y"], doppler_width=line.ΔD,
                              alpha_cont=Q["α_cont"], destruction=Q["ελ"], C=np.ascontiguousarray(Q["C"].T), lte_pops=np.ascontiguousarray(Q["lte"].T))
        pick = [0, 2, 8] if nq >= 9 else [0]
        oq = O.make_quadrature(w[pick], th[pick], ph[pick])
        S = np.ascontiguousarray(atom.B_λ(line.λ[None, :], Q["flat"]["temperature"][:, None]))

        def run():
            t = time.perf_counter()
            O.J_lambda_regular(Q["z"], Q["x"], Q["y"], line.as_struct(), line.λ, sd, oq, S, Q["lte"].T)
            return time.perf_counter() - t
        upd = float(nzs - 1) * (nxs - 2) * (nys - 2) * len(line.λ) * len(pick)
        sample = (f"all {len(line.λ)} wavelengths x {len(pick)} of {ndirs} directions (one per ray routine) on a {nzs} x {nxs} x {nys} box of the same "
                  "atmosphere; C/OpenMP port of J_λ_regular (Julia not installed)")
        return run, upd, sample, O.num_threads()

    if args.impl == "reference":
        run, upd, sample, threads = cpu_sample(os.cpu_count() or 1)
        for _ in range(min(W, 1)):
            run()
        t = float(np.mean([run() for _ in range(K)]))
        val = upd / t
        print(json.dumps({"impl": "reference", "metric": metric, "value": val, "unit": "updates/s", "n_gpus": args.gpus, "steps": K,
                          "warmup": min(W, 1), "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": {"workload": f"{args.workload}: regular grid {shape}, {qname}"},
                          "cpu_baseline": {"value": val, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().vrt_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t_setup = time.time()
    Q = regular_problem(*shape, nbb, nbf)
    line, n, f = Q["line"], Q["n"], Q["fields"]
    nlam = len(line.λ)
    atm = V.Atmosphere(Q["z"], Q["x"], Q["y"], f["temperature"], f["electron_density"], f["hydrogen_density"], f["velocity_z"],
                       f["velocity_x"], f["velocity_y"])
    lam_chunk = int(os.environ.get("VRT_REG_BENCH_CHUNK", "0"))   # 0: as many wavelengths per pass as the free HBM allows
    D = min(world, int(nq))
    dlo, dhi = shard_range(int(nq), D, rank % D)
    solver = V.Solver(atm, qpath, line=line, α_cont=Q["α_cont"], ελ=Q["ελ"], C_rates=Q["C"], LTE_pops=Q["lte"], lam_chunk=lam_chunk,
                      dir_range=(dlo, dhi) if world > 1 else None)
    comm_how = None
    if world > 1:
        comm_how = attach_collectives(solver, dist, torch, local_rank, rank, world, D, 1, rank % D, 0, {"ms": 0.0, "bytes": 0})
    log(f"setup {time.time() - t_setup:.1f}s: regular grid {shape} = {n} cells, dirs={ndirs} nlam={nlam}, rank 0 directions [{dlo},{dhi})")

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
    solver.iterate(-1.0, W)
    sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = solver.iterate(-1.0, K)
    e1.record()
    sync()
    clocks = sampler.stop()
    tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt[0])
    stats = _lib.last_stats()
    hist = res["history"]
    checksum = solver.checksum()
    if rank != 0 or world > 1:
        # N > 1: device-resident line only (the end-to-end and CPU legs are reported at N = 1)
        if rank == 0:
            interior = float(shape[0] - 1) * (shape[1] - 2) * (shape[2] - 2)
            updates = interior * ndirs * nlam
            print(json.dumps({"metric": metric, "value": updates / (ms / K / 1e3), "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
                              "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                              "config": {"workload": f"{args.workload}: NLTE line Lambda-iteration on the regular grid {shape}, {qname}, {nlam} wavelengths",
                                         "parallelism": f"{D} direction shards; {comm_how}"},
                              "checksum": checksum, "gpu_launches": int(max(stats["kernels"], 1)), "clocks": clocks, "e2e": None, "cpu_baseline": None,
                              "roofline": None}))
        solver.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0
    interior = float(shape[0] - 1) * (shape[1] - 2) * (shape[2] - 2)
    updates = interior * ndirs * nlam
    value = updates / (ms / K / 1e3)
    B_alg = 40.0            # S read 8, alpha read 8, I write 8, J read-modify-write 16 per (cell, direction, wavelength)
    achieved = B_alg * updates * K / (stats["sweep_ms"] / 1e3) / 1e9 if stats["sweep_ms"] > 0 else 0.0
    peak, peak_src = measured_peak()
    roofline = {"bound": "hbm", "kernel": "regular plane walk (k_reg_xy | k_reg_coef + k_reg_rec)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_update": B_alg,
                "sweep_ms_per_step": stats["sweep_ms"] / K, "sweep_share_of_step": stats["sweep_ms"] / ms,
                "note": "k_reg_rec is latency-bound (one CTA per wavelength), see DESIGN.md 5b"}
    e2e = None
    if not args.no_e2e:
        try:
            import ctypes as C
            L = _lib.lib()
            hS = torch.empty((n, nlam), dtype=torch.float64).pin_memory()
            hP = torch.empty((3, n), dtype=torch.float64).pin_memory()
            _lib.check(L.vrt_get_state(solver.h, C.c_void_p(hS.data_ptr()), None, C.c_void_p(hP.data_ptr())))
            null_cb = _abi_null_cb()

            def step():
                _lib.check(L.vrt_set_state(solver.h, C.c_void_p(hS.data_ptr()), C.c_void_p(hP.data_ptr())))
                _lib.check(L.vrt_lambda_iterate(solver.h, -1.0, 1, null_cb, None, None))
                _lib.check(L.vrt_get_state(solver.h, C.c_void_p(hS.data_ptr()), None, C.c_void_p(hP.data_ptr())))
            step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(K):
                step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e = {"value": updates / (dt / K), "unit": "updates/s", "h2d_bytes_per_step": int(8 * (n * nlam + 3 * n)),
                   "d2h_bytes_per_step": int(8 * (n * nlam + 3 * n)), "ms_per_step": 1e3 * dt / K,
                   "api": "vrt_set_state(S, populations) + vrt_lambda_iterate(1 iteration) + vrt_get_state(S, populations) on a regular-grid handle, pinned host buffers"}
        except Exception as ex:
            log(f"e2e leg failed: {ex}")
    cpu = None
    if not args.no_cpu_baseline:
        run, upd, sample, threads = cpu_sample(os.cpu_count() or 1)
        t = run()
        cpu = {"value": upd / t, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample + f" ({t:.1f} s)"}
    print(json.dumps({"metric": metric, "value": value, "unit": "updates/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"{args.workload}: NLTE line Lambda-iteration on the regular grid {shape[0]} x {shape[1]} x {shape[2]} (ghost columns included), "
                                             f"{qname}, {nlam} wavelengths, synthetic Bifrost-shaped atmosphere", "cells": n, "quadrature": qname,
                                 "n_dirs": ndirs, "n_lambda": nlam, "parallelism": "single GPU",
                                 "l2_policy": f"inputs larger than L2 (S+J+alpha = {8 * n * nlam * 3 / 1e9:.1f} GB)", "n_sweeps": 3},
                      "s_per_lambda_iteration": ms / K / 1e3, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                      "gpu_launches": int(max(stats["kernels"], 1)), "clocks": clocks,
                      "stage_ms": {k: float(np.mean([h[k] for h in hist])) for k in ("t_opacity_ms", "t_sweep_ms", "t_source_ms", "t_rates_ms", "t_stateq_ms", "t_total_ms")} if hist else None}))
    solver.close()
    return 0


def workload_name(key, P, nlam):
    kx, ky = P["tiles"]
    tile = f" (voro++ tessellation of {P['base']} sites tiled {kx}x{ky} periodically)" if kx * ky > 1 else ""
    if P.get("native"):
        tile = " (sampled from the atmosphere cube and tessellated on the GPU: vrt_rejection_sampling, vrt_voronoi_neighbours, vrt_trilinear)"
    nsites = WORKLOADS[key][0] * kx * ky
    return f"{key}: NLTE line Lambda-iteration, {nsites} Voronoi sites{tile}, {P['qname']}, {nlam} wavelengths, synthetic Bifrost-shaped atmosphere"


def _abi_null_cb():
    from voronoirt_b200 import _abi
    import ctypes as C
    return C.cast(None, _abi.vrt_iter_cb)


if __name__ == "__main__":
    sys.exit(main())
