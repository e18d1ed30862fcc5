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
    # name: (base sites tessellated by voro++, tiles in x, tiles in y, quadrature, nλ_bb, nλ_bf)
    "small": (20000, 1, 1, "ul7n12", 50, 20),
    "nlte_1m": (250000, 2, 2, "ul7n12", 50, 20),
    "nlte_1m_direct": (1000000, 1, 1, "ul7n12", 50, 20),   # 1 M sites tessellated directly (no tiling): slower set-up, same solve
    "nlte_4m": (250000, 4, 4, "ul9n20", 50, 20),
    "nlte_16m": (250000, 8, 8, "ul9n20", 50, 20),
    # directly sampled sites, tessellated on the GPU by vrt_voronoi_neighbours (no voro++, no tiling)
    "nlte_4m_native": (4000000, 1, 1, "ul9n20", 50, 20),
    "nlte_16m_native": (16000000, 1, 1, "ul9n20", 50, 20),
    # BASELINE configs[3]: the regular-grid comparison solver, 256 x 256 x 400 (+ ghost columns), ul7n12, 91 wavelengths (N = 1 only)
    "regular_400": (0, 1, 1, "ul7n12", 50, 20),
}


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


def build_problem(workload):
    """-> dict(pos, nbr, bounds, atm, n, quadrature path).  The base tessellation is cached under .vrt_cache/."""
    from voronoirt_b200 import api, synth
    base, kx, ky, qname, nbb, nbf = WORKLOADS[workload]
    cache = os.environ.get("VRT_CACHE", os.path.join(ROOT, ".vrt_cache"))
    os.makedirs(cache, exist_ok=True)
    f = os.path.join(cache, f"base_{base}_seed2022.npz")
    rank = int(os.environ.get("RANK", "0"))
    if workload.endswith("_native"):
        t = time.time()
        pos = synth.sample_sites(base, seed=2022)
        t1 = time.time()
        B = synth.BOX
        nbr = api.voronoi_neighbours(pos, B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"])
        log(f"sampled {base} sites in {t1 - t:.1f}s, tessellated them on the GPU in {time.time() - t1:.2f}s (max {int(nbr[:, 0].max())} faces)")
        atm = synth.atmosphere(pos[0], pos[1], pos[2])
        return dict(pos=pos, nbr=nbr, bounds=dict(synth.BOX), atm=atm, n=pos.shape[1], qpath=api.quadrature_path(qname), qname=qname,
                    nbb=nbb, nbf=nbf, tiles=(1, 1), base=base, native=True)
    if not os.path.exists(f):
        if rank == 0:
            t = time.time()
            pos = synth.sample_sites(base, seed=2022)
            if api.default_voro_exec() is not None:
                nbr = synth.voronoi_neighbours(pos)
                how = "voro++"
            else:   # the reference's driver was not staged (oracle/_ref/output_sites): same neighbour sets from the GPU
                B = synth.BOX
                nbr = api.voronoi_neighbours(pos, B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"])
                how = "vrt_voronoi_neighbours (voro++ driver not found)"
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
    return dict(pos=pos, nbr=nbr, bounds=bounds, atm=atm, n=pos.shape[1], qpath=api.quadrature_path(qname), qname=qname,
                nbb=nbb, nbf=nbf, tiles=(kx, ky), base=base)


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


def cpu_reference_sample(P, line, inputs, threads, target_updates=2.0e8, hoist=0):
    """times the CPU oracle (port of the reference algorithm, threads over wavelengths like lambda_iteration.jl:91)
    on a bounded sample: all directions x a contiguous block of wavelengths around the line centre."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from voronoirt_b200 import api, atom
    lte, α_cont, ελ, Cr = inputs
    atm = P["atm"]
    b = P["bounds"]
    bounds = [b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"]]
    osites = O.Sites(np.ascontiguousarray(P["pos"].T), np.ascontiguousarray(P["nbr"].T), bounds)
    sd = O.make_site_data(temperature=atm["temperature"], electron_density=atm["electron_density"], hydrogen_density=atm["hydrogen_density"],
                          velocity_z=atm["velocity_z"], velocity_x=atm["velocity_x"], velocity_y=atm["velocity_y"], doppler_width=line.ΔD,
                          alpha_cont=α_cont, destruction=ελ, C=np.ascontiguousarray(Cr.T), lte_pops=np.ascontiguousarray(lte.T))
    w, th, ph, nq = api.read_quadrature(P["qpath"])
    nlam = len(line.λ)
    nl_ = min(nlam, max(1, threads))
    kd = int(min(nq, max(1, round(target_updates / (P["n"] * nl_)))))      # bounded sample: first kd directions of the table
    w, th, ph, nq = w[:kd], th[:kd], ph[:kd], kd
    oq = O.make_quadrature(w, th, ph)
    S = np.ascontiguousarray(atom.B_λ(line.λ[None, :], atm["temperature"][:, None]))
    ls = line.as_struct()
    nl = min(nlam, max(1, threads))
    l0 = max(0, line.λidx[1] // 2 - nl // 2)

    def run():
        t = time.perf_counter()
        O.J_lambda_voronoi(osites, ls, line.λ, sd, oq, S, lte.T, l0=l0, l1=l0 + nl, hoist=hoist)
        return time.perf_counter() - t
    return run, dict(nlam_sample=nl, l0=l0, ndirs=int(nq), n=P["n"], threads=O.num_threads())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("VRT_WORKLOAD", "nlte_16m"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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
    if args.impl == "reference":
        if rank != 0:
            return 0
        from voronoirt_b200 import synth
        P = build_problem(args.workload)
        atm = P["atm"]
        line, lte, α_cont, ελ, Cr = synth.line_inputs(atm["temperature"], atm["electron_density"], atm["hydrogen_density"], P["nbb"], P["nbf"])
        threads = os.cpu_count() or 1
        run, info = cpu_reference_sample(P, line, (lte, α_cont, ελ, Cr), threads)
        for _ in range(min(W, 1)):
            run()
        ts = [run() for _ in range(K)]
        t = float(np.mean(ts))
        updates = info["n"] * info["ndirs"] * info["nlam_sample"]
        val = updates / t
        sample = (f"{info['nlam_sample']} of {len(line.λ)} wavelengths (index {info['l0']}..) x the first {info['ndirs']} of the quadrature's directions x {info['n']} sites, "
                  "stencil recomputed at every visit like irregular_ray_tracing.jl:50; C/OpenMP port of the Julia reference (Julia not installed)")
        out = {"impl": "reference", "metric": "cell*angle*freq updates/s per formal solution (NLTE Lambda-iteration)", "value": val,
               "unit": "updates/s", "n_gpus": args.gpus, "steps": K, "warmup": min(W, 1), "ms_per_step": 1e3 * t, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(args.workload, P, len(line.λ)), "sites": P["n"], "quadrature": P["qname"], "n_lambda": len(line.λ)},
               "cpu_baseline": {"value": val, "unit": "updates/s", "cores": info["threads"], "kind": "port", "sample": sample},
               "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return 0

    import torch
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib, atom, synth
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().vrt_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t_setup = time.time()
    P = build_problem(args.workload)
    atm, b, n = P["atm"], P["bounds"], P["n"]
    line, lte, α_cont, ελ, Cr = synth.line_inputs(atm["temperature"], atm["electron_density"], atm["hydrogen_density"], P["nbb"], P["nbf"])
    nlam = len(line.λ)
    w, th, ph, nq = V.read_quadrature(P["qpath"])
    # multi-GPU decomposition: D direction shards x G wavelength shards (D*G = world).  Directions first: the sweep's cost
    # is per (cell, direction) visit, so fewer directions per GPU scales it, narrower wavelength rows barely do.
    D, G = shard_grid(world, int(nq))
    di, gi = rank % D, rank // D
    lo, hi = shard_range(nlam, G, gi)
    # directions are dealt round-robin (rank di takes di, di+D, ...): neighbouring lines of the quadrature files are up/down
    # pairs of similar inclination, so every rank gets a similar mix of cheap (steep) and expensive (grazing) directions
    mine = list(range(di, int(nq), D))
    my_quad = (w[mine], th[mine], ph[mine]) if D > 1 else P["qpath"]
    dlo, dhi = 0, len(mine)
    cell = V.read_cell(P["nbr"], n, P["pos"], b["x_min"], b["x_max"], b["y_min"], b["y_max"])
    sites = V.VoronoiSites(*cell, atm["temperature"], atm["electron_density"], atm["hydrogen_density"], atm["velocity_z"],
                           atm["velocity_x"], atm["velocity_y"], b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"], n)
    ndirs = int(np.sum(th != 90))
    solver = V.Solver(sites, my_quad, line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte,
                      lam_range=(lo, hi) if G > 1 else None, dir_range=(dlo, dhi) if D > 1 else None,
                      cell_shard=(di, D) if D > 1 and not os.environ.get("VRT_NO_CELL_SHARD") else None)
    coll = {"ms": 0.0, "bytes": 0}
    if world > 1:
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
                torch.cuda.synchronize()
                coll["ms"] += 1e3 * (time.perf_counter() - t0c)
                coll["bytes"] += 8 * count
                return 0
            if op == 1:
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
    log(f"setup {time.time() - t_setup:.1f}s: n={n} dirs={ndirs} nlam={nlam} shards: {D} direction x {G} wavelength; rank 0 has directions {mine} wavelengths [{lo},{hi}) layers up/down={len(sites.layers_up) - 1}/{len(sites.layers_down) - 1}")

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
    coll_ms, coll_bytes = coll["ms"] / K, coll["bytes"] / K
    tt = torch.tensor([ms, stats["sweep_ms"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max, sweep_ms_max = float(tt[0]), float(tt[1])
    updates_total = float(n) * ndirs * nlam          # all ranks together, per step
    value = updates_total / (ms_max / K / 1e3)

    # ---- roofline of the dominant kernel (k_sweep): algorithmic bytes / measured kernel time (CUDA events in the library)
    B_alg = 40.0 + 104.0 / nlam
    local_updates = float(n) * int(np.sum(th[mine] != 90)) * (hi - lo)
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
    roofline = {"bound": "hbm", "kernel": "k_sweep", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_update": B_alg,
                "sweep_ms_per_step": sweep_ms_max / K, "sweep_share_of_step": sweep_ms_max / ms_max,
                "cell_visits_per_step": stats["visits"] / K * (1.0), "dependent_steps_per_step": stats["steps"] / K}

    # ---- end to end through the C ABI with pinned host buffers: state in, one Λ-iteration, S/J/populations out
    # Per step: S and populations go host -> device (restart state, like recover_simulation.jl), one Λ-iteration runs, and S and
    # populations come back device -> host (what the reference writes to its HDF5 file after every iteration,
    # lambda_iteration.jl:280-281).
    e2e = None
    dt_local = float("inf")
    nl = hi - lo
    try:
        if args.no_e2e:
            raise StopIteration
        import ctypes as C
        L = _lib.lib()
        hS = torch.empty((n, nl), dtype=torch.float64).pin_memory()
        hP = torch.empty((3, n), dtype=torch.float64).pin_memory()
        _lib.check(L.vrt_get_state(solver.h, C.c_void_p(hS.data_ptr()), None, C.c_void_p(hP.data_ptr())))

        null_cb = _abi_null_cb()
        dbg = os.environ.get("VRT_DEBUG")

        def step():
            ta = time.perf_counter()
            _lib.check(L.vrt_set_state(solver.h, C.c_void_p(hS.data_ptr()), C.c_void_p(hP.data_ptr())))
            tb = time.perf_counter()
            _lib.check(L.vrt_lambda_iterate(solver.h, -1.0, 1, null_cb, None, None))
            tc = time.perf_counter()
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
            e2e = {"value": updates_total / (dt / K), "unit": "updates/s", "h2d_bytes_per_step": int(8 * (n * nl + 3 * n)),
                   "d2h_bytes_per_step": int(8 * (n * nl + 3 * n)), "ms_per_step": 1e3 * dt / K,
                   "api": "vrt_set_state(S, populations) + vrt_lambda_iterate(1 iteration) + vrt_get_state(S, populations) with pinned host buffers; best of 2 repetitions of K steps"}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        run, info = cpu_reference_sample(P, line, (lte, α_cont, ελ, Cr), threads)
        t = run()
        cpu = {"value": info["n"] * info["ndirs"] * info["nlam_sample"] / t, "unit": "updates/s", "cores": info["threads"], "kind": "port",
               "sample": f"{info['nlam_sample']} of {nlam} wavelengths x the first {info['ndirs']} of {int(nq)} directions x {info['n']} sites, one formal solution "
                         f"({t:.1f} s), stencil recomputed per visit like the reference; C/OpenMP port (Julia not installed)"}
        if n <= 4_000_000:   # second flavour (stencil hoisted out of the wavelength loop: the fairer CPU bar); skipped on huge grids
            run_h, _ = cpu_reference_sample(P, line, (lte, α_cont, ελ, Cr), threads, hoist=1)
            th_ = run_h()
            cpu["hoisted_value"] = info["n"] * info["ndirs"] * info["nlam_sample"] / th_

    if rank == 0:
        out = {"metric": "cell*angle*freq updates/s per formal solution (NLTE Lambda-iteration)", "value": value, "unit": "updates/s",
               "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(args.workload, P, nlam), "sites": n, "quadrature": P["qname"], "n_dirs": ndirs, "n_lambda": nlam,
                          "parallelism": f"{D} direction shards x {G} wavelength shards, NCCL reduce-scatter of J + all-gather of S ({8 * n * (hi - lo) / 1e6:.0f} MB each) per iteration; source update, rates and statistical equilibrium sharded over cells" if world > 1 else "single GPU", "l2_policy": "inputs larger than L2 "
                          f"(S+J+I+alpha = {8 * n * nlam * (2 + 2 * ndirs) / 1e9:.1f} GB)", "n_sweeps": 3, "p": 7.0},
               "s_per_lambda_iteration": ms_max / K / 1e3, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
               "gpu_launches": int(launches), "clocks": clocks,
               "collectives": {"ms_per_step": coll_ms, "bytes_per_step": coll_bytes, "what": "NCCL reduce-scatter of J + all-gather of S and populations over the direction shards, max of the criterion" if world > 1 else None},
               "stage_ms": {k: float(np.mean([h[k] for h in hist])) for k in ("t_opacity_ms", "t_sweep_ms", "t_source_ms", "t_rates_ms", "t_stateq_ms", "t_total_ms")} if hist else None}
        print(json.dumps(out))
    solver.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


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
    """BASELINE configs[3]: NLTE line Λ-iteration on the regular grid (Λ_regular, lambda_iteration.jl:116-205), one GPU."""
    rank = int(os.environ.get("RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if rank == 0:
            print(json.dumps({"workload": args.workload, "unavailable": "the regular-grid path is single-GPU this round (replicas only)"}))
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
    torch.cuda.set_device(0)
    t_setup = time.time()
    Q = regular_problem(*shape, nbb, nbf)
    line, n, f = Q["line"], Q["n"], Q["fields"]
    nlam = len(line.λ)
    atm = V.Atmosphere(Q["z"], Q["x"], Q["y"], f["temperature"], f["electron_density"], f["hydrogen_density"], f["velocity_z"],
                       f["velocity_x"], f["velocity_y"])
    lam_chunk = int(os.environ.get("VRT_REG_BENCH_CHUNK", "0"))   # 0: as many wavelengths per pass as the free HBM allows
    solver = V.Solver(atm, qpath, line=line, α_cont=Q["α_cont"], ελ=Q["ελ"], C_rates=Q["C"], LTE_pops=Q["lte"], lam_chunk=lam_chunk)
    log(f"setup {time.time() - t_setup:.1f}s: regular grid {shape} = {n} cells, dirs={ndirs} nlam={nlam}")
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
    hist = res["history"]
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
        tile = " (sampled directly, tessellated on the GPU by vrt_voronoi_neighbours)"
    return f"{key}: NLTE line Lambda-iteration, {P['n']} Voronoi sites{tile}, {P['qname']}, {nlam} wavelengths, synthetic Bifrost-shaped atmosphere"


def _abi_null_cb():
    from voronoirt_b200 import _abi
    import ctypes as C
    return C.cast(None, _abi.vrt_iter_cb)


if __name__ == "__main__":
    sys.exit(main())
