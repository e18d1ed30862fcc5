"""Order probe (GPU): times the Λ-iteration of one workload under several visit orders of the sweep program
(schedule.cu rule 5: column blocks x level slabs) and checks that J does not change by a bit.

    python profiles/microbench/order_probe.py --workload nlte_16m_native --configs "1,1,0,2;4,4,0,2;8,8,150,2" [--dirs 0:4]

A config is  bx,by,slab,max_dirs[,NAME=VALUE...]  (extra NAME=VALUE pairs are exported for that config only).
Prints one JSON line per config: sweep / opacity / total ms per iteration, steps per launch and a bitwise checksum of J.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="nlte_4m_native")
    ap.add_argument("--configs", default="1,1,0,2;4,4,0,2")
    ap.add_argument("--dirs", default="")
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import bench
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib, synth
    import ctypes as C
    torch.cuda.set_device(0)
    t0 = time.time()
    P = bench.build_problem(args.workload)
    atm, b, n = P["atm"], P["bounds"], P["n"]
    line, lte, α_cont, ελ, Cr = synth.line_inputs(atm["temperature"], atm["electron_density"], atm["hydrogen_density"], P["nbb"], P["nbf"])
    nlam = len(line.λ)
    w, th, ph, nq = V.read_quadrature(P["qpath"])
    cell = V.read_cell(P["nbr"], n, P["pos"], b["x_min"], b["x_max"], b["y_min"], b["y_max"])
    sites = V.VoronoiSites(*cell, atm["temperature"], atm["electron_density"], atm["hydrogen_density"], atm["velocity_z"],
                           atm["velocity_x"], atm["velocity_y"], b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"], n)
    print(f"[probe] set-up {time.time() - t0:.1f}s: n={n} layers {len(sites.layers_up) - 1}/{len(sites.layers_down) - 1}", file=sys.stderr, flush=True)
    dr = None
    nd = int(np.sum(th != 90))
    if args.dirs:
        a, e = (int(v) for v in args.dirs.split(":"))
        dr = (a, e)
        nd = int(np.sum(th[a:e] != 90))
    Jdev = torch.empty((n, nlam), dtype=torch.float64, device="cuda")
    results = []
    for cfg in args.configs.split(";"):
        parts = cfg.split(",")
        bx, by, slab, md = (int(v) for v in parts[:4])
        extra = dict(p.split("=", 1) for p in parts[4:])
        env = {"VRT_BLOCKS": f"{bx},{by}", "VRT_SLAB": str(slab), "VRT_MAX_DIRS": str(md), **extra}
        for k, v in env.items():
            os.environ[k] = v
        rec = {"config": cfg, "workload": args.workload, "n": n, "dirs": nd}
        try:
            t1 = time.time()
            solver = V.Solver(sites, P["qpath"], line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte, dir_range=dr)
            rec["setup_s"] = time.time() - t1
            solver.iterate(-1.0, 1)
            torch.cuda.synchronize()
            res = solver.iterate(-1.0, args.iters)
            torch.cuda.synchronize()
            st = _lib.last_stats()
            h = res["history"]
            rec.update(sweep_ms=float(np.mean([x["t_sweep_ms"] for x in h])), opacity_ms=float(np.mean([x["t_opacity_ms"] for x in h])),
                       total_ms=float(np.mean([x["t_total_ms"] for x in h])), launches=st["kernels"] / args.iters,
                       steps_per_iter=st["steps"] / args.iters, visits_per_iter=st["visits"] / args.iters)
            _lib.check(_lib.lib().vrt_get_state(solver.h, None, C.c_void_p(Jdev.data_ptr()), None))
            torch.cuda.synchronize()
            rec["J_checksum"] = int(Jdev.view(torch.int64).sum().item())
            rec["J_max"] = float(Jdev.max().item())
            upd = float(n) * nd * nlam
            rec["roofline_frac"] = (40.0 + 104.0 / nlam) * upd / (rec["sweep_ms"] / 1e3) / 1e9 / bench.measured_peak()[0]
            solver.close()
        except Exception as ex:  # noqa: BLE001
            rec["error"] = str(ex)
        _lib.check(_lib.lib().vrt_grid_release_schedules(sites._grid.h))
        for k in env:
            os.environ.pop(k, None)
        print(json.dumps(rec), flush=True)
        results.append(rec)
    ok = len({r.get("J_checksum") for r in results if "error" not in r}) <= 1
    print(json.dumps({"J_bitwise_identical_across_orders": ok}), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for r in results:
                f.write(json.dumps(r) + "\n")
            f.write(json.dumps({"J_bitwise_identical_across_orders": ok}) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
