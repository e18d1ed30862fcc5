"""BASELINE.json configs[1]: continuum Λ-iteration (single wavelength, ul7n12) on the 1 M-site grid — timing probe.
usage (on a B200): python profiles/microbench/continuum_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import voronoirt_b200 as V
from voronoirt_b200 import _lib, synth

P = bench.build_problem("nlte_1m")
atm, b, n = P["atm"], P["bounds"], P["n"]
cell = V.read_cell(P["nbr"], n, P["pos"], b["x_min"], b["x_max"], b["y_min"], b["y_max"])
sites = V.VoronoiSites(*cell, atm["temperature"], atm["electron_density"], atm["hydrogen_density"], atm["velocity_z"], atm["velocity_x"],
                       atm["velocity_y"], b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"], n)
α, ε, B0 = synth.continuum_inputs(atm["temperature"], atm["electron_density"], atm["hydrogen_density"])
s = V.Solver(sites, P["qpath"], α_cont=α, ελ=ε, B_0=B0)
s.iterate(-1.0, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = s.iterate(-1.0, 10); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
st = _lib.last_stats()
upd = n * 12
print(f"continuum 1M sites ul7n12: {ms:.3f} ms per Λ-iteration, {upd / ms * 1e3:.3e} updates/s, sweep {st['sweep_ms'] / 10:.3f} ms, "
      f"frac of HBM roofline (144 B/update) {144 * upd / (st['sweep_ms'] / 10 / 1e3) / 6554.2e9:.3f}")
res = s.iterate(1e-3, 150)
print("converged", res["converged"], "in", res["iterations"], "more iterations; diff", res["diff"])
