"""Writes tests/golden/tiny_voro.npz: voro++'s neighbour lists (the reference's rt_preprocessing/output_sites driver) for
boxes with 1..40 sites, where cells are bounded by periodic images of other sites and of themselves.  Run once in a
container that has the reference mounted:  python tests/make_tiny_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voronoirt_b200 import synth  # noqa: E402

bounds = dict(z_min=0.0, z_max=1.0, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0)
out = {}
case = 0
for n in (1, 2, 3, 5, 8, 15, 40):
    for seed in range(3):
        rng = np.random.default_rng(100 * n + seed)
        pos = np.asfortranarray(rng.random((3, n)))
        nbr = np.asarray(synth.voronoi_neighbours(pos, bounds=bounds))
        out[f"pos_{case}"] = pos
        out[f"nbr_{case}"] = nbr.astype(np.int32)
        case += 1
out["n_cases"] = np.array(case)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tiny_voro.npz"), **out)
print("wrote", case, "cases")
