"""Native Voronoi neighbour generation (vrt_voronoi_neighbours) on one B200 against the reference's voro++
driver on the host: set equality and time at 250 k sites, GPU time at larger counts (voro++ is single-threaded with fixed
6x6x6 blocks and is not run there).  Usage: python profiles/microbench/tessellate_probe.py [n_large ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import voronoirt_b200 as V  # noqa: E402
from voronoirt_b200 import _lib, synth  # noqa: E402

B = synth.BOX
b = [B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"]]
rows = []


def gpu(pos):
    t = time.perf_counter()
    nbr = V.voronoi_neighbours(pos, *b)
    wall = time.perf_counter() - t
    return nbr, wall, _lib.last_stats()["sweep_ms"]


pos = synth.sample_sites(250000, seed=2022)
gpu(pos[:, :1000])                                   # context / module load
nbr, wall, ms = gpu(pos)
t = time.perf_counter()
ref = np.asarray(synth.voronoi_neighbours(pos))
t_voro = time.perf_counter() - t
same = np.array_equal(nbr[:, 0], ref[:, 0])
if same:
    a = np.sort(np.where(np.arange(nbr.shape[1])[None, :] <= nbr[:, :1], nbr, np.iinfo(np.int64).max)[:, 1:], axis=1)
    r = np.sort(np.where(np.arange(ref.shape[1])[None, :] <= ref[:, :1], ref, np.iinfo(np.int64).max)[:, 1:], axis=1)
    w = min(a.shape[1], r.shape[1])
    same = bool(np.array_equal(a[:, :w], r[:, :w]))
rows.append({"n": 250000, "same_sets_as_voro": bool(same), "gpu_kernels_ms": ms, "gpu_wall_s": wall, "voro_write_run_parse_s": t_voro,
             "max_faces": int(nbr[:, 0].max()), "mean_faces": float(nbr[:, 0].mean())})
print(json.dumps(rows[-1]), flush=True)
for n in [int(v) for v in sys.argv[1:]] or [1000000, 4000000]:
    pos = synth.sample_sites(n, seed=11)
    nbr, wall, ms = gpu(pos)
    rows.append({"n": n, "gpu_kernels_ms": ms, "gpu_wall_s": wall, "max_faces": int(nbr[:, 0].max()), "mean_faces": float(nbr[:, 0].mean()),
                 "walls": [int((nbr == -5).sum()), int((nbr == -6).sum())]})
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tessellate_probe.json"), "w"), indent=1)
