"""Generates the committed fixtures under tests/golden/ (run here, where /root/reference exists):

  grid_unit1000.npz   the grid of SURVEY.md App. F: default_rng(0).random((1000,3)) as (x,y,z), unit box
  grid_strat3000.npz  3000 sites of the synthetic Bifrost-shaped box (voronoirt_b200.synth, seed 7)
  grid_unit300.npz    300 uniform sites (tiny; edge cases)

Neighbour lists come from the reference's own prebuilt voro++ driver rt_preprocessing/output_sites.
Usage: python tests/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voronoirt_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
UNIT = dict(z_min=0.0, z_max=1.0, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0)


def save(name, pos, nbr, b):
    bounds = np.array([b["z_min"], b["z_max"], b["x_min"], b["x_max"], b["y_min"], b["y_max"]])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), positions=pos, neighbours=nbr.astype(np.int32), bounds=bounds)
    print(name, pos.shape, nbr.shape)


def main():
    p = np.random.default_rng(0).random((1000, 3))  # columns x, y, z
    pos = np.asfortranarray(np.stack([p[:, 2], p[:, 0], p[:, 1]]))
    save("grid_unit1000", pos, synth.voronoi_neighbours(pos, bounds=UNIT), UNIT)
    p = np.random.default_rng(5).random((300, 3))
    pos = np.asfortranarray(np.stack([p[:, 2], p[:, 0], p[:, 1]]))
    save("grid_unit300", pos, synth.voronoi_neighbours(pos, bounds=UNIT), UNIT)
    pos = synth.sample_sites(3000, seed=7)
    save("grid_strat3000", pos, synth.voronoi_neighbours(pos), synth.BOX)


if __name__ == "__main__":
    main()
