"""Extracts the only usable statistics of the reference's Voronoi searchlight rasters (the sites were
unseeded, so only the beam centroid is comparable) into tests/golden/searchlight_stats.json.
Run where /root/reference exists:  python tests/make_searchlight_golden.py"""
import json
import os

import numpy as np

REF = "/root/reference/data/searchlight_data"
out = {}
x = np.load(f"{REF}/x_voronoi.npy")
y = np.load(f"{REF}/y_voronoi.npy")
for name in ("I_160_45_voronoi", "I_20_15_voronoi"):
    a = np.load(f"{REF}/{name}.npy")
    out[name] = {"shape": list(a.shape), "mean": float(a.mean()), "max": float(a.max()),
                 "centroid_x": float((a.sum(axis=1) * x).sum() / a.sum()), "centroid_y": float((a.sum(axis=0) * y).sum() / a.sum())}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "searchlight_stats.json"), "w"), indent=1)
print(out)
