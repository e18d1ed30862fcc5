"""Round-2 preparation, CPU only: how much of the sweep's DRAM traffic is set by the ORDER of the visits?
Builds the dependency levels of one direction on a stratified grid (oracle stencil, one visit per cell: no re-sweeps), replays
the row accesses of k_sweep (own S, alpha; I, S, alpha of the two upwind cells; write of I) through an LRU cache whose
capacity is the B200's 126 MB L2 scaled by the ratio of the grid sizes (to mimic 16 M sites), and counts misses for
  A  level order, cells by rank inside a level (what schedule.cu emits today),
  B  level order, cells by Morton code of (x, y, z) inside a level,
  C  layer order (the reference's own order: no level parallelism at all),
  D  like B with K directions interleaved level by level (K = 2, the final plan at 16 M sites).
Usage: python profiles/microbench/l2_order_sim.py [n_sites]"""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
from voronoirt_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pos = synth.sample_sites(n, seed=3)
nbr = np.asarray(synth.voronoi_neighbours(pos))
B = synth.BOX
b = np.array([B["z_min"], B["z_max"], B["x_min"], B["x_max"], B["y_min"], B["y_max"]])
sites = O.Sites(np.ascontiguousarray(pos.T), np.ascontiguousarray(nbr.T), b)


def morton(p):
    q = ((p - p.min(axis=1, keepdims=True)) / (np.ptp(p, axis=1)[:, None] + 1e-300) * 1023).astype(np.uint64)
    code = np.zeros(p.shape[1], dtype=np.uint64)
    for bit in range(10):
        for a in range(3):
            code |= ((q[a] >> np.uint64(bit)) & np.uint64(1)) << np.uint64(3 * bit + a)
    return code


def program(theta, phi):
    t, p = theta * np.pi / 180, phi * np.pi / 180
    k = np.array([np.cos(t), np.cos(p) * np.sin(t), np.sin(p) * np.sin(t)])
    down = int(theta < 90)
    perm, off = sites.layers(down)
    rank = np.empty(n, dtype=np.int64)
    rank[perm - 1] = np.arange(n)
    up, dots, w, r = sites.stencil(k)
    u = up - 1                                            # (n, 2) 0-based upwind cells
    n1 = off[1] - 1
    level = np.zeros(n, dtype=np.int64)
    order_ref = perm - 1
    solved = np.zeros(n, dtype=bool)
    solved[order_ref[:n1]] = True                          # boundary layer
    for c in order_ref[n1:n - 1]:                          # reference order; the last-rank site is never solved (Q1)
        lv = 0
        for m in (0, 1):
            if rank[u[c, m]] < rank[c]:                    # FINAL or THIS: a real dependency
                lv = max(lv, level[u[c, m]])
        level[c] = lv + 1
    cells = order_ref[n1:n - 1]
    return dict(cells=cells, level=level, rank=rank, u=u)


def trace(cells, P, d):
    """row ids touched by the visits, in order: (array, direction, cell) packed into one integer"""
    u = P["u"]
    S, A, I = 0, 1 + 2 * d, 2 + 2 * d                      # S is shared by the directions
    out = np.empty((len(cells), 9), dtype=np.int64)
    for col, (arr, who) in enumerate(((S, cells), (A, cells), (I, u[cells, 0]), (S, u[cells, 0]), (A, u[cells, 0]),
                                      (I, u[cells, 1]), (S, u[cells, 1]), (A, u[cells, 1]), (I, cells))):
        out[:, col] = arr * n + who
    return out


def misses(rows, capacity):
    lru = OrderedDict()
    miss = 0
    for r in rows.ravel():
        if r in lru:
            lru.move_to_end(r)
        else:
            miss += 1
            lru[r] = None
            if len(lru) > capacity:
                lru.popitem(last=False)
    return miss


cap = int(126e6 / 728 * n / 16e6)                          # L2 rows, scaled to this grid
code = morton(pos)
dirs = [(152.7, 315.5), (67.2, 155.8)]
progs = [program(*d) for d in dirs]
P = progs[0]
c = P["cells"]
algorithmic = 4 * len(c)                                   # per visit: S, alpha read once, I written once and read once downstream
res = {}
oA = c[np.lexsort((P["rank"][c], P["level"][c]))]
oB = c[np.lexsort((code[c], P["level"][c]))]
oC = c[np.argsort(P["rank"][c])]
for name, o in (("A level, rank", oA), ("B level, Morton", oB), ("C reference order (layer, rank)", oC)):
    res[name] = misses(trace(o, P, 0), cap)
# two directions interleaved level by level
def interleave(order_fn):
    per = [order_fn(Q) for Q in progs]
    lv = [Q["level"][o] for Q, o in zip(progs, per)]
    rows = []
    for L in range(1, max(int(l.max()) for l in lv) + 1):
        for d, (Q, o, l) in enumerate(zip(progs, per, lv)):
            sel = o[l == L]
            if len(sel):
                rows.append(trace(sel, Q, d))
    return np.concatenate(rows)
res["A x2 directions interleaved"] = misses(interleave(lambda Q: Q["cells"][np.lexsort((Q["rank"][Q["cells"]], Q["level"][Q["cells"]]))]), cap) / 2
res["B x2 directions interleaved"] = misses(interleave(lambda Q: Q["cells"][np.lexsort((code[Q["cells"]], Q["level"][Q["cells"]]))]), cap) / 2
print(f"n = {n}, levels = {int(P['level'].max())}, L2 capacity scaled = {cap} rows, visits = {len(c)}")
for k_, v in res.items():
    print(f"{k_:38s} misses per visit {v / len(c):5.2f}   = {v / algorithmic:4.2f} x algorithmic")

# E: idealised spatial blocking (NOT a valid schedule as it stands: it ignores dependencies that cross a tile boundary
# against the tile order) — an upper bound on what a tiled wavefront schedule could save
t, p = dirs[0][0] * np.pi / 180, dirs[0][1] * np.pi / 180
kvec = np.array([np.cos(t), np.cos(p) * np.sin(t), np.sin(p) * np.sin(t)])
for cells_per_tile in (64, 256, 1024):
    ntile = max(1, int(round((n / cells_per_tile) ** (1 / 3))))
    q = ((pos - pos.min(axis=1, keepdims=True)) / (np.ptp(pos, axis=1)[:, None] + 1e-300) * ntile).astype(np.int64).clip(0, ntile - 1)
    tile = q[0] + ntile * (q[1] + ntile * q[2])
    centre = (q + 0.5) / ntile * np.ptp(pos, axis=1)[:, None]
    proj = -(kvec[:, None] * centre).sum(axis=0)            # upwind tiles first (k points from the cell to its upwind side)
    key = np.round(proj / np.ptp(proj) * 1e6).astype(np.int64) * (ntile ** 3) + tile
    oE = c[np.lexsort((P["level"][c], key[c]))]
    posn = np.empty(n, dtype=np.int64); posn[:] = -1
    posn[oE] = np.arange(len(oE))
    viol = 0
    for m in (0, 1):
        uu = P["u"][oE, m]
        dep = (P["rank"][uu] < P["rank"][oE]) & (posn[uu] >= 0)
        viol += int((dep & (posn[uu] > posn[oE])).sum())
    mE = misses(trace(oE, P, 0), cap)
    print(f"E tiles of ~{cells_per_tile:4d} cells, upwind tiles first   misses per visit {mE / len(c):5.2f}   = {mE / algorithmic:4.2f} x algorithmic   "
          f"({viol} of {2 * len(c)} dependencies point against the order)")

# F: the valid version of E: key = max(own (tile, level) key, key of the producers) — a topological order by construction
for cells_per_tile in (256,):
    ntile = max(1, int(round((n / cells_per_tile) ** (1 / 3))))
    q = ((pos - pos.min(axis=1, keepdims=True)) / (np.ptp(pos, axis=1)[:, None] + 1e-300) * ntile).astype(np.int64).clip(0, ntile - 1)
    tile = q[0] + ntile * (q[1] + ntile * q[2])
    centre = (q + 0.5) / ntile * np.ptp(pos, axis=1)[:, None]
    proj = -(kvec[:, None] * centre).sum(axis=0)
    tkey = np.round(proj / np.ptp(proj) * 1e6).astype(np.int64) * (ntile ** 3) + tile
    uniq, tord = np.unique(tkey, return_inverse=True)          # position of the cell's tile in the tile order
    own = tord.astype(np.float64) * 1e3 + P["level"]            # (tile, level)
    key = own.copy()
    rank = P["rank"]
    for cc in c[np.argsort(rank[c])]:                           # reference order: producers come first
        for m in (0, 1):
            uu = P["u"][cc, m]
            if rank[uu] < rank[cc] and key[uu] >= key[cc]:
                key[cc] = np.nextafter(key[uu], np.inf)
    oF = c[np.lexsort((rank[c], key[c]))]
    posn = np.full(n, -1, dtype=np.int64)
    posn[oF] = np.arange(len(oF))
    viol = sum(int((((rank[P["u"][oF, m]] < rank[oF]) & (posn[P["u"][oF, m]] >= 0)) & (posn[P["u"][oF, m]] > posn[oF])).sum()) for m in (0, 1))
    mF = misses(trace(oF, P, 0), cap)
    moved = int((key[c] != own[c]).sum())
    print(f"F valid tile order (~{cells_per_tile} cells per tile)          misses per visit {mF / len(c):5.2f}   = {mF / algorithmic:4.2f} x algorithmic   "
          f"({viol} violations, {moved} cells pushed behind a producer)")
