"""Round-2 preparation, CPU only: traffic against parallelism for visit orders between "all levels across the grid"
(today) and "one small tile at a time".  Order key = (group of g consecutive tiles in upwind order, level inside the
group), made valid by pushing cells behind their producers.  For each (tile size, g): LRU misses per visit relative to the
algorithmic 4 (l2_order_sim.py) and makespan / ideal of the claim-in-order model (tile_order_parallelism.py), two
directions in flight."""
import heapq
import io
import contextlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "profiles", "microbench"))
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["100000"])
with contextlib.redirect_stdout(io.StringIO()):
    import l2_order_sim as L

n = L.n
W = max(4, int(round(3100 * n / 16e6)))


def order_for(d, cells_per_tile, g):
    Q = L.progs[d]
    c = Q["cells"]
    if cells_per_tile is None:
        return c[np.lexsort((Q["rank"][c], Q["level"][c]))]
    t, p = L.dirs[d][0] * np.pi / 180, L.dirs[d][1] * np.pi / 180
    kv = np.array([np.cos(t), np.cos(p) * np.sin(t), np.sin(p) * np.sin(t)])
    ntile = max(1, int(round((n / cells_per_tile) ** (1 / 3))))
    q = ((L.pos - L.pos.min(axis=1, keepdims=True)) / (np.ptp(L.pos, axis=1)[:, None] + 1e-300) * ntile).astype(np.int64).clip(0, ntile - 1)
    tile = q[0] + ntile * (q[1] + ntile * q[2])
    centre = (q + 0.5) / ntile * np.ptp(L.pos, axis=1)[:, None]
    proj = -(kv[:, None] * centre).sum(axis=0)
    tkey = np.round(proj / np.ptp(proj) * 1e6).astype(np.int64) * (ntile ** 3) + tile
    _, tord = np.unique(tkey, return_inverse=True)
    key = (tord // g).astype(np.float64) * 1e3 + Q["level"]
    rank = Q["rank"]
    for cc in c[np.argsort(rank[c])]:
        for m in (0, 1):
            uu = Q["u"][cc, m]
            if rank[uu] < rank[cc] and key[uu] >= key[cc]:
                key[cc] = np.nextafter(key[uu], np.inf)
    return c[np.lexsort((rank[c], key[c]))]


def makespan(per_dir):
    K = len(per_dir)
    seq, ptr = [], [0] * K
    while any(ptr[d] < len(per_dir[d]) for d in range(K)):
        for d in range(K):
            seq.extend((d, cc) for cc in per_dir[d][ptr[d]:ptr[d] + 32])
            ptr[d] += 32
    finish = [dict() for _ in range(K)]
    workers = [0.0] * W
    for d, cc in seq:
        Q = L.progs[d]
        ready = heapq.heappop(workers)
        for m in (0, 1):
            uu = Q["u"][cc, m]
            if Q["rank"][uu] < Q["rank"][cc]:
                ready = max(ready, finish[d].get(uu, 0.0))
        finish[d][cc] = ready + 1.0
        heapq.heappush(workers, ready + 1.0)
    return max(workers) / (len(seq) / W)


print(f"n = {n}, workers = {W}, cache = {L.cap} rows;  traffic = LRU misses / algorithmic (one direction);  makespan / ideal (two directions in flight)")
print(f"{'order':34s} traffic  makespan")
for cpt, g in ((None, 1), (256, 1), (256, 4), (256, 16), (256, 64), (64, 16), (1024, 4)):
    per = [order_for(d, cpt, g) for d in range(2)]
    tr = L.misses(L.trace(per[0], L.progs[0], 0), L.cap) / L.algorithmic
    ms = makespan(per)
    name = "levels over the whole grid (today)" if cpt is None else f"tiles of ~{cpt} cells, groups of {g}"
    print(f"{name:34s} {tr:6.2f}x  {ms:6.2f}x")
