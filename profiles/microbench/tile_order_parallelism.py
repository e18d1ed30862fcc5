"""Round-2 preparation, CPU only: does the cache-friendly tile order (l2_order_sim.py, variant F) leave enough parallelism
for the dataflow kernel?  Discrete-event model of k_sweep's claim discipline: W workers (warps) claim visits strictly in
program order; a claimed visit waits until its producers have finished, then takes one time unit.  Reports the makespan
relative to the ideal N / W for the level order and the tile order, with K directions interleaved (chunks of 32 visits of
each direction in turn).  W is the B200's ~3100 consumer warps scaled by the grid ratio to 16 M sites."""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "profiles", "microbench"))
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["100000"])
import io
import contextlib
with contextlib.redirect_stdout(io.StringIO()):
    import l2_order_sim as L                                  # builds the grid, the two programs and the helper functions

n = L.n
W = max(4, int(round(3100 * n / 16e6)))


def orders(kind):
    out = []
    for d, Q in enumerate(L.progs):
        c = Q["cells"]
        if kind == "level":
            o = c[np.lexsort((Q["rank"][c], Q["level"][c]))]
        else:
            t, p = L.dirs[d][0] * np.pi / 180, L.dirs[d][1] * np.pi / 180
            kv = np.array([np.cos(t), np.cos(p) * np.sin(t), np.sin(p) * np.sin(t)])
            ntile = max(1, int(round((n / 256) ** (1 / 3))))
            q = ((L.pos - L.pos.min(axis=1, keepdims=True)) / (np.ptp(L.pos, axis=1)[:, None] + 1e-300) * ntile).astype(np.int64).clip(0, ntile - 1)
            tile = q[0] + ntile * (q[1] + ntile * q[2])
            centre = (q + 0.5) / ntile * np.ptp(L.pos, axis=1)[:, None]
            proj = -(kv[:, None] * centre).sum(axis=0)
            tkey = np.round(proj / np.ptp(proj) * 1e6).astype(np.int64) * (ntile ** 3) + tile
            _, tord = np.unique(tkey, return_inverse=True)
            key = tord.astype(np.float64) * 1e3 + Q["level"]
            rank = Q["rank"]
            for cc in c[np.argsort(rank[c])]:
                for m in (0, 1):
                    uu = Q["u"][cc, m]
                    if rank[uu] < rank[cc] and key[uu] >= key[cc]:
                        key[cc] = np.nextafter(key[uu], np.inf)
            o = c[np.lexsort((rank[c], key[c]))]
        out.append(o)
    return out


def makespan(per_dir, K):
    """interleave the first K directions in chunks of 32 visits; simulate"""
    seq = []
    ptr = [0] * K
    while any(ptr[d] < len(per_dir[d]) for d in range(K)):
        for d in range(K):
            o = per_dir[d]
            for cc in o[ptr[d]:ptr[d] + 32]:
                seq.append((d, cc))
            ptr[d] += 32
    finish = [dict() for _ in range(K)]
    workers = [0.0] * W
    heapq.heapify(workers)
    for d, cc in seq:
        Q = L.progs[d]
        free = heapq.heappop(workers)
        ready = free
        for m in (0, 1):
            uu = Q["u"][cc, m]
            if Q["rank"][uu] < Q["rank"][cc]:
                ready = max(ready, finish[d].get(uu, 0.0))
        end = ready + 1.0
        finish[d][cc] = end
        heapq.heappush(workers, end)
    return max(workers) / (len(seq) / W)


print(f"n = {n}, workers = {W}")
for kind in ("level", "tile"):
    per = orders(kind)
    for K in (1, 2):
        print(f"{kind:5s} order, {K} direction(s) in flight: makespan / ideal = {makespan(per, K):.2f}")
