"""Round-2 preparation, CPU only: an executable specification of the tile-ordered sweep program.

The reference sweep (irregular_ray_tracing.jl:37-80) is sequential; libvrt turns it into visits (cell, sweep) whose operands
are FINAL / THIS / LAG / ZERO references and runs them in ANY topological order.  This script builds the
visits of one direction on tests/golden/grid_strat3000.npz, orders them by the cache-friendly key of l2_order_sim.py
(tile in upwind order, then DAG level, pushed behind the producers), executes them in that order with one storage slot
per (cell, sweep), and checks the result against the sequential oracle bit for bit.  It is the order schedule.cu should
emit in round 2; the kernel is unchanged by it."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from conftest import load_grid  # noqa: E402
from pyref import linear_weights  # noqa: E402

pos, nbr, b = load_grid("grid_strat3000")
n = pos.shape[1]
sites = O.Sites(np.ascontiguousarray(pos.T), np.ascontiguousarray(nbr.T), b)
rng = np.random.default_rng(0)
S = rng.uniform(0.2, 2.0, n)
alpha = 10.0 ** rng.uniform(-8, -4, n)
NS = 3


def run(theta, phi, cells_per_tile=64, blocks=None, slab=0):
    """blocks = (bx, by), slab = levels per slab: the key of schedule.cu rule 5 (round 2) instead of the tile key:
    key = (slab of the visit's level, rank of its column block in upwind order, level), pushed behind the producers by
    key(v) = max(base(v), key(p) + 1) on the level field — equal keys are then mutually independent visits."""
    t, p = theta * np.pi / 180, phi * np.pi / 180
    k = np.array([np.cos(t), np.cos(p) * np.sin(t), np.sin(p) * np.sin(t)])
    down = int(theta < 90)
    perm, off = sites.layers(down)
    n1 = off[1] - 1
    I0 = rng.uniform(0.0, 1.0, n1)
    ref = sites.formal_solve(k, down, S[:, None], alpha[:, None], I0[:, None])[:, 0]
    up, dots, w, r = sites.stencil(k)
    u = up - 1
    layer = np.zeros(n, dtype=np.int64)
    for L in range(1, len(off)):
        layer[perm[off[L - 1] - 1: (off[L] - 1 if L < len(off) - 1 else n)] - 1] = L
    # processing position inside the reference loop: ascending rank for up, descending inside a layer for down (:41 / :122)
    proc = np.full(n, -1, dtype=np.int64)
    q = 0
    processed = []
    for L in range(2, len(off)):
        lo, hi = off[L - 1], off[L]                      # 1-based [lo, hi): the loop never reaches the last rank (Q1)
        ranks = range(hi - 1, lo - 1, -1) if down else range(lo, hi)
        for i in ranks:
            c = perm[i - 1] - 1
            proc[c] = q
            q += 1
            processed.append(c)
    processed = np.array(processed)
    boundary = perm[:n1] - 1
    final0 = np.zeros(n)
    final0[boundary] = I0                                # cells that are never processed keep I_0 (layer 1) or zero (Q1)

    # operands of visit (c, s): kind FINAL (value of sweep NS or the fixed value), THIS (c', s), LAG (c', s-1), ZERO
    def operand(c, m, s):
        v = u[c, m]
        if proc[v] < 0:
            return ("fixed", v)
        if layer[v] < layer[c]:
            return ("visit", v, NS)
        if layer[v] > layer[c]:
            return ("zero",)
        if proc[v] < proc[c]:
            return ("visit", v, s)
        return ("visit", v, s - 1) if s > 1 else ("zero",)

    # DAG level of every visit, in reference order (producers first)
    level = {}
    order_ref = [(c, s) for L in range(2, len(off)) for s in range(1, NS + 1) for c in processed[layer[processed] == L]]
    deps = {}
    for (c, s) in order_ref:
        d = [op[1:] for op in (operand(c, 0, s), operand(c, 1, s)) if op[0] == "visit"]
        deps[(c, s)] = d
        level[(c, s)] = 1 + max([level[x] for x in d], default=0)
    if blocks is not None:
        bx, by = blocks
        ix = np.clip(((pos[1] - b[2]) / (b[3] - b[2]) * bx).astype(np.int64), 0, bx - 1)
        iy = np.clip(((pos[2] - b[4]) / (b[5] - b[4]) * by).astype(np.int64), 0, by - 1)
        cx = b[2] + (np.arange(bx) + 0.5) * (b[3] - b[2]) / bx
        cy = b[4] + (np.arange(by) + 0.5) * (b[5] - b[4]) / by
        proj = -(k[1] * cx[:, None] + k[2] * cy[None, :]).ravel()
        rank = np.empty(bx * by, dtype=np.int64)
        rank[np.argsort(proj, kind="stable")] = np.arange(bx * by)
        brank = rank[ix * by + iy]
        key = {}
        pushed = 0
        for v in order_ref:                      # producers first: one pass reaches the least fixed point
            lv = level[v]
            kk = ((lv - 1) // slab if slab > 0 else 0, int(brank[v[0]]), lv)
            for x in deps[v]:
                kx = key[x]
                cand = (kx[0], kx[1], kx[2] + 1)
                if cand > kk:
                    kk = cand
                    pushed += 1
            key[v] = kk
        # equal keys never depend on each other, and every producer has a smaller key
        assert all(key[x] < key[v] for v in order_ref for x in deps[v])
        order = sorted(order_ref, key=lambda v: (key[v], proc[v[0]], v[1]))
        val = {}
        for (c, s) in order:
            acc = 0.0
            for m in (0, 1):
                op = operand(c, m, s)
                Iu = final0[op[1]] if op[0] == "fixed" else (0.0 if op[0] == "zero" else val[(op[1], op[2])])
                v_ = u[c, m]
                a, bb, e = linear_weights(r[c, m] * (alpha[c] + alpha[v_]) / 2)
                acc += (e * Iu + a * S[v_] + bb * S[c]) * w[c, m]
            val[(c, s)] = acc
        out = final0.copy()
        for c in processed:
            out[c] = val[(c, NS)]
        same = np.array_equal(out, ref)
        print(f"theta {theta:6.1f} phi {phi:6.1f}: blocks {bx}x{by}, slab {slab}: {len(order)} visits, {len(set(key.values()))} distinct keys, "
              f"{pushed} pushes, blocked execution == sequential oracle bit for bit: {same}")
        return same
    # tile key, pushed behind the producers
    ntile = max(1, int(round((n / cells_per_tile) ** (1 / 3))))
    qq = ((pos - pos.min(axis=1, keepdims=True)) / (np.ptp(pos, axis=1)[:, None] + 1e-300) * ntile).astype(np.int64).clip(0, ntile - 1)
    tile = qq[0] + ntile * (qq[1] + ntile * qq[2])
    centre = (qq + 0.5) / ntile * np.ptp(pos, axis=1)[:, None]
    proj = -(k[:, None] * centre).sum(axis=0)
    _, tord = np.unique(np.round(proj / np.ptp(proj) * 1e6).astype(np.int64) * ntile ** 3 + tile, return_inverse=True)
    key = {}
    pushed = 0
    for v in order_ref:
        kk = tord[v[0]] * 1e4 + level[v]
        for x in deps[v]:
            if key[x] >= kk:
                kk = np.nextafter(key[x], np.inf)
                pushed += 1
        key[v] = kk
    order = sorted(order_ref, key=lambda v: (key[v], proc[v[0]], v[1]))
    # execute
    val = {}
    for (c, s) in order:
        acc = 0.0
        for m in (0, 1):
            op = operand(c, m, s)
            Iu = final0[op[1]] if op[0] == "fixed" else (0.0 if op[0] == "zero" else val[(op[1], op[2])])   # KeyError = order not topological
            v = u[c, m]
            a, bb, e = linear_weights(r[c, m] * (alpha[c] + alpha[v]) / 2)
            acc += (e * Iu + a * S[v] + bb * S[c]) * w[c, m]
        val[(c, s)] = acc
    out = final0.copy()
    for c in processed:
        out[c] = val[(c, NS)]
    same = np.array_equal(out, ref)
    print(f"theta {theta:6.1f} phi {phi:6.1f}: {len(order)} visits, {max(level.values())} DAG levels, {pushed} keys pushed behind a producer, "
          f"tile-ordered execution == sequential oracle bit for bit: {same}  (max |diff| {np.abs(out - ref).max():.2e})")
    return same


if __name__ == "__main__":
    ok = all(run(th, ph) for th, ph in ((152.7, 315.5), (67.2, 155.8), (101.8, 235.4), (27.3, 135.5)))
    sys.exit(0 if ok else 1)
