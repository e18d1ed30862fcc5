"""Independent pure-Python restatement of the reference's irregular formal solver, used only to cross-check the
C oracle on small grids (two independent restatements must agree).  Written directly from the
Julia sources, 1-based indices kept inside the functions to stay close to them:
read_cell layers (src/voronoi_utils.jl:93-174, 253-269), calc_Delaunay_lines (:186-245),
smallest_angle (:360-396), Delaunay_upII/downII (src/irregular_ray_tracing.jl:15-163),
linear_weights (src/functions.jl:484-500)."""
import math


def sort_by_layer(nbr, n, wall):
    layers = [0] * (n + 1)
    for i in range(1, n + 1):
        for j in range(1, nbr[i][0] + 1):
            if nbr[i][j] == wall:
                layers[i] = 1
    lower = 1
    while True:
        for i in range(1, n + 1):
            if layers[i] == 0:
                for j in range(1, nbr[i][0] + 1):
                    nb = nbr[i][j]
                    if nb > 0 and layers[nb] == lower:
                        layers[i] = lower + 1
                        break
        if all(layers[i] != 0 for i in range(1, n + 1)):
            break
        lower += 1
    return layers


def sortperm_reduce(layers, n):
    perm = sorted(range(1, n + 1), key=lambda i: layers[i])  # Python's sort is stable, like Julia's sortperm
    srt = [layers[i] for i in perm]
    red = [0] * (max(srt) + 1)
    red[0] = 1
    layer = 2
    for i, v in enumerate(srt, 1):
        if v == layer:
            red[layer - 1] = i
            layer += 1
    red[-1] = n
    return perm, red


def delaunay_line(P, Pn, x_min, x_max, y_min, y_max):
    pn = list(Pn)
    x_r_r, x_r_l = x_max - P[1], P[1] - x_min
    y_r_r, y_r_l = y_max - P[2], P[2] - y_min
    x_i_r, x_i_l = abs(x_max - pn[1]), abs(pn[1] - x_min)
    if x_r_r + x_i_l < P[1] - pn[1]:
        pn[1] = x_max + pn[1] - x_min
    elif x_r_l + x_i_r < pn[1] - P[1]:
        pn[1] = x_min + x_max - pn[1]
    y_i_r, y_i_l = abs(y_max - pn[2]), abs(pn[2] - y_min)
    if y_r_r + y_i_l < P[2] - pn[2]:
        pn[2] = y_max + pn[2] - y_min
    elif y_r_l + y_i_r < pn[2] - P[2]:
        pn[2] = y_min + y_max - pn[2]
    d = [pn[0] - P[0], pn[1] - P[1], pn[2] - P[2]]
    nrm = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    return [d[0] / nrm, d[1] / nrm, d[2] / nrm]


def smallest_angle(i, nbr, pos, k, bounds):
    dots = [-1.0, -1.0]
    ind = [0, 0]
    for j in range(1, nbr[i][0] + 1):
        nb = nbr[i][j]
        if nb > 0:
            l = delaunay_line(pos[i], pos[nb], *bounds)
            d = k[0] * l[0] + k[1] * l[1] + k[2] * l[2]
            if d > dots[1]:
                if d > dots[0]:
                    dots[0], ind[0] = d, nb
                else:
                    dots[1], ind[1] = d, nb
    if dots[1] <= 0:
        dots[1] = 0.0
        ind[1] = ind[0]
    return dots, ind


def linear_weights(dtau):
    if dtau < 5e-4:
        return dtau * (1 / 2 - dtau / 3), dtau * (1 / 2 - dtau / 6), 1 - dtau + 0.5 * dtau ** 2
    if dtau > 50:
        a = 1 / dtau
        return a, 1.0 - a, 0.0
    e = math.exp(-dtau)
    a = (1 - e) / dtau - e
    return a, 1 - a - e, e


def delaunay(k, S, I_0, alpha, n, nbr, pos, bounds, perm, lay, down, n_sweeps=3, p=7.0):
    """S, alpha: 1-based lists (index 0 unused); I_0 list for perm[1:n1]; returns 1-based list I"""
    I = [0.0] * (n + 1)
    lower = lay[1] - 1
    for r in range(lower):
        I[perm[r]] = I_0[r]
    for layer in range(2, len(lay)):
        lo, hi = lay[layer - 1], lay[layer]
        for _ in range(n_sweeps):
            rng = range(hi - 1, lo - 1, -1) if down else range(lo, hi)
            for i in rng:
                idx = perm[i - 1]
                dots, ind = smallest_angle(idx, nbr, pos, k, bounds)
                sw = dots[0] ** p + dots[1] ** p
                w = [dots[0] ** p / sw, dots[1] ** p / sw]
                I[idx] = 0.0
                for m in range(2):
                    u = ind[m]
                    r = math.sqrt(sum((pos[idx][a] - pos[u][a]) ** 2 for a in range(3)))
                    dtau = r * (alpha[idx] + alpha[u]) / 2
                    a, b, e = linear_weights(dtau)
                    I[idx] += (e * I[u] + a * S[u] + b * S[idx]) * w[m]
    return I
