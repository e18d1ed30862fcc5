// voronoi_cell.cuh — one Voronoi cell by half-space clipping, for the native neighbour generation that replaces the
// voro++ text round trip (reference: rt_preprocessing/output_sites.cc:35-49 `container(..., true, true,
// false, 8)` + `print_custom("%i %n")`, i/o src/io.jl:8-40, parsing src/voronoi_utils.jl:42-70).
//
// What has to be reproduced is, per site, the SET of face neighbours (1-based site ids; walls -5 = z_min, -6 = z_max) of
// the Voronoi tessellation of the box that is periodic in x and y and walled in z.  The order in which voro++ prints
// the faces is an artefact of its cell construction and is not defined by the reference.
//
// Algorithm (the dual-mesh clipping of Ray, Sokolov, Lefebvre & Lévy 2018): the cell is the intersection of half-spaces
// a x + b y + c z + d >= 0 in coordinates relative to the site; it is stored as its dual triangulation: a triangle
// (u, v, w) of plane indices is the cell vertex where the three planes meet, consistently oriented.  Clipping by a new
// plane removes the vertices outside it; every boundary edge (u, v) of the removed region gets the new vertex (u, v, p).
// Candidates come from a uniform grid in rings of growing Chebyshev distance until the security radius is reached: no
// site farther than twice the farthest vertex can cut the cell.
//
// Host/device code: the CUDA kernel (tessellate.cu) and the CPU harness of the tests (tests/voronoi_harness.cpp, test
// infrastructure only) compile this same file, so the geometry is validated against voro++'s lists without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define VC_HD __host__ __device__ __forceinline__
#else
#define VC_HD inline
#endif

namespace vrt {

constexpr int VC_MAXP = 64;    // planes kept per cell (6 box planes + contributing bisectors, never compacted)
constexpr int VC_MAXT = 124;   // vertices (dual triangles); a cell with f faces has 2f - 4

enum { VC_OK = 0, VC_OVERFLOW = 1, VC_EMPTY = 2 };

struct VoroGrid {
    int gx, gy, gz;             // cells per axis
    double x0, y0, z0;          // box minimum
    double Lx, Ly, Lz;          // box extent (x, y periodic)
    double hx, hy, hz;          // cell size
    const int32_t* start;       // gx*gy*gz + 1 offsets into `order`
    const int32_t* order;       // site indices (0-based) sorted by cell
    const double* pos;          // 3 x n (z, x, y) as at the ABI
};

struct ConvexCell {
    double pa[VC_MAXP], pb[VC_MAXP], pc[VC_MAXP], pd[VC_MAXP];
    int64_t pid[VC_MAXP];
    unsigned char t0[VC_MAXT], t1[VC_MAXT], t2[VC_MAXT], rem[VC_MAXT];
    double vx[VC_MAXT], vy[VC_MAXT], vz[VC_MAXT];
    int np, nt, status;
    double r2max;

    VC_HD void vertex(int t) {
        const int u = t0[t], v = t1[t], w = t2[t];
        const double ax = pa[u], ay = pb[u], az = pc[u], bx = pa[v], by = pb[v], bz = pc[v], cx = pa[w], cy = pb[w], cz = pc[w];
        // n2 x n3, n3 x n1, n1 x n2
        const double ux = by * cz - bz * cy, uy = bz * cx - bx * cz, uz = bx * cy - by * cx;
        const double wx = cy * az - cz * ay, wy = cz * ax - cx * az, wz = cx * ay - cy * ax;
        const double qx = ay * bz - az * by, qy = az * bx - ax * bz, qz = ax * by - ay * bx;
        const double det = ax * ux + ay * uy + az * uz;
        const double s = -1.0 / det;
        vx[t] = s * (pd[u] * ux + pd[v] * wx + pd[w] * qx);
        vy[t] = s * (pd[u] * uy + pd[v] * wy + pd[w] * qy);
        vz[t] = s * (pd[u] * uz + pd[v] * wz + pd[w] * qz);
    }

    VC_HD void update_r2() {
        double m = 0.0;
        for (int t = 0; t < nt; t++) {
            const double d = vx[t] * vx[t] + vy[t] * vy[t] + vz[t] * vz[t];
            m = d > m ? d : m;
        }
        r2max = m;
    }

    // box around the site: periodic extents in x, y (the bisectors with the site's own images, id = the site itself),
    // walls in z.  Coordinates are (x, y, z) relative to the site.
    VC_HD void init(double Lx, double Ly, double dz_lo, double dz_hi, int64_t self_id) {
        np = 6;
        status = VC_OK;
        pa[0] = 1;  pb[0] = 0;  pc[0] = 0;  pd[0] = 0.5 * Lx; pid[0] = self_id;
        pa[1] = -1; pb[1] = 0;  pc[1] = 0;  pd[1] = 0.5 * Lx; pid[1] = self_id;
        pa[2] = 0;  pb[2] = 1;  pc[2] = 0;  pd[2] = 0.5 * Ly; pid[2] = self_id;
        pa[3] = 0;  pb[3] = -1; pc[3] = 0;  pd[3] = 0.5 * Ly; pid[3] = self_id;
        pa[4] = 0;  pb[4] = 0;  pc[4] = 1;  pd[4] = dz_lo;    pid[4] = -5;   // z >= z_min
        pa[5] = 0;  pb[5] = 0;  pc[5] = -1; pd[5] = dz_hi;    pid[5] = -6;   // z <= z_max
        nt = 0;
        // the dual of the box is an octahedron; faces oriented consistently (outward)
        for (int sx = 0; sx < 2; sx++)
            for (int sy = 0; sy < 2; sy++)
                for (int sz = 0; sz < 2; sz++) {
                    const int X = sx, Y = 2 + sy, Z = 4 + sz;
                    const bool even = ((sx + sy + sz) & 1) == 0;
                    t0[nt] = (unsigned char)X;
                    t1[nt] = (unsigned char)(even ? Y : Z);
                    t2[nt] = (unsigned char)(even ? Z : Y);
                    vertex(nt);
                    nt++;
                }
        update_r2();
    }

    // drop the planes that no longer own a vertex (the six box planes keep their slots) and renumber the triangles
    VC_HD void compact_planes() {
        unsigned char map[VC_MAXP];
        int w = 6;
        for (int p = 0; p < 6; p++) map[p] = (unsigned char)p;
        for (int p = 6; p < np; p++) {
            bool used = false;
            for (int t = 0; t < nt && !used; t++) used = t0[t] == p || t1[t] == p || t2[t] == p;
            if (used) {
                pa[w] = pa[p]; pb[w] = pb[p]; pc[w] = pc[p]; pd[w] = pd[p]; pid[w] = pid[p];
                map[p] = (unsigned char)w++;
            } else {
                map[p] = 0;
            }
        }
        for (int t = 0; t < nt; t++) { t0[t] = map[t0[t]]; t1[t] = map[t1[t]]; t2[t] = map[t2[t]]; }
        np = w;
    }

    // clip by a x + b y + c z + d >= 0; returns true when the plane cuts the cell (it then stays as a face candidate)
    VC_HD bool clip(double a, double b, double c, double d, int64_t id) {
        // a vertex counts as cut only when it is outside by more than a relative tolerance (voro++ does the same with its
        // 1e-11): the bisector with a diagonal periodic image of a site of a nearly empty box passes exactly through an
        // edge of the cell and must not become a zero-area face
        const double tol = -1e-11 * d;
        int nrem = 0;
        for (int t = 0; t < nt; t++) {
            const bool out = a * vx[t] + b * vy[t] + c * vz[t] + d < tol;
            rem[t] = out ? 1 : 0;
            nrem += out ? 1 : 0;
        }
        if (nrem == 0) return false;
        if (nrem == nt) { status = VC_EMPTY; return false; }
        if (np >= VC_MAXP) {
            compact_planes();   // planes cut away since they were added still occupy slots
            if (np >= VC_MAXP) { status = VC_OVERFLOW; return false; }
        }
        const int p = np++;
        pa[p] = a; pb[p] = b; pc[p] = c; pd[p] = d; pid[p] = id;
        // new vertices along the boundary of the removed region (appended after the old ones, flagged 2 = new)
        const int nt_old = nt;
        for (int t = 0; t < nt_old; t++) {
            if (rem[t] != 1) continue;
            const int e[3][2] = {{t0[t], t1[t]}, {t1[t], t2[t]}, {t2[t], t0[t]}};
            for (int k = 0; k < 3; k++) {
                // the neighbour across (x, y) holds the directed edge (y, x)
                bool nb_removed = false;
                for (int q = 0; q < nt_old; q++) {
                    const int u = t0[q], v = t1[q], w = t2[q];
                    const int x = e[k][1], y = e[k][0];
                    if ((u == x && v == y) || (v == x && w == y) || (w == x && u == y)) { nb_removed = rem[q] == 1; break; }
                }
                if (nb_removed) continue;
                if (nt >= VC_MAXT) { status = VC_OVERFLOW; return false; }
                t0[nt] = (unsigned char)e[k][0];
                t1[nt] = (unsigned char)e[k][1];
                t2[nt] = (unsigned char)p;
                rem[nt] = 2;
                vertex(nt);
                nt++;
            }
        }
        // compact
        int w = 0;
        for (int t = 0; t < nt; t++) {
            if (rem[t] == 1) continue;
            if (w != t) { t0[w] = t0[t]; t1[w] = t1[t]; t2[w] = t2[t]; vx[w] = vx[t]; vy[w] = vy[t]; vz[w] = vz[t]; }
            w++;
        }
        nt = w;
        update_r2();
        return true;
    }

    // planes that still own a vertex = faces of the cell; ids written in plane order; returns the count.
    // A face whose vertices are (nearly) collinear has no area: it is the trace of a bisector that only touches an edge
    // or a corner of the cell (lattice-aligned sites) and is dropped, like voro++ drops it.
    VC_HD int faces(int64_t* out, int cap) const {
        int n = 0;
        const double eps2 = 1e-24 * r2max;   // (1e-12 of the cell radius)^2
        for (int p = 0; p < np; p++) {
            bool used = false;
            double ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0, far2 = 0;
            int first = -1;
            for (int t = 0; t < nt; t++) {
                if (!(t0[t] == p || t1[t] == p || t2[t] == p)) continue;
                if (first < 0) { first = t; ax = vx[t]; ay = vy[t]; az = vz[t]; continue; }
                const double ex = vx[t] - ax, ey = vy[t] - ay, ez = vz[t] - az;
                const double d2 = ex * ex + ey * ey + ez * ez;
                if (d2 > far2) { far2 = d2; bx = ex; by = ey; bz = ez; }
            }
            if (first >= 0 && far2 > eps2) {
                // height of the remaining vertices over the line a -> a + b
                for (int t = 0; t < nt && !used; t++) {
                    if (!(t0[t] == p || t1[t] == p || t2[t] == p)) continue;
                    const double ex = vx[t] - ax, ey = vy[t] - ay, ez = vz[t] - az;
                    const double cx = ey * bz - ez * by, cy = ez * bx - ex * bz, cz = ex * by - ey * bx;
                    used = (cx * cx + cy * cy + cz * cz) > eps2 * far2;   // |e x b|^2 = h^2 |b|^2
                }
            }
            if (used) {
                if (n < cap) out[n] = pid[p];
                n++;
            }
        }
        return n;
    }
};

VC_HD int vc_floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// the cell of site i (0-based) -> number of faces, ids in out[0..cap); cell.status tells about overflow
VC_HD int voronoi_cell_of(const VoroGrid& G, int64_t n, int64_t i, ConvexCell& cell, int64_t* out, int cap) {
    const double zi = G.pos[3 * i], xi = G.pos[3 * i + 1], yi = G.pos[3 * i + 2];
    cell.init(G.Lx, G.Ly, zi - G.z0, G.z0 + G.Lz - zi, i + 1);
    int ix0 = (int)((xi - G.x0) / G.hx), iy0 = (int)((yi - G.y0) / G.hy), iz0 = (int)((zi - G.z0) / G.hz);
    ix0 = ix0 < 0 ? 0 : (ix0 >= G.gx ? G.gx - 1 : ix0);
    iy0 = iy0 < 0 ? 0 : (iy0 >= G.gy ? G.gy - 1 : iy0);
    iz0 = iz0 < 0 ? 0 : (iz0 >= G.gz ? G.gz - 1 : iz0);
    const double hmin = G.hx < G.hy ? (G.hx < G.hz ? G.hx : G.hz) : (G.hy < G.hz ? G.hy : G.hz);
    for (int r = 0;; r++) {
        if (r > 0) {
            const double dmin = (r - 1) * hmin;   // no site of ring r is closer than this
            if (dmin * dmin > 4.0 * cell.r2max) break;
        }
        for (int dz = -r; dz <= r; dz++) {
            const int iz = iz0 + dz;
            if (iz < 0 || iz >= G.gz) continue;
            const bool zface = dz == -r || dz == r;
            for (int dy = -r; dy <= r; dy++) {
                const bool yface = dy == -r || dy == r;
                const int wy = vc_floor_div(iy0 + dy, G.gy), iy = iy0 + dy - wy * G.gy;
                const double sy = wy * G.Ly;
                const int step = (zface || yface) ? 1 : (2 * r > 0 ? 2 * r : 1);   // interior of the slab: only dx = -r, r
                for (int dx = -r; dx <= r; dx += step) {
                    const int wx = vc_floor_div(ix0 + dx, G.gx), ix = ix0 + dx - wx * G.gx;
                    const double sx = wx * G.Lx;
                    const int64_t c = ix + (int64_t)G.gx * (iy + (int64_t)G.gy * iz);
                    for (int32_t k = G.start[c]; k < G.start[c + 1]; k++) {
                        const int64_t j = G.order[k];
                        if (j == i && wx == 0 && wy == 0) continue;
                        const double qz = G.pos[3 * j] - zi, qx = G.pos[3 * j + 1] + sx - xi, qy = G.pos[3 * j + 2] + sy - yi;
                        const double d2 = qx * qx + qy * qy + qz * qz;
                        if (d2 > 4.0 * cell.r2max) continue;
                        cell.clip(-qx, -qy, -qz, 0.5 * d2, j + 1);
                        if (cell.status != VC_OK) return -1;
                    }
                }
            }
        }
        if (r > 8 * (G.gx + G.gy + G.gz) + 16) break;   // safety only: the box planes bound r2max, so the ring test above ends the loop
    }
    (void)n;
    return cell.faces(out, cap);
}

// Nearest site of a query point (z, x, y) in plain Euclidean distance, no periodic wrap — what `nn(KDTree(positions), p)`
// returns in Voronoi_to_Raster (voronoi_utils.jl:442-444).  Ring search over the same grid; ties go to the smaller index.
// Every product and sum of the squared distance is rounded separately so that a host evaluation gives the same argmin.
VC_HD double vc_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
VC_HD double vc_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

VC_HD int64_t nearest_site_of(const VoroGrid& G, double qz, double qx, double qy, double* d2_out) {
    int ix0 = (int)floor((qx - G.x0) / G.hx), iy0 = (int)floor((qy - G.y0) / G.hy), iz0 = (int)floor((qz - G.z0) / G.hz);
    const bool inside = ix0 >= 0 && ix0 < G.gx && iy0 >= 0 && iy0 < G.gy && iz0 >= 0 && iz0 < G.gz;
    ix0 = ix0 < 0 ? 0 : (ix0 >= G.gx ? G.gx - 1 : ix0);
    iy0 = iy0 < 0 ? 0 : (iy0 >= G.gy ? G.gy - 1 : iy0);
    iz0 = iz0 < 0 ? 0 : (iz0 >= G.gz ? G.gz - 1 : iz0);
    const double hmin = G.hx < G.hy ? (G.hx < G.hz ? G.hx : G.hz) : (G.hy < G.hz ? G.hy : G.hz);
    const int rmax = (G.gx > G.gy ? (G.gx > G.gz ? G.gx : G.gz) : (G.gy > G.gz ? G.gy : G.gz));
    double best = 1.0e300;
    int64_t arg = -1;
    for (int r = 0; r <= rmax; r++) {
        // sites of ring r are at least (r-1)*hmin away from a query inside the grid
        if (inside && r > 0 && arg >= 0) {
            const double dmin = (r - 1) * hmin;
            if (dmin * dmin > best) break;
        }
        for (int dz = -r; dz <= r; dz++) {
            const int iz = iz0 + dz;
            if (iz < 0 || iz >= G.gz) continue;
            const bool zface = dz == -r || dz == r;
            for (int dy = -r; dy <= r; dy++) {
                const int iy = iy0 + dy;
                if (iy < 0 || iy >= G.gy) continue;
                const bool yface = dy == -r || dy == r;
                const int step = (zface || yface) ? 1 : (2 * r > 0 ? 2 * r : 1);
                for (int dx = -r; dx <= r; dx += step) {
                    const int ix = ix0 + dx;
                    if (ix < 0 || ix >= G.gx) continue;
                    const int64_t c = ix + (int64_t)G.gx * (iy + (int64_t)G.gy * iz);
                    for (int32_t k = G.start[c]; k < G.start[c + 1]; k++) {
                        const int64_t j = G.order[k];
                        const double ez = G.pos[3 * j] - qz, ex = G.pos[3 * j + 1] - qx, ey = G.pos[3 * j + 2] - qy;
                        const double d2 = vc_add(vc_add(vc_mul(ez, ez), vc_mul(ex, ex)), vc_mul(ey, ey));
                        if (d2 < best || (d2 == best && j < arg)) { best = d2; arg = j; }
                    }
                }
            }
        }
    }
    if (d2_out) *d2_out = best;
    return arg;
}

// The k nearest sites (k <= VC_KMAX), ascending by (distance, index): `knn(tree, p, n_k)` of Voronoi_to_Raster_inv_dist
// (voronoi_utils.jl:797-805; the reference does not sort its result, the inverse-distance sum does not depend on the order).
constexpr int VC_KMAX = 8;

VC_HD int nearest_k_sites_of(const VoroGrid& G, int64_t n, int k, double qz, double qx, double qy, int64_t* idx_out, double* d2_out) {
    if (k > VC_KMAX) k = VC_KMAX;
    if ((int64_t)k > n) k = (int)n;
    int ix0 = (int)floor((qx - G.x0) / G.hx), iy0 = (int)floor((qy - G.y0) / G.hy), iz0 = (int)floor((qz - G.z0) / G.hz);
    const bool inside = ix0 >= 0 && ix0 < G.gx && iy0 >= 0 && iy0 < G.gy && iz0 >= 0 && iz0 < G.gz;
    ix0 = ix0 < 0 ? 0 : (ix0 >= G.gx ? G.gx - 1 : ix0);
    iy0 = iy0 < 0 ? 0 : (iy0 >= G.gy ? G.gy - 1 : iy0);
    iz0 = iz0 < 0 ? 0 : (iz0 >= G.gz ? G.gz - 1 : iz0);
    const double hmin = G.hx < G.hy ? (G.hx < G.hz ? G.hx : G.hz) : (G.hy < G.hz ? G.hy : G.hz);
    const int rmax = (G.gx > G.gy ? (G.gx > G.gz ? G.gx : G.gz) : (G.gy > G.gz ? G.gy : G.gz));
    double bd[VC_KMAX];
    int64_t bi[VC_KMAX];
    int have = 0;
    for (int r = 0; r <= rmax; r++) {
        if (inside && r > 0 && have == k) {
            const double dmin = (r - 1) * hmin;
            if (dmin * dmin > bd[k - 1]) break;
        }
        for (int dz = -r; dz <= r; dz++) {
            const int iz = iz0 + dz;
            if (iz < 0 || iz >= G.gz) continue;
            const bool zface = dz == -r || dz == r;
            for (int dy = -r; dy <= r; dy++) {
                const int iy = iy0 + dy;
                if (iy < 0 || iy >= G.gy) continue;
                const bool yface = dy == -r || dy == r;
                const int step = (zface || yface) ? 1 : (2 * r > 0 ? 2 * r : 1);
                for (int dx = -r; dx <= r; dx += step) {
                    const int ix = ix0 + dx;
                    if (ix < 0 || ix >= G.gx) continue;
                    const int64_t c = ix + (int64_t)G.gx * (iy + (int64_t)G.gy * iz);
                    for (int32_t q = G.start[c]; q < G.start[c + 1]; q++) {
                        const int64_t j = G.order[q];
                        const double ez = G.pos[3 * j] - qz, ex = G.pos[3 * j + 1] - qx, ey = G.pos[3 * j + 2] - qy;
                        const double d2 = vc_add(vc_add(vc_mul(ez, ez), vc_mul(ex, ex)), vc_mul(ey, ey));
                        if (have == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && j < bi[k - 1]))) continue;
                        int p = have < k ? have++ : k - 1;   // insertion into the sorted list
                        while (p > 0 && (d2 < bd[p - 1] || (d2 == bd[p - 1] && j < bi[p - 1]))) {
                            bd[p] = bd[p - 1]; bi[p] = bi[p - 1];
                            p--;
                        }
                        bd[p] = d2; bi[p] = j;
                    }
                }
            }
        }
    }
    for (int p = 0; p < have; p++) { idx_out[p] = bi[p]; d2_out[p] = bd[p]; }
    return have;
}

}  // namespace vrt
