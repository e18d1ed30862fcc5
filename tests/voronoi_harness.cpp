// CPU harness for voronoirt_b200/csrc/voronoi_cell.cuh — TEST INFRASTRUCTURE ONLY (built and loaded by
// tests/test_voronoi_native.py).  It compiles the same host/device cell code the CUDA kernel uses, so the geometry can
// be checked against voro++'s neighbour lists (tests/golden/*.npz) without a GPU.  Not part of the product.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../voronoirt_b200/csrc/voronoi_cell.cuh"
#include "../voronoirt_b200/csrc/sampling.cuh"

using namespace vrt;

static void build_grid(int64_t n, const double* pos, const double* bounds, int gx, int gy, int gz, VoroGrid& G, std::vector<int32_t>& start,
                       std::vector<int32_t>& order) {
    G.gx = gx; G.gy = gy; G.gz = gz;
    G.z0 = bounds[0]; G.Lz = bounds[1] - bounds[0];
    G.x0 = bounds[2]; G.Lx = bounds[3] - bounds[2];
    G.y0 = bounds[4]; G.Ly = bounds[5] - bounds[4];
    G.hx = G.Lx / gx; G.hy = G.Ly / gy; G.hz = G.Lz / gz;
    const int64_t nc = (int64_t)gx * gy * gz;
    start.assign(nc + 1, 0);
    order.resize(n);
    std::vector<int32_t> cellof(n);
    for (int64_t i = 0; i < n; i++) {
        int ix = (int)((pos[3 * i + 1] - G.x0) / G.hx), iy = (int)((pos[3 * i + 2] - G.y0) / G.hy), iz = (int)((pos[3 * i] - G.z0) / G.hz);
        ix = ix < 0 ? 0 : (ix >= gx ? gx - 1 : ix);
        iy = iy < 0 ? 0 : (iy >= gy ? gy - 1 : iy);
        iz = iz < 0 ? 0 : (iz >= gz ? gz - 1 : iz);
        cellof[i] = (int32_t)(ix + gx * (iy + gy * iz));
        start[cellof[i] + 1]++;
    }
    for (int64_t c = 0; c < nc; c++) start[c + 1] += start[c];
    std::vector<int32_t> fill(start.begin(), start.end() - 1);
    for (int64_t i = 0; i < n; i++) order[fill[cellof[i]]++] = (int32_t)i;
    G.start = start.data(); G.order = order.data(); G.pos = pos;
}

extern "C" void vc_nn_harness(int64_t n, const double* pos, const double* bounds, int gx, int gy, int gz, int64_t m, const double* q,
                              int64_t* idx, double* d2) {
    VoroGrid G;
    std::vector<int32_t> start, order;
    build_grid(n, pos, bounds, gx, gy, gz, G, start, order);
    for (int64_t k = 0; k < m; k++) idx[k] = nearest_site_of(G, q[3 * k], q[3 * k + 1], q[3 * k + 2], d2 + k) + 1;
}

extern "C" int vc_harness(int64_t n, const double* pos, const double* bounds, int gx, int gy, int gz, int64_t* nbr, int64_t ld,
                          int32_t* status) {
    VoroGrid G;
    G.gx = gx; G.gy = gy; G.gz = gz;
    G.z0 = bounds[0]; G.Lz = bounds[1] - bounds[0];
    G.x0 = bounds[2]; G.Lx = bounds[3] - bounds[2];
    G.y0 = bounds[4]; G.Ly = bounds[5] - bounds[4];
    G.hx = G.Lx / gx; G.hy = G.Ly / gy; G.hz = G.Lz / gz;
    const int64_t nc = (int64_t)gx * gy * gz;
    std::vector<int32_t> start(nc + 1, 0), order(n), cellof(n);
    for (int64_t i = 0; i < n; i++) {
        int ix = (int)((pos[3 * i + 1] - G.x0) / G.hx), iy = (int)((pos[3 * i + 2] - G.y0) / G.hy), iz = (int)((pos[3 * i] - G.z0) / G.hz);
        ix = ix < 0 ? 0 : (ix >= gx ? gx - 1 : ix);
        iy = iy < 0 ? 0 : (iy >= gy ? gy - 1 : iy);
        iz = iz < 0 ? 0 : (iz >= gz ? gz - 1 : iz);
        cellof[i] = (int32_t)(ix + gx * (iy + gy * iz));
        start[cellof[i] + 1]++;
    }
    for (int64_t c = 0; c < nc; c++) start[c + 1] += start[c];
    std::vector<int32_t> fill(start.begin(), start.end() - 1);
    for (int64_t i = 0; i < n; i++) order[fill[cellof[i]]++] = (int32_t)i;
    G.start = start.data(); G.order = order.data(); G.pos = pos;
    int bad = 0;
    ConvexCell* cell = new ConvexCell();
    for (int64_t i = 0; i < n; i++) {
        int64_t* row = nbr + i * ld;
        int cnt = voronoi_cell_of(G, n, i, *cell, row + 1, (int)(ld - 1));
        status[i] = cell->status;
        row[0] = cnt;
        if (cnt < 0 || cnt > ld - 1) bad++;
    }
    delete cell;
    return bad;
}

extern "C" void vc_philox(uint32_t* ctr, uint32_t k0, uint32_t k1) { philox4x32_10(ctr, k0, k1); }

extern "C" int64_t vc_sample_harness(int64_t n_sites, int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                                     const double* q, uint64_t seed, double q_min, double dq, double* pos) {
    TriGrid T = {nz, nx, ny, z, x, y, q};
    int64_t trials = 0;
    for (int64_t i = 0; i < n_sites; i++) {
        int64_t t = rejection_site(T, seed, i, q_min, dq, 1 << 20, pos + 3 * i);
        if (t < 0) return -1;
        trials += t;
    }
    return trials;
}

extern "C" int64_t vc_trilinear_harness(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, const double* v,
                                        int64_t n, const double* pos, double* out) {
    TriGrid T = {nz, nx, ny, z, x, y, v};
    int64_t bad = 0;
    for (int64_t k = 0; k < n; k++)
        if (!trilinear_at(T, pos[3 * k], pos[3 * k + 1], pos[3 * k + 2], out + k)) { out[k] = NAN; bad++; }
    return bad;
}

extern "C" void vc_knn_harness(int64_t n, const double* pos, const double* bounds, int gx, int gy, int gz, int64_t m, const double* q, int k,
                               int64_t* idx, double* d2) {
    VoroGrid G;
    std::vector<int32_t> start, order;
    build_grid(n, pos, bounds, gx, gy, gz, G, start, order);
    for (int64_t p = 0; p < m; p++) {
        int got = nearest_k_sites_of(G, n, k, q[3 * p], q[3 * p + 1], q[3 * p + 2], idx + p * k, d2 + p * k);
        for (int t = 0; t < got; t++) idx[p * k + t] += 1;
    }
}

extern "C" int64_t vc_corner_harness(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, const double* v,
                                     int64_t n, const double* pos, double* out) {
    TriGrid T = {nz, nx, ny, z, x, y, v};
    int64_t bad = 0;
    for (int64_t k = 0; k < n; k++)
        if (!nearest_corner_at(T, pos[3 * k], pos[3 * k + 1], pos[3 * k + 2], out + k)) { out[k] = NAN; bad++; }
    return bad;
}
