// tessellate.cu — native Voronoi neighbour generation on the GPU.
//
// Replaces the reference's preprocessing round trip: write_arrays (io.jl:8-40) -> the single-threaded voro++ driver
// (rt_preprocessing/output_sites.cc: container periodic in x, y, walls in z, `print_custom("%i %n")`) -> the text parser of
// read_cell (voronoi_utils.jl:42-70).  Output is the NeighbourMatrix read_cell builds: n x ld, column 0 = number of
// faces, then the 1-based ids of the face neighbours, -5 / -6 for the z_min / z_max walls.  The neighbour SETS equal
// voro++'s (tests: the committed voro++ lists of tests/golden/*.npz, and a 100 k-site stratified box); the order inside a
// row is this file's (box planes first, then by ring of the search grid), not voro++'s, which the reference leaves undefined.
//
// K-T1 k_vt_cell_index: uniform grid cell of every site;  CUB radix sort by cell;  K-T2 k_vt_cell_start: offsets;
// K-T3 k_vt_cells: one thread per site (in cell order, so neighbouring threads walk the same grid cells) clips its cell
// with voronoi_cell.cuh; the cell (planes + dual triangles, 6 KB) lives in thread-local memory;  K-T4 k_vt_emit: rows of the
// int32 scratch -> the caller's column-major int64 matrix.
#include <algorithm>
#include <math.h>
#include <cub/cub.cuh>
#include "voronoi_cell.cuh"
#include "vrt_internal.h"

namespace vrt {
namespace {

constexpr int VT_CAP = 64;   // faces per cell the scratch can hold (column 0 + 63 ids)

__global__ void k_vt_cell_index(int64_t n, const double* __restrict__ pos, VoroGrid G, int32_t* __restrict__ cell, int32_t* __restrict__ iota) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ix = (int)((pos[3 * i + 1] - G.x0) / G.hx), iy = (int)((pos[3 * i + 2] - G.y0) / G.hy), iz = (int)((pos[3 * i] - G.z0) / G.hz);
    ix = ix < 0 ? 0 : (ix >= G.gx ? G.gx - 1 : ix);
    iy = iy < 0 ? 0 : (iy >= G.gy ? G.gy - 1 : iy);
    iz = iz < 0 ? 0 : (iz >= G.gz ? G.gz - 1 : iz);
    cell[i] = ix + G.gx * (iy + G.gy * iz);
    iota[i] = (int32_t)i;
}

// start[c] = first position in the sorted list whose cell is >= c (start has nc + 1 entries)
__global__ void k_vt_cell_start(int64_t n, int64_t nc, const int32_t* __restrict__ sorted_cell, int32_t* __restrict__ start) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > n) return;
    const int64_t lo = k == 0 ? 0 : (int64_t)sorted_cell[k - 1] + 1;
    const int64_t hi = k == n ? nc : (int64_t)sorted_cell[k];
    for (int64_t c = lo; c <= hi; c++) start[c] = (int32_t)k;
}

__global__ void __launch_bounds__(128) k_vt_cells(int64_t n, VoroGrid G, int32_t* __restrict__ rows, int32_t* __restrict__ status) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = G.order[k];
    ConvexCell cell;
    int64_t ids[VT_CAP];
    const int cnt = voronoi_cell_of(G, n, i, cell, ids, VT_CAP - 1);
    int32_t* row = rows + i * VT_CAP;
    int st = cell.status;
    if (st == VC_OK && cnt > VT_CAP - 1) st = VC_OVERFLOW;
    status[i] = st;
    row[0] = st == VC_OK ? cnt : 0;
    if (st == VC_OK)
        for (int q = 0; q < cnt; q++) row[1 + q] = (int32_t)ids[q];
}

__global__ void k_vt_emit(int64_t n, int64_t ld, const int32_t* __restrict__ rows, int64_t* __restrict__ nbr) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * ld) return;
    const int64_t c = idx / n, i = idx - c * n;   // column-major: consecutive threads write consecutive rows of a column
    const int32_t cnt = rows[i * VT_CAP];
    nbr[idx] = c == 0 ? cnt : (c <= cnt && c < VT_CAP ? rows[i * VT_CAP + c] : 0);
}

__global__ void k_vt_nearest(int64_t m, VoroGrid G, const double* __restrict__ q, int64_t* __restrict__ idx, double* __restrict__ dist) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= m) return;
    double d2;
    const int64_t j = nearest_site_of(G, q[3 * k], q[3 * k + 1], q[3 * k + 2], &d2);
    idx[k] = j + 1;
    if (dist) dist[k] = __dsqrt_rn(d2);
}

__global__ void k_vt_nearest_k(int64_t n, int64_t m, int k, VoroGrid G, const double* __restrict__ q, int64_t* __restrict__ idx,
                               double* __restrict__ dist) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= m) return;
    int64_t bi[VC_KMAX];
    double bd[VC_KMAX];
    const int got = nearest_k_sites_of(G, n, k, q[3 * p], q[3 * p + 1], q[3 * p + 2], bi, bd);
    for (int t = 0; t < k; t++) {
        idx[p * k + t] = t < got ? bi[t] + 1 : 0;
        if (dist) dist[p * k + t] = t < got ? __dsqrt_rn(bd[t]) : nan("");
    }
}

// the uniform search grid over the sites: cell of every site, CUB radix sort by cell, cell offsets
struct SearchGrid {
    VoroGrid G;
    DevBuf<double> dpos;
    DevBuf<int32_t> cell, cell_sorted, iota, order, start;
    int build(const char* who, int64_t n, const double* positions, const double* bounds, double sites_per_cell) {
        double hb[6];
        VRT_CUDA(cudaMemcpy(hb, bounds, sizeof(hb), cudaMemcpyDefault));
        G.z0 = hb[0]; G.Lz = hb[1] - hb[0];
        G.x0 = hb[2]; G.Lx = hb[3] - hb[2];
        G.y0 = hb[4]; G.Ly = hb[5] - hb[4];
        if (!(G.Lx > 0) || !(G.Ly > 0) || !(G.Lz > 0)) {
            set_error("%s: empty box", who);
            return VRT_E_INVALID;
        }
        // near-cubic search cells
        const double h = cbrt(G.Lx * G.Ly * G.Lz * sites_per_cell / (double)n);
        auto cells = [&](double L) { return (int)std::max<int64_t>(1, std::min<int64_t>(1024, llround(L / h))); };
        G.gx = cells(G.Lx); G.gy = cells(G.Ly); G.gz = cells(G.Lz);
        G.hx = G.Lx / G.gx; G.hy = G.Ly / G.gy; G.hz = G.Lz / G.gz;
        const int64_t nc = (int64_t)G.gx * G.gy * G.gz;
        const double* pos = positions;
        if (!is_device_ptr(positions)) {
            VRT_TRY(dpos.alloc((size_t)3 * n));
            VRT_CUDA(cudaMemcpy(dpos.p, positions, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
            pos = dpos.p;
        }
        VRT_TRY(cell.alloc(n)); VRT_TRY(cell_sorted.alloc(n)); VRT_TRY(iota.alloc(n)); VRT_TRY(order.alloc(n));
        VRT_TRY(start.alloc(nc + 1));
        const int bs = 256;
        k_vt_cell_index<<<(unsigned)((n + bs - 1) / bs), bs>>>(n, pos, G, cell.p, iota.p);
        VRT_CUDA(cudaGetLastError());
        size_t tmp_bytes = 0;
        int bits = 1;
        while (((int64_t)1 << bits) < nc) bits++;
        VRT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, cell.p, cell_sorted.p, iota.p, order.p, (int)n, 0, bits));
        DevBuf<char> tmp;
        VRT_TRY(tmp.alloc(tmp_bytes));
        VRT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, cell.p, cell_sorted.p, iota.p, order.p, (int)n, 0, bits));
        k_vt_cell_start<<<(unsigned)((n + 1 + bs - 1) / bs), bs>>>(n, nc, cell_sorted.p, start.p);
        VRT_CUDA(cudaGetLastError());
        VRT_CUDA(cudaDeviceSynchronize());   // `tmp` goes out of scope
        G.start = start.p; G.order = order.p; G.pos = pos;
        return VRT_OK;
    }
};

__global__ void k_vt_reduce(int64_t n, const int32_t* __restrict__ rows, const int32_t* __restrict__ status, int32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicMax(&out[0], rows[i * VT_CAP]);
    if (status[i] != VC_OK) atomicAdd(&out[1], 1);
}

}  // namespace
}  // namespace vrt

using namespace vrt;

extern "C" int vrt_voronoi_neighbours(int64_t n, const double* positions, const double bounds[6], int64_t* nbr, int64_t ld,
                                      int64_t* ld_needed) {
    if (n <= 0 || !positions || !bounds || n >= ((int64_t)1 << 31)) {
        set_error("vrt_voronoi_neighbours: bad arguments");
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("vrt_voronoi_neighbours: no CUDA device (this library has no CPU path)");
        return VRT_E_CUDA;
    }
    cudaEvent_t e0, e1;
    VRT_CUDA(cudaEventCreate(&e0));
    VRT_CUDA(cudaEventCreate(&e1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{e0, e1};
    VRT_CUDA(cudaEventRecord(e0));
    SearchGrid SG;
    VRT_TRY(SG.build("vrt_voronoi_neighbours", n, positions, bounds, 4.0));   // about four sites per cell
    const VoroGrid& G = SG.G;
    DevBuf<int32_t> rows, status, red;
    VRT_TRY(rows.alloc((size_t)n * VT_CAP)); VRT_TRY(status.alloc(n)); VRT_TRY(red.alloc(2));
    const int bs = 256;
    k_vt_cells<<<(unsigned)((n + 127) / 128), 128>>>(n, G, rows.p, status.p);
    VRT_CUDA(cudaGetLastError());
    VRT_CUDA(cudaMemset(red.p, 0, sizeof(int32_t) * 2));
    k_vt_reduce<<<(unsigned)((n + bs - 1) / bs), bs>>>(n, rows.p, status.p, red.p);
    VRT_CUDA(cudaGetLastError());
    VRT_CUDA(cudaEventRecord(e1));
    int32_t hred[2] = {0, 0};
    VRT_CUDA(cudaMemcpy(hred, red.p, sizeof(hred), cudaMemcpyDeviceToHost));
    float ms = 0;
    VRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    SweepStats st;
    st.kernels = 6;   // index, 2 sort passes counted as one each, start, cells, reduce (+ emit below)
    st.sweep_ms = ms;
    if (hred[1] > 0) {
        set_error("vrt_voronoi_neighbours: %d cells exceed the per-cell capacity (%d planes / %d faces): degenerate input?", (int)hred[1],
                  VC_MAXP, VT_CAP - 1);
        return VRT_E_GRID;
    }
    if (ld_needed) *ld_needed = (int64_t)hred[0] + 1;
    if (nbr) {
        if (ld < (int64_t)hred[0] + 1) {
            set_error("vrt_voronoi_neighbours: ld = %lld but %d columns are needed", (long long)ld, (int)hred[0] + 1);
            return VRT_E_INVALID;
        }
        DevBuf<int64_t> dout;
        int64_t* out = nbr;
        const bool dev_out = is_device_ptr(nbr);
        if (!dev_out) {
            VRT_TRY(dout.alloc((size_t)n * ld));
            out = dout.p;
        }
        k_vt_emit<<<(unsigned)(((size_t)n * ld + bs - 1) / bs), bs>>>(n, ld, rows.p, out);
        VRT_CUDA(cudaGetLastError());
        st.kernels += 1;
        if (!dev_out) VRT_CUDA(cudaMemcpy(nbr, dout.p, sizeof(int64_t) * (size_t)n * ld, cudaMemcpyDeviceToHost));
    }
    VRT_CUDA(cudaDeviceSynchronize());
    g_last_stats = st;
    return VRT_OK;
}

extern "C" int vrt_nearest_site(int64_t n, const double* positions, const double bounds[6], int64_t m, const double* points,
                                int64_t* idx, double* dist) {
    if (n <= 0 || !positions || !bounds || m <= 0 || !points || !idx || n >= ((int64_t)1 << 31)) {
        set_error("vrt_nearest_site: bad arguments");
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("vrt_nearest_site: no CUDA device (this library has no CPU path)");
        return VRT_E_CUDA;
    }
    SearchGrid SG;
    VRT_TRY(SG.build("vrt_nearest_site", n, positions, bounds, 2.0));
    DevBuf<double> dq, dd;
    DevBuf<int64_t> di;
    const double* q = points;
    if (!is_device_ptr(points)) {
        VRT_TRY(dq.alloc((size_t)3 * m));
        VRT_CUDA(cudaMemcpy(dq.p, points, sizeof(double) * 3 * (size_t)m, cudaMemcpyHostToDevice));
        q = dq.p;
    }
    int64_t* oi = idx;
    double* od = dist;
    const bool dev_i = is_device_ptr(idx), dev_d = dist && is_device_ptr(dist);
    if (!dev_i) { VRT_TRY(di.alloc(m)); oi = di.p; }
    if (dist && !dev_d) { VRT_TRY(dd.alloc(m)); od = dd.p; }
    k_vt_nearest<<<(unsigned)((m + 127) / 128), 128>>>(m, SG.G, q, oi, od);
    VRT_CUDA(cudaGetLastError());
    if (!dev_i) VRT_CUDA(cudaMemcpy(idx, di.p, sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToHost));
    if (dist && !dev_d) VRT_CUDA(cudaMemcpy(dist, dd.p, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

extern "C" int vrt_nearest_sites(int64_t n, const double* positions, const double bounds[6], int64_t m, const double* points, int32_t k,
                                 int64_t* idx, double* dist) {
    if (n <= 0 || !positions || !bounds || m <= 0 || !points || !idx || n >= ((int64_t)1 << 31) || k < 1 || k > VC_KMAX) {
        set_error("vrt_nearest_sites: bad arguments (1 <= k <= %d)", VC_KMAX);
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("vrt_nearest_sites: no CUDA device (this library has no CPU path)");
        return VRT_E_CUDA;
    }
    SearchGrid SG;
    VRT_TRY(SG.build("vrt_nearest_sites", n, positions, bounds, 2.0));
    DevBuf<double> dq, dd;
    DevBuf<int64_t> di;
    const double* q = points;
    if (!is_device_ptr(points)) {
        VRT_TRY(dq.alloc((size_t)3 * m));
        VRT_CUDA(cudaMemcpy(dq.p, points, sizeof(double) * 3 * (size_t)m, cudaMemcpyHostToDevice));
        q = dq.p;
    }
    int64_t* oi = idx;
    double* od = dist;
    const bool dev_i = is_device_ptr(idx), dev_d = dist && is_device_ptr(dist);
    if (!dev_i) { VRT_TRY(di.alloc((size_t)m * k)); oi = di.p; }
    if (dist && !dev_d) { VRT_TRY(dd.alloc((size_t)m * k)); od = dd.p; }
    k_vt_nearest_k<<<(unsigned)((m + 127) / 128), 128>>>(n, m, k, SG.G, q, oi, od);
    VRT_CUDA(cudaGetLastError());
    if (!dev_i) VRT_CUDA(cudaMemcpy(idx, di.p, sizeof(int64_t) * (size_t)m * k, cudaMemcpyDeviceToHost));
    if (dist && !dev_d) VRT_CUDA(cudaMemcpy(dist, dd.p, sizeof(double) * (size_t)m * k, cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}
