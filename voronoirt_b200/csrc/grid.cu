// grid.cu — K1 (BFS layers from the z-walls + stable sort) and K2 (per-direction upwind stencil).
//
// Replaces read_cell's grid analysis (reference src/voronoi_utils.jl:65-84): _sort_by_layer_up/_down
// (:93-174), sortperm + reduce_layers (:72-79, :253-269), calc_Delaunay_lines (:186-245) and
// smallest_angle (:360-396).  All integer results are bit-exact with the reference semantics; the
// floating-point comparisons that select the stencil use explicitly rounded (_rn) operations so that
// no FMA contraction can flip a near-tie with respect to the CPU oracle.
#include <cub/device/device_radix_sort.cuh>
#include <limits.h>
#include <math.h>
#include "vrt_internal.h"

namespace vrt {

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

// ---------------------------------------------------------------- K1: layers
// first loop of _sort_by_layer_* (voronoi_utils.jl:97-106 / :141-150); also validates the ids
__global__ void k_wall_layers(const int64_t* __restrict__ nbr, int64_t n, int64_t ld, int64_t wall,
                              int32_t* __restrict__ layers, int* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t cnt = nbr[i];
    if (cnt < 0 || cnt >= ld) {
        atomicExch(bad, 1);
        cnt = 0;
    }
    int32_t l = 0;
    for (int64_t j = 1; j <= cnt; j++) {
        int64_t v = nbr[i + n * j];
        if (v == wall) l = 1;
        if (v > n) atomicExch(bad, 2);
    }
    layers[i] = l;
}

// one pass of the `while true` loop (voronoi_utils.jl:109-127).  In-place is safe: a pass only writes
// lower+1 and only tests == lower, exactly like the reference's in-place scan.
__global__ void k_bfs_pass(const int64_t* __restrict__ nbr, int64_t n, int32_t lower, int32_t* layers,
                           unsigned long long* __restrict__ counters) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int assigned = 0, remaining = 0;
    if (i < n && layers[i] == 0) {
        int64_t cnt = nbr[i];
        bool found = false;
        for (int64_t j = 1; j <= cnt; j++) {
            int64_t nb = nbr[i + n * j];
            if (nb > 0 && layers[nb - 1] == lower) {
                found = true;
                break;
            }
        }
        if (found) {
            layers[i] = lower + 1;
            assigned = 1;
        } else
            remaining = 1;
    }
    int a = __syncthreads_count(assigned);
    int r = __syncthreads_count(remaining);
    if (threadIdx.x == 0) {
        if (a) atomicAdd(&counters[0], (unsigned long long)a);
        if (r) atomicAdd(&counters[1], (unsigned long long)r);
    }
}

__global__ void k_iota(int32_t* a, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int32_t)i;
}

// reduce_layers (voronoi_utils.jl:253-269): off[l-1] = first 1-based index with layer l
__global__ void k_layer_offsets(const int32_t* __restrict__ sorted, int64_t n, int64_t* __restrict__ off) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) off[0] = 1;
    else if (sorted[i] != sorted[i - 1]) off[sorted[i] - 1] = i + 1;
}

static int bfs_layers(const int64_t* d_nbr, int64_t n, int64_t ld, int64_t wall, int32_t* d_layers, int* d_bad,
                      unsigned long long* d_counters) {
    const int bs = 256;
    k_wall_layers<<<nblocks(n, bs), bs>>>(d_nbr, n, ld, wall, d_layers, d_bad);
    VRT_CUDA(cudaGetLastError());
    int bad = 0;
    VRT_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) {
        set_error(bad == 1 ? "neighbour count out of range" : "neighbour id larger than n");
        return VRT_E_INVALID;
    }
    for (int32_t lower = 1;; lower++) {
        VRT_CUDA(cudaMemsetAsync(d_counters, 0, 2 * sizeof(unsigned long long)));
        k_bfs_pass<<<nblocks(n, bs), bs>>>(d_nbr, n, lower, d_layers, d_counters);
        VRT_CUDA(cudaGetLastError());
        unsigned long long c[2];
        VRT_CUDA(cudaMemcpy(c, d_counters, sizeof(c), cudaMemcpyDeviceToHost));
        if (c[1] == 0) break;
        if (c[0] == 0) {
            set_error("%llu sites are not connected to the %s wall through the neighbour graph (the reference would loop forever)",
                      c[1], wall == -5 ? "z_min" : "z_max");
            return VRT_E_GRID;
        }
    }
    return VRT_OK;
}

// stable argsort by layer (sortperm, voronoi_utils.jl:72,77): LSD radix sort is stable
static int sort_by_layer(const int32_t* d_layers, int64_t n, int32_t* d_sorted, int32_t* d_perm) {
    DevBuf<int32_t> iota;
    VRT_TRY(iota.alloc(n));
    k_iota<<<nblocks(n, 256), 256>>>(iota.p, n);
    size_t tmp_bytes = 0;
    VRT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_layers, d_sorted, iota.p, d_perm, (int)n, 0, 32));
    DevBuf<char> tmp;
    VRT_TRY(tmp.alloc(tmp_bytes));
    VRT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, d_layers, d_sorted, iota.p, d_perm, (int)n, 0, 32));
    return VRT_OK;
}

// ---------------------------------------------------------------- internal (perm_up) ordering
__global__ void k_invert_perm(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) inv[perm[r]] = (int32_t)r;
}

__global__ void k_build_internal(const double* __restrict__ pos_host, const int64_t* __restrict__ nbr, int64_t n, int64_t max_nb,
                                 const int32_t* __restrict__ site_of, const int32_t* __restrict__ rank_of,
                                 const int32_t* __restrict__ layers_dn_host,
                                 double* __restrict__ pos, int32_t* __restrict__ nbr_int, int32_t* __restrict__ nnb,
                                 int32_t* __restrict__ layer_dn) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int64_t s = site_of[c];
    pos[3 * c + 0] = pos_host[3 * s + 0];
    pos[3 * c + 1] = pos_host[3 * s + 1];
    pos[3 * c + 2] = pos_host[3 * s + 2];
    int64_t cnt = nbr[s];
    nnb[c] = (int32_t)cnt;
    layer_dn[c] = layers_dn_host[s];
    for (int64_t j = 0; j < max_nb; j++) {
        int32_t v = INT_MIN;
        if (j < cnt) {
            int64_t nb = nbr[s + n * (j + 1)];
            v = nb > 0 ? rank_of[nb - 1] : (nb < 0 ? (int32_t)nb : -1);
        }
        nbr_int[j * n + c] = v;
    }
}

__global__ void k_rank_dn(const int32_t* __restrict__ perm_dn_host, const int32_t* __restrict__ rank_of, int64_t n,
                          int32_t* __restrict__ rank_dn, int32_t* __restrict__ perm_dn_int) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    int32_t c = rank_of[perm_dn_host[q]];
    rank_dn[c] = (int32_t)q;
    perm_dn_int[q] = c;
}

int grid_build(vrt_grid* g, const double* positions, const int64_t* nbr, int64_t ld) {
    const int64_t n = g->n;
    const int bs = 256;
    if (n >= (int64_t)ROW_MASK) {
        set_error("n=%lld exceeds the 2^29-1 rows the sweep program can address", (long long)n);
        return VRT_E_INVALID;
    }
    DevBuf<int64_t> d_nbr;
    DevBuf<double> d_pos_host;
    VRT_TRY(d_nbr.alloc((size_t)n * ld));
    VRT_TRY(d_pos_host.alloc((size_t)3 * n));
    VRT_TRY(copy_in(d_nbr.p, nbr, sizeof(int64_t) * (size_t)n * ld));
    VRT_TRY(copy_in(d_pos_host.p, positions, sizeof(double) * 3 * (size_t)n));

    DevBuf<int32_t> lay_up, lay_dn, sorted, perm_up, perm_dn;
    DevBuf<int> bad;
    DevBuf<unsigned long long> counters;
    VRT_TRY(lay_up.alloc(n)); VRT_TRY(lay_dn.alloc(n)); VRT_TRY(sorted.alloc(n));
    VRT_TRY(perm_up.alloc(n)); VRT_TRY(perm_dn.alloc(n));
    VRT_TRY(bad.alloc(1)); VRT_TRY(counters.alloc(2));
    VRT_CUDA(cudaMemset(bad.p, 0, sizeof(int)));

    // max neighbours (read_cell trims the matrix to it, voronoi_utils.jl:65-70)
    {
        std::vector<int64_t> cnt((size_t)n);
        VRT_CUDA(cudaMemcpy(cnt.data(), d_nbr.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost));
        int64_t m = 0;
        for (int64_t i = 0; i < n; i++) m = cnt[i] > m ? cnt[i] : m;
        if (m + 1 > ld) {
            set_error("neighbour count %lld exceeds ld-1=%lld", (long long)m, (long long)ld - 1);
            return VRT_E_INVALID;
        }
        g->max_nb = m;
    }

    // ---- up
    VRT_TRY(bfs_layers(d_nbr.p, n, ld, -5, lay_up.p, bad.p, counters.p));
    VRT_TRY(sort_by_layer(lay_up.p, n, sorted.p, perm_up.p));
    VRT_TRY(g->layer_up.alloc(n));
    VRT_CUDA(cudaMemcpy(g->layer_up.p, sorted.p, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice));
    int32_t L = 0;
    VRT_CUDA(cudaMemcpy(&L, sorted.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    g->L_up = L;
    {
        DevBuf<int64_t> off;
        VRT_TRY(off.alloc(L + 1));
        k_layer_offsets<<<nblocks(n, bs), bs>>>(sorted.p, n, off.p);
        g->off_up.resize(L + 1);
        VRT_CUDA(cudaMemcpy(g->off_up.data(), off.p, sizeof(int64_t) * (L + 1), cudaMemcpyDeviceToHost));
        g->off_up[L] = n;  // reduced_layers[end] = length(layers) (voronoi_utils.jl:266)
    }
    // ---- down
    VRT_TRY(bfs_layers(d_nbr.p, n, ld, -6, lay_dn.p, bad.p, counters.p));
    VRT_TRY(sort_by_layer(lay_dn.p, n, sorted.p, perm_dn.p));
    VRT_CUDA(cudaMemcpy(&L, sorted.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    g->L_down = L;
    {
        DevBuf<int64_t> off;
        VRT_TRY(off.alloc(L + 1));
        k_layer_offsets<<<nblocks(n, bs), bs>>>(sorted.p, n, off.p);
        g->off_down.resize(L + 1);
        VRT_CUDA(cudaMemcpy(g->off_down.data(), off.p, sizeof(int64_t) * (L + 1), cudaMemcpyDeviceToHost));
        g->off_down[L] = n;
    }
    // host copies of the permutations (1-based, ABI)
    {
        std::vector<int32_t> tmp((size_t)n);
        g->perm_up.resize(n);
        g->perm_down.resize(n);
        VRT_CUDA(cudaMemcpy(tmp.data(), perm_up.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; i++) g->perm_up[i] = (int64_t)tmp[i] + 1;
        VRT_CUDA(cudaMemcpy(tmp.data(), perm_dn.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; i++) g->perm_down[i] = (int64_t)tmp[i] + 1;
    }
    // ---- internal ordering = perm_up rank
    VRT_TRY(g->site_of.alloc(n)); VRT_TRY(g->rank_of.alloc(n));
    VRT_CUDA(cudaMemcpy(g->site_of.p, perm_up.p, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice));
    k_invert_perm<<<nblocks(n, bs), bs>>>(perm_up.p, n, g->rank_of.p);
    VRT_TRY(g->pos.alloc((size_t)3 * n));
    VRT_TRY(g->nbr.alloc((size_t)(g->max_nb > 0 ? g->max_nb : 1) * n));
    VRT_TRY(g->nnb.alloc(n)); VRT_TRY(g->layer_dn.alloc(n));
    VRT_TRY(g->rank_dn.alloc(n)); VRT_TRY(g->perm_dn_int.alloc(n));
    k_build_internal<<<nblocks(n, bs), bs>>>(d_pos_host.p, d_nbr.p, n, g->max_nb, g->site_of.p, g->rank_of.p, lay_dn.p,
                                             g->pos.p, g->nbr.p, g->nnb.p, g->layer_dn.p);
    k_rank_dn<<<nblocks(n, bs), bs>>>(perm_dn.p, g->rank_of.p, n, g->rank_dn.p, g->perm_dn_int.p);
    VRT_CUDA(cudaGetLastError());
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

// ---------------------------------------------------------------- K2: stencil
// calc_Delaunay_lines (voronoi_utils.jl:186-245) for one neighbour, including the mirror quirk (Q3)
__device__ __forceinline__ void delaunay_line(const double* __restrict__ P, const double* __restrict__ Pn,
                                              double x_min, double x_max, double y_min, double y_max, double* line) {
    double x_r_r = x_max - P[1];
    double x_r_l = P[1] - x_min;
    double y_r_r = y_max - P[2];
    double y_r_l = P[2] - y_min;
    double pn0 = Pn[0], pn1 = Pn[1], pn2 = Pn[2];
    double x_i_r = fabs(x_max - pn1);
    double x_i_l = fabs(pn1 - x_min);
    if (x_r_r + x_i_l < P[1] - pn1)
        pn1 = x_max + pn1 - x_min;
    else if (x_r_l + x_i_r < pn1 - P[1])
        pn1 = x_min + x_max - pn1;
    double y_i_r = fabs(y_max - pn2);
    double y_i_l = fabs(pn2 - y_min);
    if (y_r_r + y_i_l < P[2] - pn2)
        pn2 = y_max + pn2 - y_min;
    else if (y_r_l + y_i_r < pn2 - P[2])
        pn2 = y_min + y_max - pn2;
    double d0 = pn0 - P[0], d1 = pn1 - P[1], d2 = pn2 - P[2];
    double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
    line[0] = __ddiv_rn(d0, nrm);
    line[1] = __ddiv_rn(d1, nrm);
    line[2] = __ddiv_rn(d2, nrm);
}

// smallest_angle(n::Int, ...) (voronoi_utils.jl:360-396) + dot_weights (irregular_ray_tracing.jl:51) + r (:66)
__global__ void k_stencil(const double* __restrict__ pos, const int32_t* __restrict__ nbr, const int32_t* __restrict__ nnb,
                          int64_t n, double k0, double k1, double k2, double p,
                          double x_min, double x_max, double y_min, double y_max,
                          int32_t* __restrict__ up, double* __restrict__ dots_out, double* __restrict__ w, double* __restrict__ r) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    double P[3] = {pos[3 * c], pos[3 * c + 1], pos[3 * c + 2]};
    double dots[2] = {-1.0, -1.0};
    int32_t ind[2] = {-1, -1};
    int32_t cnt = nnb[c];
    for (int32_t j = 0; j < cnt; j++) {
        int32_t nb = nbr[(int64_t)j * n + c];
        if (nb >= 0) {
            double Pn[3] = {pos[3 * (int64_t)nb], pos[3 * (int64_t)nb + 1], pos[3 * (int64_t)nb + 2]};
            double l[3];
            delaunay_line(P, Pn, x_min, x_max, y_min, y_max, l);
            double d = __dadd_rn(__dadd_rn(__dmul_rn(k0, l[0]), __dmul_rn(k1, l[1])), __dmul_rn(k2, l[2]));
            if (d > dots[1]) {
                if (d > dots[0]) {
                    dots[0] = d;
                    ind[0] = nb;
                } else {
                    dots[1] = d;
                    ind[1] = nb;
                }
            }
        }
    }
    if (dots[1] <= 0) {
        dots[1] = 0;
        ind[1] = ind[0];
    }
    double p1 = pow(dots[0], p), p2 = pow(dots[1], p);
    double sum = __dadd_rn(p1, p2);
    for (int m = 0; m < 2; m++) {
        up[2 * c + m] = ind[m];
        dots_out[2 * c + m] = dots[m];
        w[2 * c + m] = __ddiv_rn(m == 0 ? p1 : p2, sum);
        double rr = 0.0;
        if (ind[m] >= 0) {
            const double* Q = pos + 3 * (int64_t)ind[m];
            double e0 = P[0] - Q[0], e1 = P[1] - Q[1], e2 = P[2] - Q[2];
            rr = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1)), __dmul_rn(e2, e2)));
        }
        r[2 * c + m] = rr;
    }
}

int grid_stencil(vrt_grid* g, const double k[3], double p, Stencil* st) {
    const int64_t n = g->n;
    VRT_TRY(st->up.ensure(2 * n)); VRT_TRY(st->dots.ensure(2 * n));
    VRT_TRY(st->w.ensure(2 * n)); VRT_TRY(st->r.ensure(2 * n));
    k_stencil<<<nblocks(n, 128), 128>>>(g->pos.p, g->nbr.p, g->nnb.p, n, k[0], k[1], k[2], p,
                                         g->bounds[2], g->bounds[3], g->bounds[4], g->bounds[5],
                                         st->up.p, st->dots.p, st->w.p, st->r.p);
    VRT_CUDA(cudaGetLastError());
    return VRT_OK;
}

// Delaunay_lines in ABI layout 3 x max_nb x n (host site order); wall slots 0
__global__ void k_lines_out(const double* __restrict__ pos, const int32_t* __restrict__ nbr, const int32_t* __restrict__ nnb,
                            const int32_t* __restrict__ site_of, int64_t n, int64_t max_nb,
                            double x_min, double x_max, double y_min, double y_max, double* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * max_nb) return;
    int64_t c = t % n, j = t / n;
    int64_t s = site_of[c];
    double l[3] = {0, 0, 0};
    if (j < nnb[c]) {
        int32_t nb = nbr[j * n + c];
        if (nb >= 0) delaunay_line(pos + 3 * c, pos + 3 * (int64_t)nb, x_min, x_max, y_min, y_max, l);
    }
    double* o = out + 3 * (j + max_nb * s);
    o[0] = l[0]; o[1] = l[1]; o[2] = l[2];
}

// internal -> host order conversions for the introspection calls
__global__ void k_stencil_out(const int32_t* __restrict__ up, const double* __restrict__ a, const double* __restrict__ b,
                              const double* __restrict__ c3, const int32_t* __restrict__ site_of, int64_t n,
                              int64_t* __restrict__ up_out, double* __restrict__ a_out, double* __restrict__ b_out, double* __restrict__ c_out) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int64_t s = site_of[c];
    for (int m = 0; m < 2; m++) {
        if (up_out) {
            int32_t u = up[2 * c + m];
            up_out[2 * s + m] = u >= 0 ? (int64_t)site_of[u] + 1 : 0;
        }
        if (a_out) a_out[2 * s + m] = a[2 * c + m];
        if (b_out) b_out[2 * s + m] = b[2 * c + m];
        if (c_out) c_out[2 * s + m] = c3[2 * c + m];
    }
}

__global__ void k_scatter_i32(const int32_t* __restrict__ src, const int32_t* __restrict__ site_of, int64_t n, int width, int32_t* __restrict__ dst) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int64_t s = site_of[c];
    for (int m = 0; m < width; m++) dst[width * s + m] = src[width * c + m];
}

}  // namespace vrt

using namespace vrt;

vrt_grid::~vrt_grid() {
    for (auto* s : cache) delete s;
    vrt::sweep_timers_free(sweep_timers);
}

extern "C" {

int vrt_grid_create(int64_t n, const double* positions, const int64_t* nbr, int64_t ld, const double bounds[6], vrt_grid** out) {
    if (!out) return VRT_E_INVALID;
    *out = nullptr;
    if (n <= 0 || !positions || !nbr || ld < 2 || !bounds) {
        set_error("vrt_grid_create: bad arguments");
        return VRT_E_INVALID;
    }
    int dev = 0;
    VRT_CUDA(cudaGetDevice(&dev));
    vrt_grid* g = new vrt_grid();
    g->n = n;
    g->ld = ld;
    g->device = dev;
    for (int i = 0; i < 6; i++) g->bounds[i] = bounds[i];
    int rc = grid_build(g, positions, nbr, ld);
    if (rc != VRT_OK) {
        delete g;
        return rc;
    }
    *out = g;
    return VRT_OK;
}

void vrt_grid_destroy(vrt_grid* g) { delete g; }

int vrt_grid_size(const vrt_grid* g, int64_t* n, int64_t* max_nb) {
    if (!g) return VRT_E_INVALID;
    if (n) *n = g->n;
    if (max_nb) *max_nb = g->max_nb;
    return VRT_OK;
}

int vrt_grid_num_layers(const vrt_grid* g, int32_t down, int64_t* L) {
    if (g && g->regular) {
        set_error("vrt_grid_num_layers: not defined on a regular grid handle");
        return VRT_E_STATE;
    }
    if (!g || !L) return VRT_E_INVALID;
    *L = down ? g->L_down : g->L_up;
    return VRT_OK;
}

int vrt_grid_get_layers(const vrt_grid* g, int32_t down, int64_t* perm, int64_t* offsets) {
    if (g && g->regular) {
        set_error("vrt_grid_get_layers: not defined on a regular grid handle");
        return VRT_E_STATE;
    }
    if (!g) return VRT_E_INVALID;
    const auto& p = down ? g->perm_down : g->perm_up;
    const auto& o = down ? g->off_down : g->off_up;
    if (perm) VRT_CUDA(cudaMemcpy(perm, p.data(), sizeof(int64_t) * p.size(), is_device_ptr(perm) ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
    if (offsets) VRT_CUDA(cudaMemcpy(offsets, o.data(), sizeof(int64_t) * o.size(), is_device_ptr(offsets) ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
    return VRT_OK;
}

int vrt_grid_get_delaunay_lines(const vrt_grid* g, double* lines) {
    if (g && g->regular) {
        set_error("vrt_grid_get_delaunay_lines: not defined on a regular grid handle");
        return VRT_E_STATE;
    }
    if (!g || !lines) return VRT_E_INVALID;
    size_t cnt = (size_t)3 * g->max_nb * g->n;
    DevBuf<double> tmp;
    double* d = lines;
    if (!is_device_ptr(lines)) {
        VRT_TRY(tmp.alloc(cnt));
        d = tmp.p;
    }
    int64_t t = g->n * g->max_nb;
    k_lines_out<<<nblocks(t, 256), 256>>>(g->pos.p, g->nbr.p, g->nnb.p, g->site_of.p, g->n, g->max_nb,
                                          g->bounds[2], g->bounds[3], g->bounds[4], g->bounds[5], d);
    VRT_CUDA(cudaGetLastError());
    if (d != lines) VRT_TRY(copy_out(lines, d, cnt * sizeof(double)));
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

int vrt_grid_get_stencil(vrt_grid* g, const double k[3], double p, int64_t* upwind, double* dots, double* weights, double* r) {
    if (g && g->regular) {
        set_error("vrt_grid_get_stencil: not defined on a regular grid handle");
        return VRT_E_STATE;
    }
    if (!g || !k) return VRT_E_INVALID;
    Stencil st;
    VRT_TRY(grid_stencil(g, k, p, &st));
    const int64_t n = g->n;
    DevBuf<int64_t> d_up;
    DevBuf<double> d_a, d_b, d_c;
    int64_t* pu = upwind;
    double *pa = dots, *pb = weights, *pc = r;
    if (upwind && !is_device_ptr(upwind)) { VRT_TRY(d_up.alloc(2 * n)); pu = d_up.p; }
    if (dots && !is_device_ptr(dots)) { VRT_TRY(d_a.alloc(2 * n)); pa = d_a.p; }
    if (weights && !is_device_ptr(weights)) { VRT_TRY(d_b.alloc(2 * n)); pb = d_b.p; }
    if (r && !is_device_ptr(r)) { VRT_TRY(d_c.alloc(2 * n)); pc = d_c.p; }
    k_stencil_out<<<nblocks(n, 256), 256>>>(st.up.p, st.dots.p, st.w.p, st.r.p, g->site_of.p, n, pu, pa, pb, pc);
    VRT_CUDA(cudaGetLastError());
    if (upwind && pu != upwind) VRT_TRY(copy_out(upwind, pu, sizeof(int64_t) * 2 * n));
    if (dots && pa != dots) VRT_TRY(copy_out(dots, pa, sizeof(double) * 2 * n));
    if (weights && pb != weights) VRT_TRY(copy_out(weights, pb, sizeof(double) * 2 * n));
    if (r && pc != r) VRT_TRY(copy_out(r, pc, sizeof(double) * 2 * n));
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

int vrt_grid_release_schedules(vrt_grid* g) {
    if (!g) return VRT_E_INVALID;
    schedule_cache_clear(g);
    return VRT_OK;
}

int vrt_grid_get_schedule(vrt_grid* g, const double k[3], int32_t down, int32_t n_sweeps, int32_t prune,
                          int32_t* cls, int32_t* sublevel, int32_t* stab, int64_t* n_steps, int64_t* n_visits) {
    if (g && g->regular) {
        set_error("vrt_grid_get_schedule: not defined on a regular grid handle");
        return VRT_E_STATE;
    }
    if (!g || !k) return VRT_E_INVALID;
    // built on request with the introspection arrays kept (the cached schedules of the solvers drop them)
    DirSchedule* sch = nullptr;
    VRT_TRY(schedule_build(g, k, down, n_sweeps, 7.0, prune, 1, order_config(g->n, 1), true, &sch));
    struct Guard { DirSchedule* s; ~Guard() { delete s; } } guard{sch};
    const int64_t n = g->n;
    DevBuf<int32_t> tmp;
    VRT_TRY(tmp.alloc(2 * n));
    if (cls) {
        k_scatter_i32<<<nblocks(n, 256), 256>>>(sch->cls.p, g->site_of.p, n, 2, tmp.p);
        VRT_TRY(copy_out(cls, tmp.p, sizeof(int32_t) * 2 * n));
        VRT_CUDA(cudaDeviceSynchronize());
    }
    if (sublevel) {
        k_scatter_i32<<<nblocks(n, 256), 256>>>(sch->sublevel.p, g->site_of.p, n, 1, tmp.p);
        VRT_TRY(copy_out(sublevel, tmp.p, sizeof(int32_t) * n));
        VRT_CUDA(cudaDeviceSynchronize());
    }
    if (stab) {
        k_scatter_i32<<<nblocks(n, 256), 256>>>(sch->stab.p, g->site_of.p, n, 1, tmp.p);
        VRT_TRY(copy_out(stab, tmp.p, sizeof(int32_t) * n));
        VRT_CUDA(cudaDeviceSynchronize());
    }
    if (n_steps) *n_steps = (int64_t)sch->step_off.size() - 1;
    if (n_visits) *n_visits = sch->n_visits;
    return VRT_OK;
}

}  // extern "C"
