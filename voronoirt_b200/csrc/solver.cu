// solver.cu — the Λ-iteration engine around the sweep: K4 opacity, boundary values, J reduction,
// K6 source update + criterion, K7 radiative rates, K8 statistical equilibrium, and the ABI entry points
// that replace Delaunay_upII/downII, J_λ_voronoi, calculate_R, get_revised_populations and Λ_voronoi
// (reference src/irregular_ray_tracing.jl, src/lambda_iteration.jl, src/lambda_continuum.jl, src/rates.jl,
// src/populations.jl).  Everything is Float64.  All nlam x n arrays are [cell][λ] in internal cell order.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <array>
#include <chrono>
#include <cub/device/device_reduce.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include "physics.cuh"
#include "vrt_internal.h"

using namespace vrt;

namespace vrt {

static inline int nblocks(int64_t n, int bs) { return (int)std::min<int64_t>((n + bs - 1) / bs, 1 << 30); }

// ---------------------------------------------------------------- layout helpers
// gather: dst[c][l] = src[map[c]][l]   (host order -> internal, map = site_of)
// scatter: dst[map[c]][l] = src[c][l]  (internal -> host order)
__global__ void k_permute_rows(const double* __restrict__ src, double* __restrict__ dst, const int32_t* __restrict__ map,
                               int64_t n, int64_t nlam, int gather) {
    int64_t total = n * nlam;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / nlam, l = i - c * nlam;
        int64_t o = (int64_t)map[c] * nlam + l;
        if (gather) dst[i] = src[o];
        else dst[o] = src[i];
    }
}

int permute_rows(const double* src, double* dst, const int32_t* map, int64_t n, int64_t nlam, int gather, cudaStream_t st) {
    k_permute_rows<<<nblocks(n * nlam, 256), 256, 0, st>>>(src, dst, map, n, nlam, gather);
    VRT_CUDA(cudaGetLastError());
    return VRT_OK;
}

// host-or-device [n][w] array in host site order -> device internal order
static int upload_rows(const vrt_grid* g, const double* src, double* dst_int, int64_t w, DevBuf<double>& stage) {
    const int64_t n = g->n;
    const double* d = src;
    if (!is_device_ptr(src)) {
        VRT_TRY(stage.ensure((size_t)n * w));
        VRT_CUDA(cudaMemcpy(stage.p, src, sizeof(double) * (size_t)n * w, cudaMemcpyHostToDevice));
        d = stage.p;
    }
    return permute_rows(d, dst_int, g->site_of.p, n, w, 1, 0);
}

static int download_rows(const vrt_grid* g, const double* src_int, double* dst, int64_t w, DevBuf<double>& stage) {
    const int64_t n = g->n;
    if (is_device_ptr(dst)) return permute_rows(src_int, dst, g->site_of.p, n, w, 0, 0);
    VRT_TRY(stage.ensure((size_t)n * w));
    VRT_TRY(permute_rows(src_int, stage.p, g->site_of.p, n, w, 0, 0));
    VRT_CUDA(cudaMemcpy(dst, stage.p, sizeof(double) * (size_t)n * w, cudaMemcpyDeviceToHost));
    return VRT_OK;
}

// column-major [n x ncol] host-order matrix (e.g. populations n x 3) -> internal SoA [ncol][n]
__global__ void k_gather_cols(const double* __restrict__ src, double* __restrict__ dst, const int32_t* __restrict__ site_of,
                              int64_t n, int ncol) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int64_t s = site_of[c];
    for (int j = 0; j < ncol; j++) dst[(int64_t)j * n + c] = src[(int64_t)j * n + s];
}
__global__ void k_scatter_cols(const double* __restrict__ src, double* __restrict__ dst, const int32_t* __restrict__ site_of,
                               int64_t n, int ncol) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int64_t s = site_of[c];
    for (int j = 0; j < ncol; j++) dst[(int64_t)j * n + s] = src[(int64_t)j * n + c];
}

// ---------------------------------------------------------------- boundary values
// up directions: internal cells [0, n1) are the bottom layer (perm_up order == internal order)
__global__ void k_boundary_rows(double* __restrict__ I, int64_t nlam, const double* __restrict__ I0, const int32_t* __restrict__ cells,
                                int64_t n1) {
    int64_t total = n1 * nlam;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t q = i / nlam, l = i - q * nlam;
        int64_t c = cells ? cells[q] : q;
        I[c * nlam + l] = I0 ? I0[i] : 0.0;
    }
}

// bottom boundary of the NLTE line solve: I_0 = B_λ(λ_l, T) (lambda_iteration.jl:99-101); B0 given: continuum (lambda_continuum.jl:47)
__global__ void k_boundary_planck(double* __restrict__ I, int64_t nlam, const double* __restrict__ lam, const double* __restrict__ T,
                                  const double* __restrict__ B0, int64_t n1) {
    int64_t total = n1 * nlam;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / nlam, l = i - c * nlam;
        I[i] = B0 ? B0[c] : B_lambda(lam[l], T[c]);
    }
}

// ---------------------------------------------------------------- K4: opacity
struct LineDev {
    double lambda0, Bij, Bji, c_line;
    double c_unsold, gamma_nat, c_lin, c_quad;
};

// γ_constant (broadening.jl:63-82) with the closed forms of Transparency.jl
__global__ void k_gamma(int64_t n, LineDev L, const double* __restrict__ T, const double* __restrict__ ne,
                        const double* __restrict__ pops /* [3][n] */, double* __restrict__ gamma) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    double nHI = pops[c] + pops[n + c];
    double g = L.c_unsold * pow(T[c], 0.3) * nHI;
    g += L.gamma_nat;
    g += L.c_lin * pow(ne[c], 2.0 / 3.0);
    g += L.c_quad * pow(T[c], 1.0 / 6.0) * ne[c];
    gamma[c] = g;
}

struct OpacityDirs {
    double k[MAX_DIRS][3];
    double* alpha[MAX_DIRS];
    int nd;
};

// compute_voigt_profile (line.jl:121-137) + αline_λ (line.jl:219-225) + α_cont, for every direction of the batch.
// A CTA owns a tile of OP_TC cells x all lc wavelengths.  The math runs with the 32 lanes of a warp on 32 DIFFERENT
// CELLS at the SAME wavelength: the Humlíček region (|v| + a) is then almost warp-uniform, whereas lanes over
// wavelengths (core .. far wing) would serialise all four regions.  The tile is transposed through shared memory
// so that the α rows go out as coalesced fp64 row writes.
// Per evaluation the only division left is the one inside the Faddeeva approximation: the damping parameter a(cell, λ) does
// not depend on the direction and is kept in a second shared tile for all directions of the batch; 1/ΔD and the factor
// hc/(4πλ0)·(n1 B12 − n2 B21)/(√π ΔD) are formed once per cell.  (v = Δλ·(1/ΔD) instead of Δλ/ΔD moves v by at most one ulp.)
constexpr int OP_TC = 32;
constexpr int OP_THREADS = 256;
__global__ void __launch_bounds__(OP_THREADS) k_opacity(int64_t n, int64_t lc, const double* __restrict__ lam /* chunk */, LineDev L,
                                                        const OpacityDirs D, const double* __restrict__ gamma,
                                                        const double* __restrict__ dD, const double* __restrict__ vz,
                                                        const double* __restrict__ vx, const double* __restrict__ vy,
                                                        const double* __restrict__ pops, const double* __restrict__ alpha_cont) {
    extern __shared__ double tile[];                 // [OP_TC][ldt] α of one direction, then [OP_TC][ldt] damping a
    const int ldt = (int)lc | 1;                     // odd row stride: conflict-light transposed writes
    double* atile = tile + OP_TC * ldt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = OP_THREADS / 32;
    for (int64_t c0 = (int64_t)blockIdx.x * OP_TC; c0 < n; c0 += (int64_t)gridDim.x * OP_TC) {
        const int64_t c = c0 + lane;
        const bool ok = c < n;
        const int64_t cc = ok ? c : n - 1;
        const double dd = dD[cc], g = gamma[cc];
        const double rdd = 1.0 / dd;
        const double scale = L.c_line * (pops[cc] * L.Bij - pops[n + cc] * L.Bji) / (sqrt(PI) * (dd * 1e-9));
        const double ac = alpha_cont[cc];
        const double v0 = vz[cc], v1 = vx[cc], v2 = vy[cc];
        const int ncell = (int)min((int64_t)OP_TC, n - c0);
        // each thread fills and later reads its own entries of atile (same cell, same wavelengths): no barrier needed
        for (int l = warp; l < lc; l += nwarp) atile[lane * ldt + l] = damping_param(g, lam[l], dd);
        for (int d = 0; d < D.nd; d++) {
            // line_of_sight_velocity(sites, -k) (line.jl:198-208); explicitly rounded like the oracle
            const double vlos = __dadd_rn(__dadd_rn(__dmul_rn(v0, -D.k[d][0]), __dmul_rn(v1, -D.k[d][1])), __dmul_rn(v2, -D.k[d][2]));
            const double shift = __ddiv_rn(__dmul_rn(L.lambda0, vlos), C_0);
            for (int l = warp; l < lc; l += nwarp) {
                const double v = __dmul_rn(__dadd_rn(__dadd_rn(lam[l], -L.lambda0), shift), rdd);
                tile[lane * ldt + l] = fma(humlicek_re(atile[lane * ldt + l], v), scale, ac);
            }
            __syncthreads();
            double* __restrict__ out = D.alpha[d] + c0 * lc;
            for (int r = warp; r < ncell; r += nwarp)                      // a warp writes a row: coalesced, no index division
                for (int l = lane; l < (int)lc; l += 32) out[(int64_t)r * lc + l] = tile[r * ldt + l];
            __syncthreads();
        }
    }
}

__global__ void k_damping(int64_t n, int64_t nlam, const double* __restrict__ lam, const double* __restrict__ gamma,
                          const double* __restrict__ dD, double* __restrict__ out) {
    int64_t total = n * nlam;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / nlam, l = i - c * nlam;
        out[i] = damping_param(gamma[c], lam[l], dD[c]);
    }
}

// ---------------------------------------------------------------- J = Σ_Ω w_Ω I_Ω  (lambda_iteration.jl:102,107)
struct JDirs {
    const double* I[MAX_DIRS];
    double w[MAX_DIRS];
    int nd;
};
// atomics-free and deterministic: one thread owns (cell, λ) and adds the directions in quadrature-file order
__global__ void k_J_reduce(int64_t n, int64_t lc, const JDirs D, double* __restrict__ J, int64_t ldJ, int first) {
    int64_t total = n * lc;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / lc, l = i - c * lc;
        double* j = J + c * ldJ + l;
        double acc = first ? 0.0 : *j;
        for (int d = 0; d < D.nd; d++) acc += D.w[d] * D.I[d][i];
        *j = acc;
    }
}

// ---------------------------------------------------------------- K6: source update + criterion
__device__ __forceinline__ void block_max_nan(double d, bool isn, unsigned long long* out_bits, int* out_nan) {
    __shared__ double smax[32];
    __shared__ int snan[32];
    for (int o = 16; o > 0; o >>= 1) {
        d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
        isn = isn | (bool)__shfl_xor_sync(0xffffffffu, (int)isn, o);
    }
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smax[w] = d; snan[w] = isn; }
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        d = lane < nw ? smax[lane] : 0.0;
        isn = lane < nw ? snan[lane] : 0;
        for (int o = 16; o > 0; o >>= 1) {
            d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
            isn = isn | (bool)__shfl_xor_sync(0xffffffffu, (int)isn, o);
        }
        if (lane == 0) {
            atomicMax(out_bits, (unsigned long long)__double_as_longlong(d));  // d >= 0: bit pattern is monotone
            if (isn) atomicExch(out_nan, 1);
        }
    }
}

// S_new = (1-ε)J + εB (lambda_iteration.jl:262-264 / lambda_continuum.jl:148) fused with
// criterion's max|1 - S_old/S_new| (lambda_iteration.jl:325-349 / lambda_continuum.jl:181-198, over `thick` only)
__global__ void k_source_update(int64_t n, int64_t nlam, const double* __restrict__ lam, const double* __restrict__ T,
                                const double* __restrict__ B0, const double* __restrict__ eps, const double* __restrict__ J,
                                double* __restrict__ S, int use_thick, unsigned long long* diff_bits, int* diff_nan) {
    double dmax = 0.0;
    bool isn = false;
    if (nlam < 16) {   // narrow rows (continuum: one wavelength): flat over the elements
        const int64_t total = n * nlam;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t c = i / nlam, l = i - c * nlam;
            const double e = eps[c];
            const double B = B0 ? B0[c] : B_lambda(lam[l], T[c]);
            const double s_old = S[i];
            const double s_new = (1 - e) * J[i] + e * B;
            S[i] = s_new;
            if (!use_thick || e > 1e-4) {
                const double d = fabs(1 - s_old / s_new);
                if (d != d) isn = true;
                else dmax = fmax(dmax, d);
            }
        }
        block_max_nan(dmax, isn, diff_bits, diff_nan);
        return;
    }
    // a warp per cell row (wavelengths across the lanes): no 64-bit index division per element, ε and T read once per row
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < n; c += nwarps) {
        const double e = eps[c];
        const double Tc = B0 ? 0.0 : T[c];
        const bool counts = !use_thick || e > 1e-4;
        for (int64_t l = lane; l < nlam; l += 32) {
            const int64_t i = c * nlam + l;
            const double B = B0 ? B0[c] : B_lambda(lam[l], Tc);
            const double s_old = S[i];
            const double s_new = (1 - e) * J[i] + e * B;
            S[i] = s_new;
            if (counts) {
                const double d = fabs(1 - s_old / s_new);
                if (d != d) isn = true;
                else dmax = fmax(dmax, d);
            }
        }
    }
    block_max_nan(dmax, isn, diff_bits, diff_nan);
}

// The same, fused with the reduction of J over the direction shards THROUGH PEER MEMORY (NVLink / NVSwitch): every process
// holds its partial J (the sum over its own directions) in a buffer the others have mapped (CUDA IPC); the owner of a cell
// slice reads that slice from all R buffers, adds them in rank order, keeps the sum as its J and updates S — one pass, no
// reduce-scatter in front of it (the collective would move the same bytes first and this kernel would read them again).
constexpr int MAX_PEERS = 16;
struct PeerJ {
    const double* J[MAX_PEERS];
    int R;
};
__global__ void k_source_update_peers(int64_t n, int64_t nlam, const double* __restrict__ lam, const double* __restrict__ T,
                                      const double* __restrict__ eps, const PeerJ P, int64_t off, double* __restrict__ Jown,
                                      double* __restrict__ S, unsigned long long* diff_bits, int* diff_nan) {
    int64_t total = n * nlam;
    double dmax = 0.0;
    bool isn = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / nlam, l = i - c * nlam;
        double J = 0.0;
        for (int r = 0; r < P.R; r++) J += __ldcg(P.J[r] + off + i);
        Jown[i] = J;
        double e = eps[c];
        double B = B_lambda(lam[l], T[c]);
        double s_old = S[i];
        double s_new = (1 - e) * J + e * B;
        S[i] = s_new;
        double d = fabs(1 - s_old / s_new);
        if (d != d) isn = true;
        else dmax = fmax(dmax, d);
    }
    block_max_nan(dmax, isn, diff_bits, diff_nan);
}

// criterion alone (first pass of the while loop: S_old = 0)
__global__ void k_criterion(int64_t n, int64_t nlam, const double* __restrict__ S_new, const double* __restrict__ S_old,
                            const double* __restrict__ eps, int use_thick, unsigned long long* diff_bits, int* diff_nan) {
    int64_t total = n * nlam;
    double dmax = 0.0;
    bool isn = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / nlam;
        if (use_thick && !(eps[c] > 1e-4)) continue;
        double so = S_old ? S_old[i] : 0.0;
        double d = fabs(1 - so / S_new[i]);
        if (d != d) isn = true;
        else dmax = fmax(dmax, d);
    }
    block_max_nan(dmax, isn, diff_bits, diff_nan);
}

// single-wavelength forms of the two kernels above for the regular-grid Λ-iteration (regular.cu): S, J, eps, B0 are n-vectors
int continuum_criterion(int64_t n, const double* S_new, const double* S_old, const double* eps, unsigned long long* diff_bits, int* diff_nan) {
    k_criterion<<<nblocks(n, 256), 256>>>(n, 1, S_new, S_old, eps, 1, diff_bits, diff_nan);
    VRT_CUDA(cudaGetLastError());
    return VRT_OK;
}
int continuum_source_update(int64_t n, const double* B0, const double* eps, const double* J, double* S, unsigned long long* diff_bits,
                            int* diff_nan) {
    k_source_update<<<nblocks(n, 256), 256>>>(n, 1, nullptr, nullptr, B0, eps, J, S, 1, diff_bits, diff_nan);
    VRT_CUDA(cudaGetLastError());
    return VRT_OK;
}

// bottom boundary of the regular grid: I_0 = B_λ.(λ, T[1,:,:]) (lambda_iteration.jl:38) or B_0[1,:,:] (lambda_continuum.jl:16);
// out is [nx*ny][lc], cell (iz = 0, i) of the (nz, nx, ny) arrays is nz*i
__global__ void k_regular_boundary(double* __restrict__ out, int64_t lc, const double* __restrict__ lam, const double* __restrict__ T,
                                   const double* __restrict__ B0, int64_t nz, int64_t plane) {
    int64_t total = plane * lc;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / lc, l = i - c * lc;
        out[i] = B0 ? B0[nz * c] : B_lambda(lam[l], T[nz * c]);
    }
}

__global__ void k_planck_rows(int64_t n, int64_t nlam, const double* __restrict__ lam, const double* __restrict__ T, double* __restrict__ S) {
    int64_t total = n * nlam;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / nlam, l = i - c * nlam;
        S[i] = B_lambda(lam[l], T[c]);
    }
}

// ---------------------------------------------------------------- K7: radiative rates
// calculate_R (rates.jl:154-201).  The pairwise trapezoid Σ_l (f_l + f_{l+1})(λ_{l+1}-λ_l) of Rij/Rji
// (rates.jl:264-278, :348-364, :226-240, :305-321) is regrouped into per-wavelength weights
// W_l = (λ_{l+1}-λ_l)[l not last] + (λ_l-λ_{l-1})[l not first] so that a wavelength shard can sum its
// own terms; the shards' partial sums are then all-reduced.
struct RateLam {           // per local wavelength, host-precomputed
    double lam_m;          // λ in m
    double W;              // trapezoid weight in m
    double sigma_bf;       // σic (rates.jl:422-438) for bf wavelengths, 0 for bb
    double planck;         // 2hc²/λ⁵ (bb, rates.jl:316) or 2·hc·c/λ⁵ (bf, :359), SI
    int32_t range;         // 0 = bb (1<->2), 1 = bf level 1, 2 = bf level 2
    int32_t pad;
};

// one warp per cell, lanes over the shard's wavelengths; out: Rp[6][n] = R12,R21,R13,R31,R23,R32
// n = cells handled (pointers already offset to the first of them), stride = cell stride of the SoA arrays lte / Rp
__global__ void k_rates(int64_t n, int64_t stride, int64_t nlam, const RateLam* __restrict__ rl, LineDev L, const double* __restrict__ T,
                        const double* __restrict__ dD, const double* __restrict__ gamma, const double* __restrict__ damping /* or null */,
                        const double* __restrict__ lte /* [3][n] */, const double* __restrict__ J, double* __restrict__ Rp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double hc = H_PLANCK * C_0;
    const double pref = 2 * PI / hc;
    for (int64_t c = warp; c < n; c += nwarps) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        double Tc = T[c], dd = dD[c];
        double n1 = lte[c], n2 = lte[stride + c], n3 = lte[2 * stride + c];
        double r0 = n1 / n2, r1 = n1 / n3, r2 = n2 / n3;
        double sc = hc / (4 * PI * (L.lambda0 * 1e-9)) * L.Bij;
        for (int64_t l = lane; l < nlam; l += 32) {
            RateLam q = rl[l];
            double Jsi = J[c * nlam + l] * 1e12;
            double sigma;
            if (q.range == 0) {
                double lam_nm = q.lam_m * 1e9;
                double a = damping ? damping[c * nlam + l] : damping_param(gamma[c], lam_nm, dd);
                double v = (lam_nm - L.lambda0) / dd;
                sigma = sc * voigt_profile(a, v, dd * 1e-9);
            } else
                sigma = q.sigma_bf;
            double ratio = q.range == 0 ? r0 : (q.range == 1 ? r1 : r2);
            double G = ratio * exp(-hc / (K_B * q.lam_m * Tc));
            double up = pref * (q.lam_m * sigma * Jsi) * q.W / 1000;   // the literal /1000 of Rij (Q8)
            double dn = pref * (sigma * G * q.lam_m * (q.planck + Jsi)) * q.W;
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (q.range == k) {
                    acc[2 * k] += up;
                    acc[2 * k + 1] += dn;
                }
        }
#pragma unroll
        for (int k = 0; k < 6; k++)
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == 0)
            for (int k = 0; k < 6; k++) Rp[(int64_t)k * stride + c] = acc[k];
    }
}

// ---------------------------------------------------------------- K8: statistical equilibrium
// get_revised_populations (populations.jl:191-221): P = R + C, per-site 2x2 inv(A)*b.
__device__ __forceinline__ void stat_eq_site(double P12, double P21, double P13, double P31, double P23, double P32, double NH,
                                             double& n1, double& n2, double& n3) {
    double A11 = P12 + P21;
    double A12 = P12 - P32;
    A11 += P23;
    double A22 = P13 + P31;
    double A21 = P13 - P23;
    A22 += P32;
    double b1 = NH * P12, b2 = NH * P13;
    double x1, x2;
    if (fabs(A11) >= fabs(A21)) {
        double m = A21 / A11;
        double u22 = A22 - m * A12;
        double y2 = b2 - m * b1;
        x2 = y2 / u22;
        x1 = (b1 - A12 * x2) / A11;
    } else {
        double m = A11 / A21;
        double u22 = A12 - m * A22;
        double y2 = b1 - m * b2;
        x2 = y2 / u22;
        x1 = (b2 - A22 * x2) / A21;
    }
    n2 = x1;
    n3 = x2;
    n1 = NH - (x1 + x2);
}

// internal SoA version: Rp, Cp [6][n] (12,21,13,31,23,32), pops [3][n]
__global__ void k_stateq_soa(int64_t n, int64_t stride, const double* __restrict__ Rp, const double* __restrict__ Cp,
                             const double* __restrict__ NH, double* __restrict__ pops) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    double P[6];
    for (int k = 0; k < 6; k++) P[k] = Rp[(int64_t)k * stride + c] + Cp[(int64_t)k * stride + c];
    double n1, n2, n3;
    stat_eq_site(P[0], P[1], P[2], P[3], P[4], P[5], NH[c], n1, n2, n3);
    pops[c] = n1;
    pops[stride + c] = n2;
    pops[2 * stride + c] = n3;
}

// populations of a cell slice [c0, c0+cs) <-> packed [3][cs] block (the unit of the all-gather over cell shards)
__global__ void k_pack_pops(int64_t n, int64_t c0, int64_t cs, const double* __restrict__ pops, double* __restrict__ blk) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= cs) return;
    for (int k = 0; k < 3; k++) blk[(int64_t)k * cs + i] = (c0 + i < n) ? pops[(int64_t)k * n + c0 + i] : 0.0;
}
__global__ void k_unpack_pops(int64_t n, int64_t cs, int R, const double* __restrict__ all, double* __restrict__ pops) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int64_t r = c / cs, i = c - r * cs;
    for (int k = 0; k < 3; k++) pops[(int64_t)k * n + c] = all[(r * 3 + k) * cs + i];
}

// ABI version: R, C 3 x 3 x n (column-major, [a + 3b + 9i] = M[a+1,b+1,i+1]), pops n x 3
__global__ void k_stateq_abi(int64_t n, const double* __restrict__ R, const double* __restrict__ Cm, const double* __restrict__ NH,
                             double* __restrict__ pops) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = R + 9 * i;
    const double* c = Cm + 9 * i;
#define PP(a, b) (r[((a)-1) + 3 * ((b)-1)] + c[((a)-1) + 3 * ((b)-1)])
    double n1, n2, n3;
    stat_eq_site(PP(1, 2), PP(2, 1), PP(1, 3), PP(3, 1), PP(2, 3), PP(3, 2), NH[i], n1, n2, n3);
#undef PP
    pops[i] = n1;
    pops[n + i] = n2;
    pops[2 * n + i] = n3;
}

// Rp [6][n] internal -> R 3x3xn in host site order (diagonal 0, rates.jl:196-198)
__global__ void k_R_out(int64_t n, const double* __restrict__ Rp, const int32_t* __restrict__ site_of, double* __restrict__ R) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    double* o = R + 9 * (int64_t)site_of[c];
    for (int k = 0; k < 9; k++) o[k] = 0.0;
    o[0 + 3 * 1] = Rp[0 * n + c];  // R[1,2]
    o[1 + 3 * 0] = Rp[1 * n + c];  // R[2,1]
    o[0 + 3 * 2] = Rp[2 * n + c];  // R[1,3]
    o[2 + 3 * 0] = Rp[3 * n + c];  // R[3,1]
    o[1 + 3 * 2] = Rp[4 * n + c];  // R[2,3]
    o[2 + 3 * 1] = Rp[5 * n + c];  // R[3,2]
}
// C 3x3xn host order -> Cp [6][n] internal
__global__ void k_C_in(int64_t n, const double* __restrict__ Cm, const int32_t* __restrict__ site_of, double* __restrict__ Cp) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    const double* o = Cm + 9 * (int64_t)site_of[c];
    Cp[0 * n + c] = o[0 + 3 * 1];
    Cp[1 * n + c] = o[1 + 3 * 0];
    Cp[2 * n + c] = o[0 + 3 * 2];
    Cp[3 * n + c] = o[2 + 3 * 0];
    Cp[4 * n + c] = o[1 + 3 * 2];
    Cp[5 * n + c] = o[2 + 3 * 1];
}

}  // namespace vrt

// ================================================================= solver handle
struct vrt_solver {
    vrt_grid* g = nullptr;
    int is_line = 0;
    vrt_config cfg;
    vrt_line line;
    LineDev ld;
    int64_t n = 0, nlam_total = 0, l_begin = 0, nlam = 0;  // nlam = local wavelength count
    std::vector<double> lambda;                              // all wavelengths, nm
    // quadrature (directions with θ == 90 are skipped like the reference does)
    int nd = 0;
    std::vector<double> qw;
    std::vector<int> qdown;
    std::vector<std::array<double, 3>> qk;
    std::vector<DirSchedule*> sch;
    std::vector<int> order;                 // order the directions are swept in (batches of `db` consecutive entries)
    // wavelength sub-range [lo, hi) of the local wavelengths a direction is solved on (vrt_solver_set_direction_lambda): a
    // direction shared between two processes, each taking part of its wavelengths.  lo == hi == 0: all local wavelengths.
    std::vector<std::array<int64_t, 2>> dir_lam;
    int64_t n1_up = 0, n1_dn = 0;
    // per-site device arrays, internal order
    DevBuf<double> T, ne, NH, vz, vx, vy, dD, alpha_cont, eps, B0, Cp, lte, lam_dev, gamma;
    DevBuf<RateLam> rl;
    // state
    DevBuf<double> S, J, pops, Rp, S_prev;
    bool have_state = false;
    bool dir_sharded = false;
    int cell_R = 1, cell_r = 0;             // cell shards of the post-J stages (source, rates, stat-eq)
    int64_t cs = 0, c0 = 0, c1 = 0, n_pad = 0; // cells per shard, own slice [c0, c1), padded cell count
    DevBuf<double> pops_blk, diff_pair;
    bool gamma_valid = false;
    // work buffers of the current (λ-chunk, direction-batch) plan
    int64_t lc = 0;
    int db = 0;
    std::vector<DevBuf<double>*> bufs;
    std::vector<double*> alpha_p, I_p;
    std::vector<std::array<double*, MAX_SWEEPS>> scr_p;
    DevBuf<double> stage;
    DevBuf<unsigned long long> diff_bits;
    DevBuf<int> diff_nan;
    vrt_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    void* comm = nullptr;                   // in-library NCCL communicators (comm.cu), preferred over the host hook
    cudaEvent_t comm_ev[2] = {nullptr, nullptr};
    const double* peerJ[16] = {nullptr};    // J buffers of the direction group mapped through CUDA IPC (own entry = own buffer)
    bool peers = false;
    cudaEvent_t gather_ev = nullptr;        // a deferred all-gather of S is running on the collectives' stream until this event
    bool gather_pending = false;
    bool has_exchange() const { return comm != nullptr || allreduce != nullptr; }
    // CUDA events are created once and reused (no create/destroy per iteration, nothing to leak on an early return)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    int event(cudaEvent_t* e) {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t n;
            VRT_CUDA(cudaEventCreate(&n));
            ev_pool.push_back(n);
        }
        *e = ev_pool[ev_used++];
        return VRT_OK;
    }
    ~vrt_solver() {
        for (auto* b : bufs) delete b;
        for (auto e : ev_pool) cudaEventDestroy(e);
        for (auto e : comm_ev)
            if (e) cudaEventDestroy(e);
        if (gather_ev) cudaEventDestroy(gather_ev);
        if (peers)
            for (int r = 0; r < 16; r++)
                if (peerJ[r] && r != cell_r) cudaIpcCloseMemHandle(const_cast<double*>(peerJ[r]));
        if (comm) vrt::comm_free(comm);
    }
};

namespace vrt {

static int solver_common_init(vrt_solver* s, vrt_grid* g, const vrt_quadrature* quad, const vrt_config* cfg, int64_t nlam_local) {
    s->g = g;
    s->n = g->n;
    vrt_config c;
    memset(&c, 0, sizeof(c));
    c.n_sweeps = 3;
    c.p = 7.0;
    c.prune = 1;
    if (cfg) {
        c = *cfg;
        if (c.n_sweeps <= 0) c.n_sweeps = 3;
        if (c.p == 0.0) c.p = 7.0;
    }
    s->cfg = c;
    s->cell_R = 1; s->cell_r = 0;
    if (c.cell_shard_count > 1) {
        if (c.cell_shard_rank < 0 || c.cell_shard_rank >= c.cell_shard_count) {
            set_error("solver: cell shard rank out of range");
            return VRT_E_INVALID;
        }
        s->cell_R = c.cell_shard_count;
        s->cell_r = c.cell_shard_rank;
    }
    s->cs = (g->n + s->cell_R - 1) / s->cell_R;
    s->n_pad = s->cs * s->cell_R;
    s->c0 = std::min<int64_t>(g->n, s->cs * s->cell_r);
    s->c1 = std::min<int64_t>(g->n, s->c0 + s->cs);
    const int cv = chunk_visits(c.lam_chunk > 0 ? std::min<int64_t>(c.lam_chunk, nlam_local) : nlam_local);
    if (!quad || quad->n_dirs <= 0 || !quad->weights || !quad->theta || !quad->phi) {
        set_error("solver: bad quadrature");
        return VRT_E_INVALID;
    }
    std::vector<double> w(quad->n_dirs), th(quad->n_dirs), ph(quad->n_dirs);
    VRT_CUDA(cudaMemcpy(w.data(), quad->weights, sizeof(double) * quad->n_dirs, cudaMemcpyDefault));
    VRT_CUDA(cudaMemcpy(th.data(), quad->theta, sizeof(double) * quad->n_dirs, cudaMemcpyDefault));
    VRT_CUDA(cudaMemcpy(ph.data(), quad->phi, sizeof(double) * quad->n_dirs, cudaMemcpyDefault));
    int64_t d_lo = 0, d_hi = quad->n_dirs;
    if (c.dir_end > c.dir_begin) {
        if (c.dir_begin < 0 || c.dir_end > quad->n_dirs) {
            set_error("solver: direction shard out of range");
            return VRT_E_INVALID;
        }
        d_lo = c.dir_begin;
        d_hi = c.dir_end;
        s->dir_sharded = true;
    }
    for (int64_t i = d_lo; i < d_hi; i++) {
        double t = th[i], p = ph[i];
        // k = [cosθ, cosϕ sinθ, sinϕ sinθ] (lambda_iteration.jl:87); θ == 90 is skipped by both branches (:98,:104)
        if (!(t > 90) && !(t < 90)) continue;
        std::array<double, 3> k = {cos(t * PI / 180), cos(p * PI / 180) * sin(t * PI / 180), sin(p * PI / 180) * sin(t * PI / 180)};
        int down = !(t > 90);
        int rc = VRT_OK;
        DirSchedule* sc = nullptr;
        if (!g->regular) {
            sc = schedule_get(g, k.data(), down, c.n_sweeps, c.p, c.prune, cv, &rc);
            if (!sc) return rc;
        }
        s->qk.push_back(k);
        s->qw.push_back(w[i]);
        s->qdown.push_back(down);
        s->sch.push_back(sc);
    }
    s->nd = (int)s->qk.size();
    // Directions in flight together share the S rows they read when they walk through the grid side by side: the batches
    // are therefore formed from directions of the same sense (up / down) and neighbouring inclination.  J then adds the
    // directions in this order instead of the order of the quadrature file (rounding-level difference; VRT_DIR_ORDER=0
    // keeps the file order).
    s->order.resize(s->nd);
    s->dir_lam.assign(s->nd, std::array<int64_t, 2>{0, 0});
    for (int d = 0; d < s->nd; d++) s->order[d] = d;
    {
        const char* e = getenv("VRT_DIR_ORDER");
        if (e && atoi(e) == 1 && !g->regular)
            std::stable_sort(s->order.begin(), s->order.end(), [&](int a, int b) {
                if (s->qdown[a] != s->qdown[b]) return s->qdown[a] < s->qdown[b];
                return fabs(s->qk[a][0]) > fabs(s->qk[b][0]);      // steep rays first
            });
    }
    s->n1_up = g->off_up[1] - 1;
    s->n1_dn = g->off_down[1] - 1;
    if (g->regular && s->cell_R > 1) {
        set_error("solver: cell shards are not implemented on the regular grid (direction and wavelength shards are)");
        return VRT_E_INVALID;
    }
    VRT_TRY(s->diff_bits.alloc(1));
    VRT_TRY(s->diff_nan.alloc(1));
    return VRT_OK;
}

static int upload_site_vec(const vrt_grid* g, const double* src, DevBuf<double>& dst, DevBuf<double>& stage, const char* name) {
    if (!src) {
        set_error("solver: site array '%s' is NULL", name);
        return VRT_E_INVALID;
    }
    VRT_TRY(dst.alloc(g->n));
    return upload_rows(g, src, dst.p, 1, stage);
}

// choose (λ-chunk, directions in flight) from free HBM and allocate the work buffers
static int plan_buffers(vrt_solver* s) {
    if (s->lc > 0) return VRT_OK;
    const int64_t n = s->n;
    if (s->g->regular) {
        // per wavelength: alpha_tot in the caller's layout (line only) next to regular.cu's five internal volumes
        int64_t lc = 0;
        VRT_TRY(regular_plan_chunk(s->g, s->nlam, s->is_line ? 1.0 : 0.0, &lc));
        if (s->cfg.lam_chunk > 0) lc = std::min<int64_t>(lc, s->cfg.lam_chunk);
        if (s->is_line) lc = std::min<int64_t>(lc, 384);   // two opacity tiles per CTA in shared memory
        const int64_t passes = (s->nlam + lc - 1) / lc;
        lc = (s->nlam + passes - 1) / passes;
        s->lc = lc;
        s->db = 1;
        s->alpha_p.assign(1, nullptr);
        s->I_p.assign(1, nullptr);
        s->scr_p.assign(1, std::array<double*, MAX_SWEEPS>{});
        auto* b0 = new DevBuf<double>();   // boundary plane [nx*ny][lc]
        s->bufs.push_back(b0);
        VRT_TRY(b0->alloc((size_t)(s->g->rnx * s->g->rny) * lc));
        s->I_p[0] = b0->p;
        if (s->is_line) {
            auto* ba = new DevBuf<double>();
            s->bufs.push_back(ba);
            VRT_TRY(ba->alloc((size_t)n * lc + 2));
            s->alpha_p[0] = ba->p;
        }
        return VRT_OK;
    }
    size_t free_b = 0, total_b = 0;
    VRT_CUDA(cudaMemGetInfo(&free_b, &total_b));
    double budget = 0.85 * (double)free_b;
    auto per_dir_rows = [&](int d) {
        double rows = (s->is_line ? 2.0 : 1.0) * (double)n;  // I_main (+ alpha for the line; the continuum shares α)
        for (int k = 0; k < s->cfg.n_sweeps - 1; k++) rows += (double)s->sch[d]->scr_rows[k];
        return rows;
    };
    double rows_all = 0, rows_max = 0;
    for (int d = 0; d < s->nd; d++) {
        rows_all += per_dir_rows(d);
        rows_max = std::max(rows_max, per_dir_rows(d));
    }
    int64_t lc = s->cfg.lam_chunk > 0 ? std::min<int64_t>(s->cfg.lam_chunk, s->nlam) : s->nlam;
    // the opacity kernel keeps two OP_TC x lc tiles in shared memory (227 KB per CTA at most): wider chunks are split
    if (s->is_line) lc = std::min<int64_t>(lc, 384);
    const char* envl = getenv("VRT_LAM_CHUNK");
    if (envl && atoi(envl) > 0) lc = std::min<int64_t>(atoi(envl), std::min<int64_t>(s->nlam, s->is_line ? 384 : s->nlam));
    int db = std::min(s->nd, MAX_DIRS);
    const char* env = getenv("VRT_MAX_DIRS");
    if (env && atoi(env) > 0) db = std::min(db, atoi(env));
    auto fits = [&](int64_t lcc, int dbb) { return rows_max * dbb * (double)lcc * 8.0 <= budget; };
    // The sweep costs per (cell, direction) visit and per pass, far more than per byte: wide wavelength rows come first
    // (fewer passes over the visit lists, wide TMA rows), then as many directions in flight as still fit.  At least
    // min(nd, 2) directions are kept in flight because the dataflow order hides dependency latency behind the other
    // directions' work.  Measured: on 1 M sites 2 directions in flight cost 33 % more than 6 or more, but a second pass
    // over half-width rows costs 71 % more; on 16 M sites (91 wavelengths x 2 directions) beats (46 x 4) by 24 %
    // (profiles/README.md), so the width wins down to two directions.
    const int db_min = std::min(db, 2);
    if (s->cfg.lam_chunk <= 0 && !(envl && atoi(envl) > 0)) {
        if (!fits(lc, db_min)) lc = std::max<int64_t>(1, (int64_t)(budget / (rows_max * db_min * 8.0)));
        // even out the chunks: ceil(nlam / passes)
        const int64_t passes = (s->nlam + lc - 1) / lc;
        lc = (s->nlam + passes - 1) / passes;
    }
    while (!fits(lc, db) && db > 1) db--;
    while (!fits(lc, db) && lc > 1) lc = (lc + 1) / 2;
    if (!fits(lc, db)) {
        set_error("not enough device memory for one direction x one wavelength (%.1f GB free)", free_b / 1e9);
        return VRT_E_NOMEM;
    }
    s->lc = lc;
    s->db = db;
    s->alpha_p.assign(db, nullptr);
    s->I_p.assign(db, nullptr);
    s->scr_p.assign(db, std::array<double*, MAX_SWEEPS>{});
    // any direction may end up in any buffer slot (whole and wavelength-split directions are batched separately): every slot is
    // sized for the largest scratch need
    int64_t scr[MAX_SWEEPS] = {0};
    for (int d = 0; d < s->nd; d++)
        for (int k = 0; k < MAX_SWEEPS; k++) scr[k] = std::max(scr[k], s->sch[d]->scr_rows[k]);
    for (int j = 0; j < db; j++) {
        auto* bi = new DevBuf<double>();
        s->bufs.push_back(bi);
        VRT_TRY(bi->alloc((size_t)n * lc + 2));
        s->I_p[j] = bi->p;
        if (s->is_line) {
            auto* ba = new DevBuf<double>();
            s->bufs.push_back(ba);
            VRT_TRY(ba->alloc((size_t)n * lc + 2));
            s->alpha_p[j] = ba->p;
        }
        for (int k = 0; k < s->cfg.n_sweeps - 1; k++) {
            auto* bs = new DevBuf<double>();
            s->bufs.push_back(bs);
            VRT_TRY(bs->alloc((size_t)std::max<int64_t>(scr[k], 1) * lc + 2));
            s->scr_p[j][k] = bs->p;
        }
    }
    return VRT_OK;
}

// One exchange step between the processes (ops of vrt_allreduce_fn).  In-library NCCL: enqueued on the collectives' stream
// between two events, so it is ordered after everything issued so far and before everything issued afterwards, without a
// host synchronisation.  Host hook: the device is synchronised and the callback does the rest.
static int exchange(vrt_solver* s, double* buf, int64_t count, int op) {
    if (s->comm) {
        cudaStream_t cs = comm_stream(s->comm);
        for (auto& e : s->comm_ev)
            if (!e) VRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        VRT_CUDA(cudaEventRecord(s->comm_ev[0], 0));
        VRT_CUDA(cudaStreamWaitEvent(cs, s->comm_ev[0], 0));
        VRT_TRY(comm_op(s->comm, buf, count, op, cs));
        VRT_CUDA(cudaEventRecord(s->comm_ev[1], cs));
        VRT_CUDA(cudaStreamWaitEvent(0, s->comm_ev[1], 0));
        return VRT_OK;
    }
    if (s->allreduce) {
        VRT_CUDA(cudaDeviceSynchronize());
        int rc = s->allreduce(buf, count, op, s->allreduce_user);
        if (rc != 0) {
            set_error("exchange hook failed (%d) in op %d", rc, op);
            return VRT_E_STATE;
        }
    }
    return VRT_OK;
}

// The all-gather of S at the end of an iteration is not needed until the next sweep: with in-library collectives it is only
// enqueued on the collectives' stream, and whoever reads S beyond the own cell slice calls wait_gather() first.  The next
// iteration's γ, boundary and opacity kernels (which do not read S) run under it.
static int exchange_S_deferred(vrt_solver* s) {
    if (!s->comm || getenv("VRT_NO_DEFERRED_GATHER")) return exchange(s, s->S.p, s->n_pad * s->nlam, 4);
    cudaStream_t cs = comm_stream(s->comm);
    for (auto& e : s->comm_ev)
        if (!e) VRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!s->gather_ev) VRT_CUDA(cudaEventCreateWithFlags(&s->gather_ev, cudaEventDisableTiming));
    VRT_CUDA(cudaEventRecord(s->comm_ev[0], 0));
    VRT_CUDA(cudaStreamWaitEvent(cs, s->comm_ev[0], 0));
    VRT_TRY(comm_op(s->comm, s->S.p, s->n_pad * s->nlam, 4, cs));
    VRT_CUDA(cudaEventRecord(s->gather_ev, cs));
    s->gather_pending = true;
    return VRT_OK;
}
static int wait_gather(vrt_solver* s) {
    if (s->gather_pending) {
        VRT_CUDA(cudaStreamWaitEvent(0, s->gather_ev, 0));
        s->gather_pending = false;
    }
    return VRT_OK;
}

// J_λ_voronoi on device state: s->S -> s->J (internal order)
// scatter: false = all-reduce J over the direction shards (every rank gets the full J); true = reduce-scatter over cells
// jmode: what happens to the partial J of a direction shard: 0 all-reduce (every rank gets the full J), 1 reduce-scatter over
// cells, 2 nothing (the owners of the cell slices read it through peer memory)
static int mean_intensity_internal(vrt_solver* s, SweepStats* stats, double* t_opacity_ms, double* t_sweep_ms, int jmode = 0) {
    const int64_t n = s->n;
    VRT_TRY(plan_buffers(s));
    s->ev_used = 0;
    cudaEvent_t e0, e1;
    VRT_TRY(s->event(&e0));
    VRT_TRY(s->event(&e1));
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> op_ev;   // opacity launches of the irregular path, resolved after the final sync
    float opacity_ms = 0;
    if (s->is_line) {
        k_gamma<<<nblocks(n, 256), 256>>>(n, s->ld, s->T.p, s->ne.p, s->pops.p, s->gamma.p);
        s->gamma_valid = true;
        stats->kernels += 1;
    }
    if (s->g->regular) {
        // J_λ_regular (lambda_iteration.jl:1-58 / lambda_continuum.jl:1-24): directions one after the other, per direction
        // alpha_tot (:32-35), the bottom boundary B_λ(T[1,:,:]) for θ > 90 (:38) or zero (:46), the plane walk, J += w I
        const int64_t plane = s->g->rnx * s->g->rny;
        for (int64_t l0 = 0; l0 < s->nlam; l0 += s->lc) {
            const int64_t lc = std::min(s->lc, s->nlam - l0);
            bool have_S[2] = {false, false};
            uint64_t ws_token = 0;
            // directions grouped by the internal layout they are solved in (S is then laid out twice per chunk, not per
            // direction); J therefore sums the directions in that order, not in the order of the quadrature file
            std::vector<int> order;
            for (int lay = 0; lay < 2; lay++)
                for (int d = 0; d < s->nd; d++) {
                    int dl = 0;
                    VRT_TRY(regular_dir_layout(s->g, s->qk[d].data(), &dl));
                    if (dl == lay) order.push_back(d);
                }
            for (size_t oi = 0; oi < order.size(); oi++) {
                const int d = order[oi];
                const double* alpha = s->alpha_cont.p;
                int64_t a_ld = 1;
                if (s->is_line) {
                    OpacityDirs od;
                    od.nd = 1;
                    for (int a = 0; a < 3; a++) od.k[0][a] = s->qk[d][a];
                    od.alpha[0] = s->alpha_p[0];
                    VRT_CUDA(cudaEventRecord(e0));
                    const size_t shm = 2 * sizeof(double) * OP_TC * (size_t)((int)lc | 1);
                    const int grid = (int)std::min<int64_t>((n + OP_TC - 1) / OP_TC, 148 * 16);
                    if (shm > 48 * 1024) VRT_CUDA(cudaFuncSetAttribute(k_opacity, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
                    k_opacity<<<grid, OP_THREADS, shm>>>(n, lc, s->lam_dev.p + s->l_begin + l0, s->ld, od, s->gamma.p, s->dD.p, s->vz.p,
                                                        s->vx.p, s->vy.p, s->pops.p, s->alpha_cont.p);
                    VRT_CUDA(cudaEventRecord(e1));
                    VRT_CUDA(cudaGetLastError());
                    stats->kernels += 1;
                    alpha = s->alpha_p[0];
                    a_ld = lc;
                }
                const double* I0 = nullptr;
                if (!s->qdown[d]) {
                    k_regular_boundary<<<nblocks(plane * lc, 256), 256>>>(s->I_p[0], lc, s->lam_dev.p + s->l_begin + l0, s->T.p,
                                                                         s->is_line ? nullptr : s->B0.p, s->g->rnz, plane);
                    VRT_CUDA(cudaGetLastError());
                    stats->kernels += 1;
                    I0 = s->I_p[0];
                }
                VRT_TRY(regular_dir_accumulate(s->g, s->qk[d].data(), s->qdown[d], s->cfg.n_sweeps, lc, s->S.p, s->nlam, l0, alpha, a_ld, 0,
                                               I0, s->J.p, s->nlam, l0, s->qw[d], oi > 0, have_S, &ws_token, stats));
                if (s->is_line) {
                    float ms = 0;
                    VRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
                    opacity_ms += ms;
                }
            }
        }
        if (s->nd == 0) VRT_CUDA(cudaMemset(s->J.p, 0, sizeof(double) * (size_t)n * s->nlam));
        VRT_CUDA(cudaDeviceSynchronize());
        // direction shards (J_λ_regular's loop over the quadrature, lambda_iteration.jl:23-55, split over processes): J = Σ shards
        if (s->dir_sharded && s->has_exchange()) VRT_TRY(exchange(s, s->J.p, n * s->nlam, 2));
        if (t_opacity_ms) *t_opacity_ms = opacity_ms;
        if (t_sweep_ms) *t_sweep_ms = stats->sweep_ms;
        return VRT_OK;
    }
    // directions solved on all local wavelengths, in sweep order, and those restricted to a wavelength sub-range
    std::vector<int> full, part;
    for (int q = 0; q < s->nd; q++) {
        const int d = s->order[q];
        (s->dir_lam[d][1] > s->dir_lam[d][0] ? part : full).push_back(d);
    }
    bool J_started = false;
    // one launch group: directions `ds` (at most db) on the local wavelengths [l0, l0 + lc), buffer slots 0 .. ds.size()-1
    auto run_batch = [&](const std::vector<int>& ds, int64_t l0, int64_t lc, bool first) -> int {
        const int nb = (int)ds.size();
        std::vector<SweepDir> dirs(nb);
        OpacityDirs od;
        JDirs jd;
        od.nd = jd.nd = nb;
        for (int j = 0; j < nb; j++) {
            const int d = ds[j];
            dirs[j].sch = s->sch[d];
            dirs[j].I_main = s->I_p[j];
            dirs[j].alpha = s->is_line ? s->alpha_p[j] : s->alpha_cont.p;
            for (int k = 0; k < MAX_SWEEPS; k++) dirs[j].scratch[k] = s->scr_p[j][k];
            for (int a = 0; a < 3; a++) od.k[j][a] = s->qk[d][a];
            od.alpha[j] = s->alpha_p[j];
            jd.I[j] = s->I_p[j];
            jd.w[j] = s->qw[d];
            // boundary values (lambda_iteration.jl:98-106 / lambda_continuum.jl:44-51) + the never-processed site (Q1)
            if (!s->qdown[d]) {
                k_boundary_planck<<<nblocks(s->n1_up * lc, 256), 256>>>(s->I_p[j], lc, s->lam_dev.p + s->l_begin + l0, s->T.p,
                                                                       s->is_line ? nullptr : s->B0.p, s->n1_up);
                VRT_CUDA(cudaMemsetAsync(s->I_p[j] + (size_t)(n - 1) * lc, 0, sizeof(double) * lc));
            } else {
                k_boundary_rows<<<nblocks(s->n1_dn * lc, 256), 256>>>(s->I_p[j], lc, nullptr, s->g->perm_dn_int.p, s->n1_dn);
                k_boundary_rows<<<nblocks(lc, 256), 256>>>(s->I_p[j], lc, nullptr, s->g->perm_dn_int.p + (n - 1), 1);
                stats->kernels += 1;
            }
            stats->kernels += 1;
        }
        if (s->is_line) {
            cudaEvent_t o0, o1;
            VRT_TRY(s->event(&o0));
            VRT_TRY(s->event(&o1));
            op_ev.emplace_back(o0, o1);
            VRT_CUDA(cudaEventRecord(o0));
            {
                const size_t shm = 2 * sizeof(double) * OP_TC * (size_t)((int)lc | 1);
                const int grid = (int)std::min<int64_t>((n + OP_TC - 1) / OP_TC, 148 * 16);
                if (shm > 48 * 1024) VRT_CUDA(cudaFuncSetAttribute(k_opacity, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
                k_opacity<<<grid, OP_THREADS, shm>>>(n, lc, s->lam_dev.p + s->l_begin + l0, s->ld, od, s->gamma.p, s->dD.p, s->vz.p,
                                                    s->vx.p, s->vy.p, s->pops.p, s->alpha_cont.p);
            }
            VRT_CUDA(cudaEventRecord(op_ev.back().second));
            stats->kernels += 1;
        }
        VRT_CUDA(cudaGetLastError());
        VRT_TRY(wait_gather(s));   // the sweep reads S of every cell
        VRT_TRY(sweep_run(s->g, nb, dirs.data(), s->S.p + l0, s->nlam, lc, 0, stats));
        k_J_reduce<<<nblocks(n * lc, 256), 256>>>(n, lc, jd, s->J.p + l0, s->nlam, first ? 1 : 0);
        stats->kernels += 1;
        VRT_CUDA(cudaGetLastError());
        return VRT_OK;
    };
    for (int64_t l0 = 0; l0 < s->nlam; l0 += s->lc) {
        const int64_t lc = std::min(s->lc, s->nlam - l0);
        for (size_t d0 = 0; d0 < full.size(); d0 += (size_t)s->db) {
            std::vector<int> ds(full.begin() + d0, full.begin() + std::min(full.size(), d0 + (size_t)s->db));
            VRT_TRY(run_batch(ds, l0, lc, d0 == 0));
        }
    }
    J_started = !full.empty();
    if (!part.empty()) {
        // J of the wavelengths no full direction has written yet starts from zero
        if (!J_started) VRT_CUDA(cudaMemsetAsync(s->J.p, 0, sizeof(double) * (size_t)n * s->nlam));
        for (int d : part) {
            const int64_t lo = s->dir_lam[d][0], hi = s->dir_lam[d][1];
            for (int64_t l0 = lo; l0 < hi; l0 += s->lc) VRT_TRY(run_batch(std::vector<int>{d}, l0, std::min(s->lc, hi - l0), false));
        }
    }
    VRT_CUDA(cudaDeviceSynchronize());
    VRT_TRY(sweep_collect(s->g, stats));
    for (auto& pr : op_ev) {
        float ms = 0;
        VRT_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        opacity_ms += ms;
    }
    if (s->nd == 0) VRT_CUDA(cudaMemset(s->J.p, 0, sizeof(double) * (size_t)n * s->nlam));   // no direction: J = 0, whatever J held
    if (s->dir_sharded && s->has_exchange()) {
        // J = sum over the direction shards (lambda_iteration.jl:102,107 add the directions one after the other)
        if (jmode == 1) VRT_TRY(exchange(s, s->J.p, s->n_pad * s->nlam, 3));
        else if (jmode == 0) VRT_TRY(exchange(s, s->J.p, n * s->nlam, 2));
    }
    if (t_opacity_ms) *t_opacity_ms = opacity_ms;
    if (t_sweep_ms) *t_sweep_ms = stats->sweep_ms;
    return VRT_OK;
}

static int rates_internal(vrt_solver* s, const double* damping_int, int64_t c0, int64_t c1) {
    const int64_t n = s->n;
    if (!s->lte.p) {
        set_error("radiative rates need the LTE populations (vrt_site_data.lte_pops or vrt_solver_set_field)");
        return VRT_E_STATE;
    }
    if (!damping_int && !s->gamma_valid) {
        set_error("vrt_calculate_R: no damping given and vrt_mean_intensity has not been called");
        return VRT_E_STATE;
    }
    int64_t warps_needed = n;
    int bs = 256;
    int64_t blocks = std::min<int64_t>((warps_needed * 32 + bs - 1) / bs, 148 * 64);
    {
        const int64_t cn = c1 - c0;   // own cell slice (all cells unless cell-sharded)
        k_rates<<<(int)blocks, bs>>>(cn, n, s->nlam, s->rl.p, s->ld, s->T.p + c0, s->dD.p + c0, s->gamma.p + c0,
                                     damping_int ? damping_int + c0 * s->nlam : nullptr, s->lte.p + c0, s->J.p + c0 * s->nlam, s->Rp.p + c0);
    }
    VRT_CUDA(cudaGetLastError());
    if (s->has_exchange()) VRT_TRY(exchange(s, s->Rp.p, 6 * n, 0));
    return VRT_OK;
}

static int read_diff(vrt_solver* s, double* diff) {
    unsigned long long bits = 0;
    int isn = 0;
    VRT_CUDA(cudaMemcpy(&bits, s->diff_bits.p, sizeof(bits), cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaMemcpy(&isn, s->diff_nan.p, sizeof(isn), cudaMemcpyDeviceToHost));
    double d;
    memcpy(&d, &bits, sizeof(d));
    if (s->has_exchange()) {
        // max over shards; NaN is carried as a flag through the same reduction
        VRT_TRY(s->diff_pair.ensure(2));
        double h[2] = {d, isn ? 1.0 : 0.0};
        VRT_CUDA(cudaMemcpy(s->diff_pair.p, h, sizeof(h), cudaMemcpyHostToDevice));
        VRT_TRY(exchange(s, s->diff_pair.p, 2, 1));
        VRT_CUDA(cudaMemcpy(h, s->diff_pair.p, sizeof(h), cudaMemcpyDeviceToHost));
        d = h[0];
        isn = h[1] != 0.0;
    }
    *diff = isn ? NAN : d;
    return VRT_OK;
}

static int set_field(vrt_solver* s, int field, const double* data) {
    const vrt_grid* g = s->g;
    const int64_t n = s->n;
    if (!data) {
        set_error("vrt_solver_set_field: NULL data");
        return VRT_E_INVALID;
    }
    DevBuf<double> tmp;
    switch (field) {
        case VRT_FIELD_ALPHA_CONT:
            return upload_site_vec(g, data, s->alpha_cont, s->stage, "alpha_cont");
        case VRT_FIELD_DESTRUCTION:
            return upload_site_vec(g, data, s->eps, s->stage, "destruction");
        case VRT_FIELD_C:
            VRT_TRY(tmp.alloc((size_t)9 * n));
            VRT_TRY(copy_in(tmp.p, data, sizeof(double) * 9 * n));
            VRT_TRY(s->Cp.ensure((size_t)6 * n));
            k_C_in<<<nblocks(n, 256), 256>>>(n, tmp.p, g->site_of.p, s->Cp.p);
            break;
        case VRT_FIELD_LTE_POPS:
            VRT_TRY(tmp.alloc((size_t)3 * n));
            VRT_TRY(copy_in(tmp.p, data, sizeof(double) * 3 * n));
            VRT_TRY(s->lte.ensure((size_t)3 * n));
            k_gather_cols<<<nblocks(n, 256), 256>>>(tmp.p, s->lte.p, g->site_of.p, n, 3);
            break;
        default:
            set_error("vrt_solver_set_field: unknown field %d", field);
            return VRT_E_INVALID;
    }
    VRT_CUDA(cudaGetLastError());
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

static int ensure_state(vrt_solver* s) {
    VRT_TRY(wait_gather(s));   // every entry point that reads or writes the state passes here
    if (s->have_state) return VRT_OK;
    const int64_t n = s->n;
    if (s->is_line && !s->lte.p) {
        set_error("solver state needs the LTE populations");
        return VRT_E_STATE;
    }
    if (s->is_line) {
        // S = B_0, populations = LTE (lambda_iteration.jl:216-241)
        k_planck_rows<<<nblocks(n * s->nlam, 256), 256>>>(n, s->nlam, s->lam_dev.p + s->l_begin, s->T.p, s->S.p);
        VRT_CUDA(cudaMemcpy(s->pops.p, s->lte.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice));
    } else {
        VRT_CUDA(cudaMemcpy(s->S.p, s->B0.p, sizeof(double) * n, cudaMemcpyDeviceToDevice));  // lambda_continuum.jl:136-137
    }
    VRT_CUDA(cudaGetLastError());
    s->have_state = true;
    return VRT_OK;
}

}  // namespace vrt

// ================================================================= ABI
extern "C" {

int vrt_formal_solve(vrt_grid* g, const double k[3], int32_t down, double p, int32_t n_sweeps, int64_t nlam,
                     const double* S, const double* alpha, const double* I0, double* I_out) {
    if (!g || !k || !S || !alpha || !I_out || nlam <= 0) {
        set_error("vrt_formal_solve: bad arguments");
        return VRT_E_INVALID;
    }
    if (g->regular) {
        set_error("vrt_formal_solve: regular grid handle — use vrt_regular_formal_solve");
        return VRT_E_STATE;
    }
    const int64_t n = g->n;
    int rc = VRT_OK;
    DirSchedule* sch = schedule_get(g, k, down ? 1 : 0, n_sweeps, p, 1, chunk_visits(nlam), &rc);
    if (!sch) return rc;
    SweepStats stats;
    DevBuf<double> S_int, a_int, I_main, stage, scr[MAX_SWEEPS];
    VRT_TRY(S_int.alloc((size_t)n * nlam + 2)); VRT_TRY(a_int.alloc((size_t)n * nlam + 2)); VRT_TRY(I_main.alloc((size_t)n * nlam + 2));
    VRT_TRY(upload_rows(g, S, S_int.p, nlam, stage));
    VRT_TRY(upload_rows(g, alpha, a_int.p, nlam, stage));
    stats.kernels += 2;
    SweepDir dir;
    dir.sch = sch;
    dir.alpha = a_int.p;
    dir.I_main = I_main.p;
    for (int s = 0; s < MAX_SWEEPS; s++) dir.scratch[s] = nullptr;
    for (int s = 0; s < n_sweeps - 1; s++) {
        VRT_TRY(scr[s].alloc((size_t)std::max<int64_t>(sch->scr_rows[s], 1) * nlam + 2));
        dir.scratch[s] = scr[s].p;
    }
    // I = zero(S); I[perm[1:n1]] = I_0 (irregular_ray_tracing.jl:23,33-35)
    VRT_CUDA(cudaMemsetAsync(I_main.p, 0, sizeof(double) * (size_t)n * nlam));
    const int64_t n1 = (down ? g->off_down[1] : g->off_up[1]) - 1;
    if (n1 > 0) {
        const double* d_I0 = I0;
        DevBuf<double> I0_dev;
        if (I0 && !is_device_ptr(I0)) {
            VRT_TRY(I0_dev.alloc((size_t)n1 * nlam));
            VRT_CUDA(cudaMemcpy(I0_dev.p, I0, sizeof(double) * (size_t)n1 * nlam, cudaMemcpyHostToDevice));
            d_I0 = I0_dev.p;
        }
        k_boundary_rows<<<nblocks(n1 * nlam, 256), 256>>>(I_main.p, nlam, d_I0, down ? g->perm_dn_int.p : nullptr, n1);
        VRT_CUDA(cudaGetLastError());
        VRT_CUDA(cudaDeviceSynchronize());
        stats.kernels += 1;
    }
    VRT_TRY(sweep_run(g, 1, &dir, S_int.p, nlam, nlam, 0, &stats));
    VRT_TRY(download_rows(g, I_main.p, I_out, nlam, stage));
    stats.kernels += 1;
    VRT_CUDA(cudaDeviceSynchronize());
    VRT_TRY(sweep_collect(g, &stats));
    g_last_stats = stats;
    return VRT_OK;
}

int vrt_solver_create_line(vrt_grid* g, const vrt_line* line, const double* lambda, const vrt_site_data* sd,
                           const vrt_quadrature* quad, const vrt_config* cfg, vrt_solver** out) {
    if (!out) return VRT_E_INVALID;
    *out = nullptr;
    if (!g || !line || !lambda || !sd) {
        set_error("vrt_solver_create_line: bad arguments");
        return VRT_E_INVALID;
    }
    vrt_solver* s = new vrt_solver();
    struct Guard { vrt_solver* s; ~Guard() { delete s; } } guard{s};
    s->is_line = 1;
    s->line = *line;
    {
        int64_t nl = line->nlam;
        if (cfg && cfg->lam_end > cfg->lam_begin) nl = cfg->lam_end - cfg->lam_begin;
        VRT_TRY(solver_common_init(s, g, quad, cfg, nl));
    }
    const int64_t n = g->n;
    s->nlam_total = line->nlam;
    if (line->nlam <= 0 || line->lidx[3] != line->nlam || line->lidx[0] != 0) {
        set_error("vrt_solver_create_line: inconsistent wavelength index table");
        return VRT_E_INVALID;
    }
    s->l_begin = 0;
    s->nlam = line->nlam;
    if (s->cfg.lam_end > s->cfg.lam_begin) {
        if (s->cfg.lam_begin < 0 || s->cfg.lam_end > line->nlam) {
            set_error("vrt_solver_create_line: wavelength shard out of range");
            return VRT_E_INVALID;
        }
        s->l_begin = s->cfg.lam_begin;
        s->nlam = s->cfg.lam_end - s->cfg.lam_begin;
    }
    s->lambda.resize(line->nlam);
    VRT_CUDA(cudaMemcpy(s->lambda.data(), lambda, sizeof(double) * line->nlam, cudaMemcpyDefault));
    VRT_TRY(s->lam_dev.alloc(line->nlam));
    VRT_CUDA(cudaMemcpy(s->lam_dev.p, s->lambda.data(), sizeof(double) * line->nlam, cudaMemcpyHostToDevice));
    s->ld.lambda0 = line->lambda0;
    s->ld.Bij = line->Bij;
    s->ld.Bji = line->Bji;
    s->ld.c_line = H_PLANCK * C_0 / (4 * PI * (line->lambda0 * 1e-9));
    s->ld.c_unsold = line->c_unsold;
    s->ld.gamma_nat = line->gamma_natural;
    s->ld.c_lin = line->c_linear_stark;
    s->ld.c_quad = line->c_quadratic_stark;

    VRT_TRY(upload_site_vec(g, sd->temperature, s->T, s->stage, "temperature"));
    VRT_TRY(upload_site_vec(g, sd->electron_density, s->ne, s->stage, "electron_density"));
    VRT_TRY(upload_site_vec(g, sd->hydrogen_density, s->NH, s->stage, "hydrogen_density"));
    VRT_TRY(upload_site_vec(g, sd->velocity_z, s->vz, s->stage, "velocity_z"));
    VRT_TRY(upload_site_vec(g, sd->velocity_x, s->vx, s->stage, "velocity_x"));
    VRT_TRY(upload_site_vec(g, sd->velocity_y, s->vy, s->stage, "velocity_y"));
    VRT_TRY(upload_site_vec(g, sd->doppler_width, s->dD, s->stage, "doppler_width"));
    VRT_TRY(upload_site_vec(g, sd->alpha_cont, s->alpha_cont, s->stage, "alpha_cont"));
    if (sd->destruction) VRT_TRY(set_field(s, VRT_FIELD_DESTRUCTION, sd->destruction));
    if (sd->C) VRT_TRY(set_field(s, VRT_FIELD_C, sd->C));
    if (sd->lte_pops) VRT_TRY(set_field(s, VRT_FIELD_LTE_POPS, sd->lte_pops));
    VRT_TRY(s->gamma.alloc(n));
    VRT_TRY(s->S.alloc((size_t)s->n_pad * s->nlam + 2)); VRT_TRY(s->J.alloc((size_t)s->n_pad * s->nlam));
    VRT_TRY(s->pops.alloc((size_t)3 * n)); VRT_TRY(s->Rp.alloc((size_t)6 * n));
    VRT_CUDA(cudaMemset(s->J.p, 0, sizeof(double) * (size_t)s->n_pad * s->nlam));
    VRT_CUDA(cudaMemset(s->S.p, 0, sizeof(double) * ((size_t)s->n_pad * s->nlam + 2)));

    // per-wavelength rate constants (σic rates.jl:422-438, gaunt_bf :562-572, trapezoid weights)
    {
        std::vector<RateLam> rl((size_t)s->nlam);
        const double hc = H_PLANCK * C_0;
        const double E_inf = R_INF * C_0 * H_PLANCK;
        const double n_eff = sqrt(E_inf / (line->chi_j - line->chi_i));
        const double charge = (double)line->Z;
        const double sc = 4 * E_CHARGE * E_CHARGE / (3 * PI * sqrt(3.0) * EPS_0 * M_ELECTRON * C_0 * C_0 * R_INF);
        for (int64_t ll = 0; ll < s->nlam; ll++) {
            int64_t l = s->l_begin + ll;
            int range = l < line->lidx[1] ? 0 : (l < line->lidx[2] ? 1 : 2);
            int64_t start = line->lidx[range], stop = line->lidx[range + 1];
            RateLam q;
            memset(&q, 0, sizeof(q));
            double lm = s->lambda[l] * 1e-9;
            q.lam_m = lm;
            q.range = range;
            double W = 0;
            if (l + 1 < stop) W += s->lambda[l + 1] * 1e-9 - lm;
            if (l > start) W += lm - s->lambda[l - 1] * 1e-9;
            q.W = W;
            double l5 = lm * lm * lm * lm * lm;
            if (range == 0) {
                q.planck = 2 * H_PLANCK * C_0 * C_0 / l5;
                q.sigma_bf = 0;
            } else {
                q.planck = 2 * hc * C_0 / l5;
                double lam_edge = s->lambda[stop - 1];
                double r = s->lambda[l] / lam_edge;
                double x = 1 / (lm * R_INF * charge * charge);
                double x3 = pow(x, 1.0 / 3);
                double nsqx = 1 / (n_eff * n_eff * x);
                double gbf = 1 + 0.1728 * x3 * (1 - 2 * nsqx) - 0.0496 * (x3 * x3) * (1 - (1 - nsqx) * 0.66666667 * nsqx);
                q.sigma_bf = sc * (charge * charge * charge * charge) * n_eff * (r * r * r) * gbf;
            }
            rl[ll] = q;
        }
        VRT_TRY(s->rl.alloc(s->nlam));
        VRT_CUDA(cudaMemcpy(s->rl.p, rl.data(), sizeof(RateLam) * s->nlam, cudaMemcpyHostToDevice));
    }
    guard.s = nullptr;
    *out = s;
    return VRT_OK;
}

int vrt_solver_create_continuum(vrt_grid* g, const double* alpha_cont, const double* eps, const double* B0,
                                const vrt_quadrature* quad, const vrt_config* cfg, vrt_solver** out) {
    if (!out) return VRT_E_INVALID;
    *out = nullptr;
    if (!g || !alpha_cont || !eps || !B0) {
        set_error("vrt_solver_create_continuum: bad arguments");
        return VRT_E_INVALID;
    }
    vrt_solver* s = new vrt_solver();
    struct Guard { vrt_solver* s; ~Guard() { delete s; } } guard{s};
    s->is_line = 0;
    VRT_TRY(solver_common_init(s, g, quad, cfg, 1));
    const int64_t n = g->n;
    s->nlam_total = s->nlam = 1;
    s->l_begin = 0;
    s->lambda.assign(1, 500.0);
    VRT_TRY(s->lam_dev.alloc(1));
    VRT_CUDA(cudaMemcpy(s->lam_dev.p, s->lambda.data(), sizeof(double), cudaMemcpyHostToDevice));
    VRT_TRY(upload_site_vec(g, alpha_cont, s->alpha_cont, s->stage, "alpha_cont"));
    VRT_TRY(upload_site_vec(g, eps, s->eps, s->stage, "eps"));
    VRT_TRY(upload_site_vec(g, B0, s->B0, s->stage, "B0"));
    VRT_TRY(s->T.alloc(1));
    VRT_TRY(s->S.alloc(n + 2)); VRT_TRY(s->J.alloc(n));
    VRT_CUDA(cudaMemset(s->J.p, 0, sizeof(double) * n));
    guard.s = nullptr;
    *out = s;
    return VRT_OK;
}

void vrt_solver_destroy(vrt_solver* s) { delete s; }

int vrt_solver_set_allreduce(vrt_solver* s, vrt_allreduce_fn fn, void* user) {
    if (!s) return VRT_E_INVALID;
    s->allreduce = fn;
    s->allreduce_user = user;
    return VRT_OK;
}

int vrt_solver_set_field(vrt_solver* s, int32_t field, const double* data) {
    if (!s || !s->is_line) {
        set_error("vrt_solver_set_field: needs a line solver");
        return VRT_E_INVALID;
    }
    return set_field(s, field, data);
}

int vrt_solver_nlam_local(const vrt_solver* s, int64_t* nlam_local) {
    if (!s || !nlam_local) return VRT_E_INVALID;
    *nlam_local = s->nlam;
    return VRT_OK;
}

int vrt_mean_intensity(vrt_solver* s, const double* S, const double* populations, double* J, double* damping) {
    if (!s || !S || !J) {
        set_error("vrt_mean_intensity: bad arguments");
        return VRT_E_INVALID;
    }
    const int64_t n = s->n;
    SweepStats stats;
    VRT_TRY(wait_gather(s));
    VRT_TRY(upload_rows(s->g, S, s->S.p, s->nlam, s->stage));
    stats.kernels += 1;
    if (s->is_line) {
        if (!populations) {
            set_error("vrt_mean_intensity: populations required for the line solver");
            return VRT_E_INVALID;
        }
        DevBuf<double> tmp;
        const double* src = populations;
        if (!is_device_ptr(populations)) {
            VRT_TRY(tmp.alloc((size_t)3 * n));
            VRT_CUDA(cudaMemcpy(tmp.p, populations, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
            src = tmp.p;
        }
        k_gather_cols<<<nblocks(n, 256), 256>>>(src, s->pops.p, s->g->site_of.p, n, 3);
        VRT_CUDA(cudaGetLastError());
        VRT_CUDA(cudaDeviceSynchronize());
        stats.kernels += 1;
    }
    s->have_state = true;
    VRT_TRY(mean_intensity_internal(s, &stats, nullptr, nullptr));
    VRT_TRY(download_rows(s->g, s->J.p, J, s->nlam, s->stage));
    stats.kernels += 1;
    if (damping && s->is_line) {
        DevBuf<double> dmp;
        VRT_TRY(dmp.alloc((size_t)n * s->nlam));
        k_damping<<<nblocks(n * s->nlam, 256), 256>>>(n, s->nlam, s->lam_dev.p + s->l_begin, s->gamma.p, s->dD.p, dmp.p);
        VRT_CUDA(cudaGetLastError());
        VRT_TRY(download_rows(s->g, dmp.p, damping, s->nlam, s->stage));
        VRT_CUDA(cudaDeviceSynchronize());
        stats.kernels += 2;
    }
    VRT_CUDA(cudaDeviceSynchronize());
    g_last_stats = stats;
    return VRT_OK;
}

int vrt_calculate_R(vrt_solver* s, const double* J, const double* damping, double* R) {
    if (!s || !s->is_line || !J || !R) {
        set_error("vrt_calculate_R: bad arguments");
        return VRT_E_INVALID;
    }
    const int64_t n = s->n;
    VRT_TRY(upload_rows(s->g, J, s->J.p, s->nlam, s->stage));
    DevBuf<double> dmp;
    const double* dmp_int = nullptr;
    if (damping) {
        VRT_TRY(dmp.alloc((size_t)n * s->nlam));
        VRT_TRY(upload_rows(s->g, damping, dmp.p, s->nlam, s->stage));
        dmp_int = dmp.p;
    }
    VRT_TRY(rates_internal(s, dmp_int, 0, n));
    DevBuf<double> tmp;
    double* out = R;
    if (!is_device_ptr(R)) {
        VRT_TRY(tmp.alloc((size_t)9 * n));
        out = tmp.p;
    }
    k_R_out<<<nblocks(n, 256), 256>>>(n, s->Rp.p, s->g->site_of.p, out);
    VRT_CUDA(cudaGetLastError());
    if (out != R) VRT_TRY(copy_out(R, out, sizeof(double) * 9 * n));
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

int vrt_get_revised_populations(int64_t n, const double* R, const double* Cm, const double* N_H, double* populations) {
    if (n <= 0 || !R || !Cm || !N_H || !populations) {
        set_error("vrt_get_revised_populations: bad arguments");
        return VRT_E_INVALID;
    }
    DevBuf<double> dR, dC, dN, dP;
    const double *pR = R, *pC = Cm, *pN = N_H;
    double* pP = populations;
    if (!is_device_ptr(R)) { VRT_TRY(dR.alloc((size_t)9 * n)); VRT_TRY(copy_in(dR.p, R, sizeof(double) * 9 * n)); pR = dR.p; }
    if (!is_device_ptr(Cm)) { VRT_TRY(dC.alloc((size_t)9 * n)); VRT_TRY(copy_in(dC.p, Cm, sizeof(double) * 9 * n)); pC = dC.p; }
    if (!is_device_ptr(N_H)) { VRT_TRY(dN.alloc(n)); VRT_TRY(copy_in(dN.p, N_H, sizeof(double) * n)); pN = dN.p; }
    if (!is_device_ptr(populations)) { VRT_TRY(dP.alloc((size_t)3 * n)); pP = dP.p; }
    k_stateq_abi<<<nblocks(n, 256), 256>>>(n, pR, pC, pN, pP);
    VRT_CUDA(cudaGetLastError());
    if (pP != populations) VRT_TRY(copy_out(populations, pP, sizeof(double) * 3 * n));
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

int vrt_lambda_iterate(vrt_solver* s, double eps, int32_t maxiter, vrt_iter_cb cb, void* user, vrt_result* out) {
    if (!s) return VRT_E_INVALID;
    const int64_t n = s->n;
    auto t_begin = std::chrono::steady_clock::now();
    if (s->is_line && (!s->eps.p || !s->Cp.p || !s->lte.p)) {
        set_error("vrt_lambda_iterate: destruction, C and lte_pops must be set");
        return VRT_E_STATE;
    }
    VRT_TRY(ensure_state(s));
    const int use_thick = s->is_line ? 0 : 1;
    // first criterion: S_old = zero(S_new) (lambda_iteration.jl:241, lambda_continuum.jl:139)
    VRT_CUDA(cudaMemset(s->diff_bits.p, 0, sizeof(unsigned long long)));
    VRT_CUDA(cudaMemset(s->diff_nan.p, 0, sizeof(int)));
    k_criterion<<<nblocks(n * s->nlam, 256), 256>>>(n, s->nlam, s->S.p, nullptr, s->eps.p, use_thick, s->diff_bits.p, s->diff_nan.p);
    VRT_CUDA(cudaGetLastError());
    double diff = 0;
    VRT_TRY(read_diff(s, &diff));
    int i = 0;
    SweepStats total;
    // post-J stages on this rank's cell slice only (reduce-scatter of J, all-gather of S and populations)
    const bool cshard = s->cell_R > 1 && s->dir_sharded && s->has_exchange() && s->is_line;
    const int64_t c0 = cshard ? s->c0 : 0, c1 = cshard ? s->c1 : n, cn = c1 - c0;
    while (diff > eps && i < maxiter) {
        auto t0 = std::chrono::steady_clock::now();
        SweepStats stats;
        vrt_iter_info info;
        memset(&info, 0, sizeof(info));
        info.diff = diff;
        const bool via_peers = cshard && s->peers && s->comm && !getenv("VRT_NO_PEER_REDUCE");
        VRT_TRY(mean_intensity_internal(s, &stats, &info.t_opacity_ms, &info.t_sweep_ms, cshard ? (via_peers ? 2 : 1) : 0));
        if (via_peers) {
            // every process's partial J must be complete before a peer reads it: a one-element all-reduce is the barrier
            VRT_TRY(s->diff_pair.ensure(2));
            VRT_TRY(exchange(s, s->diff_pair.p, 1, 1));
        }
        cudaEvent_t e[4];
        for (auto& ev : e) VRT_TRY(s->event(&ev));   // pooled: mean_intensity_internal reset the pool for this iteration
        VRT_CUDA(cudaMemset(s->diff_bits.p, 0, sizeof(unsigned long long)));
        VRT_CUDA(cudaMemset(s->diff_nan.p, 0, sizeof(int)));
        VRT_CUDA(cudaEventRecord(e[0]));
        if (cn > 0 && via_peers) {
            PeerJ pj;
            pj.R = s->cell_R;
            for (int r = 0; r < s->cell_R; r++) pj.J[r] = s->peerJ[r];
            k_source_update_peers<<<nblocks(cn * s->nlam, 256), 256>>>(cn, s->nlam, s->lam_dev.p + s->l_begin, s->T.p + c0, s->eps.p + c0, pj,
                                                                       c0 * s->nlam, s->J.p + c0 * s->nlam, s->S.p + c0 * s->nlam,
                                                                       s->diff_bits.p, s->diff_nan.p);
        } else if (cn > 0)
            k_source_update<<<std::min(nblocks(cn * s->nlam, 256), 148 * 64), 256>>>(cn, s->nlam, s->lam_dev.p + s->l_begin, s->T.p + c0,
                                                                 s->is_line ? nullptr : s->B0.p + c0, s->eps.p + c0, s->J.p + c0 * s->nlam,
                                                                 s->S.p + c0 * s->nlam, use_thick, s->diff_bits.p, s->diff_nan.p);
        VRT_CUDA(cudaEventRecord(e[1]));
        stats.kernels += 1;
        if (s->is_line) {
            VRT_TRY(rates_internal(s, nullptr, c0, c1));
            VRT_CUDA(cudaEventRecord(e[2]));
            if (cn > 0) k_stateq_soa<<<nblocks(cn, 256), 256>>>(cn, n, s->Rp.p + c0, s->Cp.p + c0, s->NH.p + c0, s->pops.p + c0);
            VRT_CUDA(cudaEventRecord(e[3]));
            stats.kernels += 2;
        }
        if (cshard) {
            // every rank now owns S and the populations of its cell slice: all-gather both over the direction group
            VRT_TRY(s->pops_blk.ensure((size_t)3 * s->n_pad));
            k_pack_pops<<<nblocks(s->cs, 256), 256>>>(n, s->cs * s->cell_r, s->cs, s->pops.p, s->pops_blk.p + (size_t)3 * s->cs * s->cell_r);
            VRT_CUDA(cudaGetLastError());
            VRT_TRY(exchange(s, s->pops_blk.p, 3 * s->n_pad, 4));
            k_unpack_pops<<<nblocks(n, 256), 256>>>(n, s->cs, s->cell_R, s->pops_blk.p, s->pops.p);
            stats.kernels += 2;
        }
        VRT_CUDA(cudaGetLastError());
        VRT_CUDA(cudaDeviceSynchronize());
        float ms = 0;
        VRT_CUDA(cudaEventElapsedTime(&ms, e[0], e[1]));
        info.t_source_ms = ms;
        if (s->is_line) {
            VRT_CUDA(cudaEventElapsedTime(&ms, e[1], e[2]));
            info.t_rates_ms = ms;
            VRT_CUDA(cudaEventElapsedTime(&ms, e[2], e[3]));
            info.t_stateq_ms = ms;
        }
        VRT_TRY(read_diff(s, &diff));
        // S of the other ranks' cells: after the criterion's reduction on the collectives' stream, so that the host does not
        // wait for it; the next sweep does (wait_gather)
        if (cshard) VRT_TRY(exchange_S_deferred(s));
        i++;
        info.iteration = i;
        info.updates = (double)n * s->nd * (double)s->nlam;
        info.t_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        total.kernels += stats.kernels; total.visits += stats.visits; total.steps += stats.steps; total.sweep_ms += stats.sweep_ms;
        if (cb && cb(&info, user) != 0) break;
    }
    VRT_TRY(wait_gather(s));
    g_last_stats = total;
    if (out) {
        out->iterations = i;
        out->converged = (diff == diff) && !(diff > eps);   // NaN stops the loop like the reference's `while diff > eps`, but is not convergence
        out->diff = diff;
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    }
    return VRT_OK;
}

int vrt_get_state(vrt_solver* s, double* S, double* J, double* populations) {
    if (!s) return VRT_E_INVALID;
    VRT_TRY(ensure_state(s));
    const int64_t n = s->n;
    if (S) VRT_TRY(download_rows(s->g, s->S.p, S, s->nlam, s->stage));
    if (J) {
        if (s->cell_R > 1 && s->dir_sharded && s->has_exchange() && s->is_line)   // J lives as cell slices: gather it (collective call)
            VRT_TRY(exchange(s, s->J.p, s->n_pad * s->nlam, 4));
        VRT_TRY(download_rows(s->g, s->J.p, J, s->nlam, s->stage));
    }
    if (populations && s->is_line) {
        DevBuf<double> tmp;
        double* o = populations;
        if (!is_device_ptr(populations)) { VRT_TRY(tmp.alloc((size_t)3 * n)); o = tmp.p; }
        k_scatter_cols<<<nblocks(n, 256), 256>>>(s->pops.p, o, s->g->site_of.p, n, 3);
        VRT_CUDA(cudaGetLastError());
        if (o != populations) VRT_TRY(copy_out(populations, o, sizeof(double) * 3 * n));
    }
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

// dst[i][l] = src[rank_of[i]][l]: rows of `cnt` consecutive sites in host site order out of the internal order
__global__ void k_gather_sites(const double* __restrict__ src, double* __restrict__ dst, const int32_t* __restrict__ rank_of,
                               int64_t cnt, int64_t nlam) {
    const int64_t total = cnt * nlam;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nlam, l = i - r * nlam;
        dst[i] = src[(int64_t)rank_of[r] * nlam + l];
    }
}

// write_to_file(S_λ, output_path) and write_to_file(populations, output_path) (io.jl:57-83): what Λ_voronoi does after every
// iteration (lambda_iteration.jl:280-281), straight from the device state: site-ordered row blocks are gathered on the
// device, copied into a pinned staging buffer and written at their place in the datasets.  Needs every wavelength on this
// process (no wavelength shard).
int vrt_output_write_state(vrt_outfile* f, vrt_solver* s) {
    if (!f || !s) return VRT_E_INVALID;
    VRT_TRY(ensure_state(s));
    int64_t fn = 0, fl = 0;
    VRT_TRY(outfile_shape(f, &fn, &fl));
    const int64_t n = s->n;
    if (fn != n || fl != s->nlam || s->nlam != s->nlam_total) {
        set_error("vrt_output_write_state: the file was created for %lld sites x %lld wavelengths, the solver holds %lld x %lld", (long long)fn,
                  (long long)fl, (long long)n, (long long)s->nlam);
        return VRT_E_INVALID;
    }
    const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(n, ((int64_t)64 << 20) / (8 * s->nlam)));
    DevBuf<double> dev;
    VRT_TRY(dev.alloc((size_t)rows * s->nlam));
    double* pin[2] = {nullptr, nullptr};
    struct Pins { double** p; ~Pins() { for (int i = 0; i < 2; i++) if (p[i]) cudaFreeHost(p[i]); } } pins{pin};
    for (int i = 0; i < 2; i++) VRT_CUDA(cudaMallocHost((void**)&pin[i], sizeof(double) * (size_t)rows * s->nlam));
    int b = 0;
    for (int64_t i0 = 0; i0 < n; i0 += rows, b ^= 1) {
        const int64_t cnt = std::min(rows, n - i0);
        k_gather_sites<<<nblocks(cnt * s->nlam, 256), 256>>>(s->S.p, dev.p, s->g->rank_of.p + i0, cnt, s->nlam);
        VRT_CUDA(cudaGetLastError());
        VRT_CUDA(cudaMemcpy(pin[b], dev.p, sizeof(double) * (size_t)cnt * s->nlam, cudaMemcpyDeviceToHost));
        VRT_TRY(outfile_write_at(f, "source_function", (uint64_t)i0 * s->nlam * 8, pin[b], sizeof(double) * (size_t)cnt * s->nlam));
    }
    if (s->is_line) {
        // populations (n_sites, 3) column-major: three site-ordered columns
        std::vector<double> host((size_t)3 * n);
        DevBuf<double> tmp;
        VRT_TRY(tmp.alloc((size_t)3 * n));
        k_scatter_cols<<<nblocks(n, 256), 256>>>(s->pops.p, tmp.p, s->g->site_of.p, n, 3);
        VRT_CUDA(cudaGetLastError());
        VRT_CUDA(cudaMemcpy(host.data(), tmp.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
        VRT_TRY(outfile_write_at(f, "populations", 0, host.data(), sizeof(double) * 3 * n));
    }
    return VRT_OK;
}

struct AbsOp {
    __host__ __device__ double operator()(const double& x) const { return fabs(x); }
};

int vrt_state_checksum(vrt_solver* s, double out[4]) {
    if (!s || !out) return VRT_E_INVALID;
    VRT_TRY(ensure_state(s));
    const int64_t n = s->n;
    const int64_t ns = n * s->nlam;
    if (ns >= (int64_t)INT32_MAX * 8) {
        set_error("vrt_state_checksum: state too large");
        return VRT_E_INVALID;
    }
    DevBuf<double> res;
    DevBuf<char> tmp;
    VRT_TRY(res.alloc(4));
    VRT_CUDA(cudaMemset(res.p, 0, 4 * sizeof(double)));
    cub::TransformInputIterator<double, AbsOp, const double*> absS(s->S.p, AbsOp());
    size_t b0 = 0, b1 = 0;
    VRT_CUDA(cub::DeviceReduce::Sum(nullptr, b0, s->S.p, res.p, ns));
    VRT_CUDA(cub::DeviceReduce::Max(nullptr, b1, absS, res.p + 1, ns));
    VRT_TRY(tmp.alloc(std::max(b0, b1)));
    size_t b = tmp.n;
    VRT_CUDA(cub::DeviceReduce::Sum(tmp.p, b, s->S.p, res.p, ns));
    b = tmp.n;
    VRT_CUDA(cub::DeviceReduce::Max(tmp.p, b, absS, res.p + 1, ns));
    if (s->is_line) {
        b = tmp.n;
        VRT_CUDA(cub::DeviceReduce::Sum(tmp.p, b, s->pops.p, res.p + 2, 3 * n));
    }
    // J lives as cell slices when the post-J stages are cell-sharded: sum of this rank's slice only
    const bool cshard = s->cell_R > 1 && s->dir_sharded && s->has_exchange() && s->is_line;
    const int64_t c0 = cshard ? s->c0 : 0, c1 = cshard ? s->c1 : n;
    b = tmp.n;
    if (c1 > c0) VRT_CUDA(cub::DeviceReduce::Sum(tmp.p, b, s->J.p + c0 * s->nlam, res.p + 3, (c1 - c0) * s->nlam));
    VRT_CUDA(cudaMemcpy(out, res.p, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return VRT_OK;
}

int vrt_solver_comm_init(vrt_solver* s, const char* dir_id, int32_t dir_rank, int32_t dir_size, const char* lam_id, int32_t lam_rank,
                         int32_t lam_size) {
    if (!s) return VRT_E_INVALID;
    if (s->cell_R > 1 && dir_size != s->cell_R) {
        set_error("vrt_solver_comm_init: the direction group has %d processes but the solver was created with %d cell shards", dir_size, s->cell_R);
        return VRT_E_INVALID;
    }
    if (s->cell_R > 1 && dir_rank != s->cell_r) {
        set_error("vrt_solver_comm_init: rank %d in the direction group but cell shard %d", dir_rank, s->cell_r);
        return VRT_E_INVALID;
    }
    if (s->comm) {
        comm_free(s->comm);
        s->comm = nullptr;
    }
    return comm_create(dir_id, dir_rank, dir_size, lam_id, lam_rank, lam_size, &s->comm);
}

// cost of each direction this solver holds (in the order of its quadrature table, θ = 90 rows left out): the number of
// (cell, sweep) visits of its sweep program.  Hosts use it to balance direction shards (longest processing time first).
int vrt_solver_direction_visits(const vrt_solver* s, int64_t* n_dirs, double* visits, int64_t capacity) {
    if (!s || !n_dirs) return VRT_E_INVALID;
    *n_dirs = s->nd;
    if (visits) {
        if (capacity < s->nd) {
            set_error("vrt_solver_direction_visits: capacity %lld < %d directions", (long long)capacity, s->nd);
            return VRT_E_INVALID;
        }
        for (int d = 0; d < s->nd; d++) visits[d] = s->sch[d] ? (double)s->sch[d]->n_visits : (double)s->n;
    }
    return VRT_OK;
}

int vrt_solver_peer_handle(vrt_solver* s, char handle[64]) {
    if (!s || !handle) return VRT_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    VRT_CUDA(cudaIpcGetMemHandle(&h, s->J.p));
    memcpy(handle, &h, 64);
    return VRT_OK;
}

int vrt_solver_peer_attach(vrt_solver* s, const char* handles, int32_t count) {
    if (!s || !handles) return VRT_E_INVALID;
    if (count != s->cell_R || count > 16 || s->cell_R < 2) {
        set_error("vrt_solver_peer_attach: %d handles for %d cell shards (at most 16)", (int)count, s->cell_R);
        return VRT_E_INVALID;
    }
    if (s->peers) return VRT_OK;
    for (int r = 0; r < count; r++) {
        if (r == s->cell_r) {
            s->peerJ[r] = s->J.p;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)64 * r, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; q++)
                if (q != s->cell_r && s->peerJ[q]) cudaIpcCloseMemHandle(const_cast<double*>(s->peerJ[q]));
            for (auto& pp : s->peerJ) pp = nullptr;
            return cuda_fail(e, "cudaIpcOpenMemHandle (peer J buffer)", __FILE__, __LINE__);
        }
        s->peerJ[r] = static_cast<const double*>(p);
    }
    s->peers = true;
    return VRT_OK;
}

int vrt_solver_peer_detach(vrt_solver* s) {
    if (!s) return VRT_E_INVALID;
    if (!s->peers) return VRT_OK;
    VRT_CUDA(cudaDeviceSynchronize());
    for (int r = 0; r < 16; r++) {
        if (s->peerJ[r] && r != s->cell_r) cudaIpcCloseMemHandle(const_cast<double*>(s->peerJ[r]));
        s->peerJ[r] = nullptr;
    }
    s->peers = false;
    return VRT_OK;
}

int vrt_solver_set_direction_lambda(vrt_solver* s, int32_t direction, int64_t lam_begin, int64_t lam_end) {
    if (!s) return VRT_E_INVALID;
    if (s->g->regular || !s->is_line) {
        set_error("vrt_solver_set_direction_lambda: only for the line solver on a Voronoi grid");
        return VRT_E_STATE;
    }
    if (direction < 0 || direction >= s->nd || lam_begin < 0 || lam_end > s->nlam || lam_end < lam_begin) {
        set_error("vrt_solver_set_direction_lambda: direction %d or range [%lld, %lld) out of bounds", (int)direction, (long long)lam_begin,
                  (long long)lam_end);
        return VRT_E_INVALID;
    }
    if (lam_end - lam_begin == s->nlam) lam_begin = lam_end = 0;   // the whole range: nothing special
    else if (lam_end > lam_begin && chunk_visits(lam_end - lam_begin) != s->sch[direction]->cv) {
        set_error("vrt_solver_set_direction_lambda: %lld wavelengths need another sweep program than the solver's (use >= 16)",
                  (long long)(lam_end - lam_begin));
        return VRT_E_INVALID;
    }
    s->dir_lam[direction] = {lam_begin, lam_end};
    return VRT_OK;
}

int vrt_solver_cell_slice(const vrt_solver* s, int64_t* first, int64_t* last) {
    if (!s || !first || !last) return VRT_E_INVALID;
    const bool cshard = s->cell_R > 1 && s->dir_sharded && s->is_line;
    *first = cshard ? s->c0 : 0;
    *last = cshard ? s->c1 : s->n;
    return VRT_OK;
}

// own cell slice, internal order: S, J rows [c0, c1) are contiguous, the populations are three runs of the SoA arrays
int vrt_get_state_slice(vrt_solver* s, double* S, double* J, double* populations) {
    if (!s) return VRT_E_INVALID;
    VRT_TRY(ensure_state(s));
    int64_t c0 = 0, c1 = 0;
    VRT_TRY(vrt_solver_cell_slice(s, &c0, &c1));
    const int64_t cn = c1 - c0, n = s->n;
    if (S) VRT_TRY(copy_out(S, s->S.p + c0 * s->nlam, sizeof(double) * (size_t)cn * s->nlam));
    if (J) VRT_TRY(copy_out(J, s->J.p + c0 * s->nlam, sizeof(double) * (size_t)cn * s->nlam));
    if (populations && s->is_line)
        for (int k = 0; k < 3; k++) VRT_TRY(copy_out(populations + (size_t)k * cn, s->pops.p + (size_t)k * n + c0, sizeof(double) * (size_t)cn));
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

int vrt_set_state_slice(vrt_solver* s, const double* S, const double* populations) {
    if (!s) return VRT_E_INVALID;
    VRT_TRY(ensure_state(s));
    int64_t c0 = 0, c1 = 0;
    VRT_TRY(vrt_solver_cell_slice(s, &c0, &c1));
    const int64_t cn = c1 - c0, n = s->n;
    const bool cshard = cn < n;
    if (cshard && !s->has_exchange()) {
        set_error("vrt_set_state_slice: cell-sharded solver without collectives (vrt_solver_comm_init / vrt_solver_set_allreduce)");
        return VRT_E_STATE;
    }
    if (S) {
        VRT_TRY(copy_in(s->S.p + c0 * s->nlam, S, sizeof(double) * (size_t)cn * s->nlam));
        // the other processes' slices come over NVLink, not over every process's PCIe link
        if (cshard) VRT_TRY(exchange(s, s->S.p, s->n_pad * s->nlam, 4));
    }
    if (populations && s->is_line) {
        for (int k = 0; k < 3; k++) VRT_TRY(copy_in(s->pops.p + (size_t)k * n + c0, populations + (size_t)k * cn, sizeof(double) * (size_t)cn));
        if (cshard) {
            VRT_TRY(s->pops_blk.ensure((size_t)3 * s->n_pad));
            k_pack_pops<<<nblocks(s->cs, 256), 256>>>(n, s->cs * s->cell_r, s->cs, s->pops.p, s->pops_blk.p + (size_t)3 * s->cs * s->cell_r);
            VRT_CUDA(cudaGetLastError());
            VRT_TRY(exchange(s, s->pops_blk.p, 3 * s->n_pad, 4));
            k_unpack_pops<<<nblocks(n, 256), 256>>>(n, s->cs, s->cell_R, s->pops_blk.p, s->pops.p);
            VRT_CUDA(cudaGetLastError());
        }
    }
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

int vrt_set_state(vrt_solver* s, const double* S, const double* populations) {
    if (!s) return VRT_E_INVALID;
    VRT_TRY(ensure_state(s));
    const int64_t n = s->n;
    if (S) VRT_TRY(upload_rows(s->g, S, s->S.p, s->nlam, s->stage));
    if (populations && s->is_line) {
        DevBuf<double> tmp;
        const double* src = populations;
        if (!is_device_ptr(populations)) {
            VRT_TRY(tmp.alloc((size_t)3 * n));
            VRT_CUDA(cudaMemcpy(tmp.p, populations, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
            src = tmp.p;
        }
        k_gather_cols<<<nblocks(n, 256), 256>>>(src, s->pops.p, s->g->site_of.p, n, 3);
        VRT_CUDA(cudaGetLastError());
    }
    VRT_CUDA(cudaDeviceSynchronize());
    return VRT_OK;
}

}  // extern "C"
