// regular.cu — short characteristics on the regular Cartesian grid (characteristics.jl:19-835).
//
// The reference walks the z planes in order; in every plane it picks the face the characteristic leaves through
// (argmin of the three path lengths, characteristics.jl:52-56) and calls one of three ray routines:
//   xy  (:191-373)  every point of the plane depends only on the previous plane            -> one parallel kernel
//   yz  (:383-604)  rows of constant x are visited in upwind order, n_sweeps times, each row reading the row before it
//   xz  (:614-835)  the same with x and y exchanged
// For yz/xz the update of a point is LINEAR in the two carried values it reads:
//       I[s][j] = cA*car[jl] + cB*car[ju] + cC,     jl = j-(sgn_j+1)/2, ju = jl+1
// with cA, cB, cC depending only on already-final data (S, alpha, previous plane).  The coefficients are computed once
// per plane by a fully parallel kernel (the reference recomputes them every sweep: 3x the exp() and the bilinear
// interpolations), and the dependent part is a thin recurrence of n_sweeps*(ns-2) row steps: one CTA per wavelength,
// the carried row in shared memory, the coefficient rows prefetched REG_PF steps ahead.
//
// Layout in HBM: the caller's arrays are the Julia arrays (nlam, nz, nx, ny) column-major (wavelength fastest).
// Internally a plane is [l][s][j] with j (the "parallel" axis of the recurrence: y for yz, x for xz) fastest, so that
// both the plane kernels and the recurrence read and write coalesced rows; which of x, y is j is fixed per direction
// because r_x and r_y do not depend on the plane.  Two tiled transposes convert in and out.
#include <algorithm>
#include <math.h>
#include <string.h>
#include <chrono>
#include <mutex>
#include "vrt_internal.h"

namespace vrt {
namespace {

constexpr int REG_PF = 12;   // coefficient rows in flight per thread in the recurrence (a third of it for rows wider than 514)

struct RegPlane {
    int np, ns;          // extents of the j and s axes (ghost columns included)
    int sgn_j, sgn_s;    // xy_intersect signs (functions.jl:430-457) mapped on the axes
    int par_is_x;        // j is x (xz branch / default) or y (yz branch)
    int lc;              // wavelengths in this chunk
    double kz, kj, ks;   // direction components along z, j, s
};

__host__ __device__ inline int reg_wrap(int i, int n) { return i == 0 ? n - 2 : (i == n - 1 ? 1 : i); }

// linear_weights (functions.jl:484-500)
__device__ __forceinline__ void reg_weights(double dtau, double& a, double& b, double& e) {
    if (dtau < 5e-4) {
        e = 1 - dtau + 0.5 * (dtau * dtau);
        a = dtau * (1.0 / 2 - dtau / 3);
        b = dtau * (1.0 / 2 - dtau / 6);
    } else if (dtau > 50) {
        e = 0.0;
        a = 1 / dtau;
        b = 1.0 - a;
    } else {
        e = exp(-dtau);
        a = (1 - e) / dtau - e;
        b = 1 - a - e;
    }
}

// ------------------------------------------------------------------ layout conversion
// src: caller's (nlam_src, nz, nx, ny) column-major, wavelengths [l0, l0+lc) -> dst [iz][l][s][j].
// Tile of 32 (q = l + lc*iz) x 32 (j) per block, one s per blockIdx.z; TO_INTERNAL = 0 is the inverse copy.
// For the inverse copy the result is w * value, added to what dst holds when `accumulate` (J += weights[i] * I).
template <int TO_INTERNAL>
__global__ void k_reg_transpose(const double* __restrict__ src, double* __restrict__ dst, int64_t nlam_src, int64_t l0,
                                int lc, int64_t nz, int64_t nx, int64_t ny, int par_is_x, double w, int accumulate) {
    __shared__ double tile[32][33];
    const int np = par_is_x ? (int)nx : (int)ny, ns = par_is_x ? (int)ny : (int)nx;
    const int64_t Q = (int64_t)lc * nz;
    const int64_t q0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32, s = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
    // user-side element: q along tx
    auto user_off = [&](int64_t q, int j) -> int64_t {
        const int64_t iz = q / lc, l = q - iz * lc;
        const int64_t ix = par_is_x ? j : s, iy = par_is_x ? s : j;
        return (l0 + l) + nlam_src * (iz + nz * (ix + nx * iy));
    };
    auto int_off = [&](int64_t q, int j) -> int64_t { return j + (int64_t)np * (s + (int64_t)ns * q); };
    if (TO_INTERNAL) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int j = j0 + ty + 8 * r;
            const int64_t q = q0 + tx;
            if (j < np && q < Q) tile[ty + 8 * r][tx] = src[user_off(q, j)];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int64_t q = q0 + ty + 8 * r;
            const int j = j0 + tx;
            if (j < np && q < Q) dst[int_off(q, j)] = tile[tx][ty + 8 * r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int64_t q = q0 + ty + 8 * r;
            const int j = j0 + tx;
            if (j < np && q < Q) tile[ty + 8 * r][tx] = src[int_off(q, j)];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int j = j0 + ty + 8 * r;
            const int64_t q = q0 + tx;
            if (j < np && q < Q) {
                const int64_t o = user_off(q, j);
                const double v = w * tile[tx][ty + 8 * r];
                dst[o] = accumulate ? dst[o] + v : v;
            }
        }
    }
}

// bilinear (functions.jl:328-355): the first coordinate selects the row of [Q11 Q12; Q21 Q22].  The three interpolations
// of one point (alpha, S, I) and all its wavelengths share the geometry: the corner weights are formed once per thread,
// with the two divisions of `bilinear` as reciprocals (same operation order otherwise).
struct RegBil {
    double ax, bx, idx, ay, by, idy;
    __device__ __forceinline__ RegBil(double xm, double ym, double x1, double x2, double y1, double y2)
        : ax(x2 - xm), bx(xm - x1), idx(1.0 / (x2 - x1)), ay(y2 - ym), by(ym - y1), idy(1.0 / (y2 - y1)) {}
    __device__ __forceinline__ double operator()(double Q11, double Q12, double Q21, double Q22) const {
        const double f1 = (ax * Q11 + bx * Q21) * idx;
        const double f2 = (ax * Q12 + bx * Q22) * idx;
        return (ay * f1 + by * f2) * idy;
    }
};

constexpr int REG_LPT = 4;   // wavelengths per thread in the plane kernels

// ------------------------------------------------------------------ xy branch
// xy_up_ray / xy_down_ray (characteristics.jl:191-280, :290-373).  One thread per (s, j) INCLUDING the ghost columns (a
// ghost evaluates the interior point it mirrors (:270-279): same arithmetic, no second pass) and REG_LPT wavelengths.
// Sc/ac: plane idz; Su/au/Iu: upwind plane; r = |dz / k_z| from the host.
__global__ void k_reg_xy(RegPlane P, double r, const double* __restrict__ cj, const double* __restrict__ cs,
                         const double* __restrict__ Sc, const double* __restrict__ ac, const double* __restrict__ Su,
                         const double* __restrict__ au, const double* __restrict__ Iu, double* __restrict__ Iout) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // flattened (s, j): np need not be a multiple of the block
    if (idx >= P.np * P.ns) return;
    const int s = idx / P.np, j = idx - s * P.np;
    const int gj = reg_wrap(j, P.np), gs = reg_wrap(s, P.ns);
    const double jinc = r * P.kj, sinc = r * P.ks;
    const int jl = gj - (P.sgn_j + 1) / 2, ju = jl + 1;
    const int sl = gs - (P.sgn_s + 1) / 2, su = sl + 1;
    const double jup = cj[gj] + jinc, sup = cs[gs] + sinc;
    const double jb1 = cj[jl], jb2 = cj[ju], sb1 = cs[sl], sb2 = cs[su];
    // bilinear's first coordinate is x: j when par_is_x, else s
    const RegBil W = P.par_is_x ? RegBil(jup, sup, jb1, jb2, sb1, sb2) : RegBil(sup, jup, sb1, sb2, jb1, jb2);
    const size_t pl = (size_t)P.np * P.ns;
    const size_t oll = (size_t)sl * P.np + jl, olu = (size_t)su * P.np + jl;   // (j low, s low), (j low, s up)
    // [Q11 Q12; Q21 Q22] = [x low,y low  x low,y up; x up,y low  x up,y up]
    const size_t o11 = oll, o22 = olu + 1;
    const size_t o12 = P.par_is_x ? olu : oll + 1, o21 = P.par_is_x ? oll + 1 : olu;
    const size_t oc = (size_t)gs * P.np + gj, oo = (size_t)s * P.np + j;
    const int l0 = blockIdx.y * REG_LPT, l1 = min(l0 + REG_LPT, P.lc);
#pragma unroll
    for (int q = 0; q < REG_LPT; q++) {
        const int l = l0 + q;
        if (l < l1) {
            const size_t b = (size_t)l * pl;
            const double a_u = W(au[b + o11], au[b + o12], au[b + o21], au[b + o22]);
            const double S_u = W(Su[b + o11], Su[b + o12], Su[b + o21], Su[b + o22]);
            const double I_u = W(Iu[b + o11], Iu[b + o12], Iu[b + o21], Iu[b + o22]);
            const double dtau = r * (ac[b + oc] + a_u) / 2;
            double wa, wb, we;
            reg_weights(dtau, wa, wb, we);
            Iout[b + oo] = we * I_u + wa * S_u + wb * Sc[b + oc];
        }
    }
}

// ------------------------------------------------------------------ yz / xz branch, parallel part
// Coefficients of the row recurrence for every interior point of plane idz (yz_*_ray :383-604, xz_*_ray :614-835).
// lo/hi: planes izl/izu of S and alpha; prev: the previous (upwind) plane of I; zc = z[idz], zb1 = z[izl], zb2 = z[izu];
// r = |Δs / k_s| from the host.  Centre values come from plane idz in the yz branch and from the UPPER plane izu in the
// xz branch: the host passes the right plane as S_c / a_c.
__global__ void k_reg_coef(RegPlane P, int up, double r, double zc, double zb1, double zb2, const double* __restrict__ cj,
                           const double* __restrict__ S_lo, const double* __restrict__ S_hi,
                           const double* __restrict__ a_lo, const double* __restrict__ a_hi, const double* __restrict__ S_c,
                           const double* __restrict__ a_c, const double* __restrict__ prev, double* __restrict__ cA,
                           double* __restrict__ cB, double* __restrict__ cC) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.np * P.ns) return;
    const int s = idx / P.np, j = idx - s * P.np;
    if (j < 1 || j > P.np - 2 || s < 1 || s > P.ns - 2) return;
    const double zup = zc + r * P.kz;
    const double jup = cj[j] + r * P.kj;
    const int jl = j - (P.sgn_j + 1) / 2, ju = jl + 1;
    const int sn = s + P.sgn_s;
    const RegBil W(zup, jup, zb1, zb2, cj[jl], cj[ju]);
    // I_u = I_fix + w_l*car[jl] + w_u*car[ju]: I_vals = [I_0 row; carried row] for up, [carried row; I_0 row] for down
    const double w_l = up ? W(0.0, 0.0, 1.0, 0.0) : W(1.0, 0.0, 0.0, 0.0);
    const double w_u = up ? W(0.0, 0.0, 0.0, 1.0) : W(0.0, 1.0, 0.0, 0.0);
    const size_t pl = (size_t)P.np * P.ns;
    const size_t ol = (size_t)sn * P.np + jl, ou = ol + 1, oc = (size_t)s * P.np + j;
    const int l0 = blockIdx.y * REG_LPT, l1 = min(l0 + REG_LPT, P.lc);
#pragma unroll
    for (int q = 0; q < REG_LPT; q++) {
        const int l = l0 + q;
        if (l < l1) {
            const size_t b = (size_t)l * pl;
            const double a_u = W(a_lo[b + ol], a_lo[b + ou], a_hi[b + ol], a_hi[b + ou]);
            const double S_u = W(S_lo[b + ol], S_lo[b + ou], S_hi[b + ol], S_hi[b + ou]);
            const double dtau = r * (a_c[b + oc] + a_u) / 2;
            double wa, wb, we;
            reg_weights(dtau, wa, wb, we);
            const double p_l = prev[b + ol], p_u = prev[b + ou];
            const double I_fix = up ? W(p_l, p_u, 0.0, 0.0) : W(0.0, 0.0, p_l, p_u);
            cA[b + oc] = we * w_l;
            cB[b + oc] = we * w_u;
            cC[b + oc] = we * I_fix + wa * S_u + wb * S_c[b + oc];
        }
    }
}

__device__ __forceinline__ double reg_ldg(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));   // ld.volatile: ptxas keeps it where it is written
                                                                            // (plain loads sink to the end of the unrolled group)
    return v;
}
__device__ __forceinline__ void reg_stg(double* p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// ------------------------------------------------------------------ yz / xz branch, dependent part
// One CTA per wavelength, one thread per INTERIOR j (thread t owns j = t+1).  Rows s are visited from the upwind side
// (range_bounds, functions.jl:466-475) n_sweeps times; the carried row survives from one sweep to the next and starts
// at zero (:403,:420 `I_carried = zero(...)` outside the sweep loop).  The periodic ghost entries of the carried row
// (:470-471) are never stored: a reader of column 0 / np-1 reads the interior column it mirrors.  Only the last sweep's
// values reach HBM; its rows 1 and ns-2 are also written to the ghost rows ns-1 and 0 (:476-480).
// The dependent chain of one row step is LDS, 2 FMA, STS, barrier; everything else (coefficient prefetch REG_PF rows
// ahead, row counters) is off the chain.
template <int MAXT, int PF>
__global__ void __launch_bounds__(MAXT) k_reg_rec(RegPlane P, int n_sweeps, const double* __restrict__ cA,
                                                   const double* __restrict__ cB, const double* __restrict__ cC,
                                                   double* __restrict__ Iout) {
    static_assert(PF % 2 == 0, "the carried-row double buffer is indexed by the parity of the unrolled step");
    extern __shared__ double car[];   // 2 x np carried rows
    const int np = P.np, ns = P.ns, nsi = P.ns - 2;
    // threads past the last interior column repeat it (same values to the same addresses) instead of idling behind a
    // predicate that would be re-evaluated in every step
    const int j = min((int)threadIdx.x + 1, np - 2);
    const size_t base = (size_t)blockIdx.x * np * ns;
    const int off = (P.sgn_j + 1) / 2;
    const int il = reg_wrap(j - off, np), iu = reg_wrap(j - off + 1, np);
    const int o0 = (P.sgn_s > 0 ? 1 : ns - 2) * np, stride = P.sgn_s * np;   // element offsets inside the plane (< 2^21)
    const int total = n_sweeps * nsi, last0 = total - nsi;
    const int o_end = o0 + stride * nsi;
    for (int q = threadIdx.x; q < 2 * np; q += blockDim.x) car[q] = 0.0;
    const double* pA = cA + base + j;
    const double* pB = cB + base + j;
    const double* pC = cC + base + j;
    double* pI = Iout + base + j;
    // keep the per-thread bases in registers: ptxas otherwise rebuilds them from blockIdx/params in every unrolled step
    asm volatile("" : "+l"(pA), "+l"(pB), "+l"(pC), "+l"(pI));
    int il1 = il + np, iu1 = iu + np, jw0 = j, jw1 = j + np, il0 = il, iu0 = iu;
    asm volatile("" : "+r"(il0), "+r"(iu0), "+r"(il1), "+r"(iu1), "+r"(jw0), "+r"(jw1));
    __syncthreads();
    double a[PF], b[PF], c[PF];
    int po = o0, so = o0;   // rows being prefetched / solved; both wrap into the next sweep
#pragma unroll
    for (int d = 0; d < PF; d++) {
        a[d] = reg_ldg(pA + po); b[d] = reg_ldg(pB + po); c[d] = reg_ldg(pC + po);
        po += stride;
        if (po == o_end) po = o0;
    }
    // one row step; PF is even, so the parity of tt is the parity of d and the double buffer is indexed statically
#define REG_STEP(d, tt)                                                                                     \
    {                                                                                                       \
        const double v = ((d)&1) ? fma(a[d], car[il1], fma(b[d], car[iu1], c[d]))                           \
                                 : fma(a[d], car[il0], fma(b[d], car[iu0], c[d]));                          \
        if ((d)&1) car[jw0] = v; else car[jw1] = v;                                                         \
        a[d] = reg_ldg(pA + po); b[d] = reg_ldg(pB + po); c[d] = reg_ldg(pC + po);                          \
        if ((tt) >= last0) reg_stg(pI + so, v);                                                             \
        po += stride;                                                                                       \
        if (po == o_end) po = o0;                                                                           \
        so += stride;                                                                                       \
        if (so == o_end) so = o0;                                                                           \
        __syncthreads();                                                                                    \
    }
    int t0 = 0;
    for (; t0 + PF <= total; t0 += PF) {
#pragma unroll
        for (int d = 0; d < PF; d++) REG_STEP(d, t0 + d)
    }
#pragma unroll
    for (int d = 0; d < PF; d++)
        if (t0 + d < total) REG_STEP(d, t0 + d)   // uniform over the CTA
#undef REG_STEP
    // periodic ghosts of the plane (:470-480): columns 0 / np-1 of the interior rows, then the two ghost rows.  The CTA
    // reads back its own stores; __syncthreads orders them.
    double* pl = Iout + base;
    for (int r = threadIdx.x; r < nsi; r += blockDim.x) {
        double* row = pl + (size_t)(r + 1) * np;
        row[0] = row[np - 2];
        row[np - 1] = row[1];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < np; q += blockDim.x) {
        pl[q] = pl[(size_t)(ns - 2) * np + q];
        pl[(size_t)(ns - 1) * np + q] = pl[np + q];
    }
}

// Device workspace of the regular-grid entries, kept between calls (a Λ-iteration solves n_directions times per
// iteration with the same shapes; cudaMalloc/cudaFree of tens of GB cost more than the solve).  Grow-only; released by
// vrt_regular_release_workspace().  S and alpha are held once per internal layout (j = x and j = y).
struct RegWorkspace {
    int device = -1;
    uint64_t token = 0;   // changes whenever an entry point writes the S / alpha buffers (see regular_dir_accumulate)
    DevBuf<double> dS[2], dA[2], dI, stage, stage0, cA, cB, cC, dx, dy, dJ;
    DevBuf<unsigned long long> diff_bits;
    DevBuf<int> diff_nan;
    size_t bytes() const {
        return 8 * (dS[0].n + dS[1].n + dA[0].n + dA[1].n + dI.n + stage.n + stage0.n + cA.n + cB.n + cC.n + dx.n + dy.n + dJ.n);
    }
    void release_arrays() {   // before a grow: the large buffers only
        token++;
        for (int i = 0; i < 2; i++) { dS[i].release(); dA[i].release(); }
        dI.release(); stage.release(); stage0.release(); cA.release(); cB.release(); cC.release(); dJ.release();
    }
    void release() {
        release_arrays();
        dx.release(); dy.release(); diff_bits.release(); diff_nan.release();
    }
};
std::mutex g_reg_mu;
RegWorkspace* g_reg_ws = nullptr;   // never destroyed at exit: the CUDA runtime may already be gone by then

struct RegGeom {
    int64_t nz = 0, nx = 0, ny = 0;
    size_t plane = 0, vol = 0;
    std::vector<double> hz, hx, hy;
};

struct RegDir {
    RegPlane P;
    double k[3];
    int down;
    double r_x, r_y;
};

int reg_check_geom(const char* who, int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, RegGeom* G) {
    if (!z || !x || !y || nz < 2 || nx < 3 || ny < 3) {
        set_error("%s: bad grid arguments", who);
        return VRT_E_INVALID;
    }
    if (nx > 1026 || ny > 1026) {
        set_error("%s: nx, ny <= 1026 (one thread per interior point of a row in the recurrence)", who);
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("%s: no CUDA device (this library has no CPU path)", who);
        return VRT_E_CUDA;
    }
    G->nz = nz; G->nx = nx; G->ny = ny;
    G->plane = (size_t)nx * ny;
    G->vol = G->plane * nz;
    G->hz.resize(nz); G->hx.resize(nx); G->hy.resize(ny);
    VRT_CUDA(cudaMemcpy(G->hz.data(), z, sizeof(double) * nz, cudaMemcpyDefault));
    VRT_CUDA(cudaMemcpy(G->hx.data(), x, sizeof(double) * nx, cudaMemcpyDefault));
    VRT_CUDA(cudaMemcpy(G->hy.data(), y, sizeof(double) * ny, cudaMemcpyDefault));
    return VRT_OK;
}

// characteristics.jl:34-40: path lengths to the x and y faces, xy_intersect signs, and the internal layout they imply
int reg_dir_setup(const char* who, const RegGeom& G, const double k[3], int down, RegDir* D) {
    if (!(k[0] == k[0]) || !(k[1] == k[1]) || !(k[2] == k[2]) || k[0] == 0.0) {
        set_error("%s: direction must be finite with k[0] != 0", who);
        return VRT_E_INVALID;
    }
    const double dx = G.hx[1] - G.hx[0], dy = G.hy[1] - G.hy[0];
    D->r_x = fabs(dx / k[1]);
    D->r_y = fabs(dy / k[2]);
    int sx = 1, sy = 1;
    if (k[1] > 0 && k[2] > 0) { sx = -1; sy = -1; }
    else if (k[1] < 0 && k[2] > 0) { sx = 1; sy = -1; }
    else if (k[1] < 0 && k[2] < 0) { sx = 1; sy = 1; }
    else if (k[1] > 0 && k[2] < 0) { sx = -1; sy = 1; }
    // argmin([r_z, r_x, r_y]) takes the first minimum: a sideways plane is xz only when r_y < r_x strictly
    RegPlane& P = D->P;
    P.par_is_x = (D->r_y < D->r_x) ? 1 : 0;
    P.np = (int)(P.par_is_x ? G.nx : G.ny);
    P.ns = (int)(P.par_is_x ? G.ny : G.nx);
    P.sgn_j = P.par_is_x ? sx : sy;
    P.sgn_s = P.par_is_x ? sy : sx;
    P.kz = k[0];
    P.kj = P.par_is_x ? k[1] : k[2];
    P.ks = P.par_is_x ? k[2] : k[1];
    P.lc = 0;
    D->k[0] = k[0]; D->k[1] = k[1]; D->k[2] = k[2];
    D->down = down ? 1 : 0;
    return VRT_OK;
}

int reg_workspace(RegWorkspace** out) {
    if (!g_reg_ws) g_reg_ws = new RegWorkspace();
    RegWorkspace& W = *g_reg_ws;
    int cur_dev = 0;
    VRT_CUDA(cudaGetDevice(&cur_dev));
    if (W.device != cur_dev) { W.release(); W.device = cur_dev; }
    *out = &W;
    return VRT_OK;
}

// wavelengths per chunk when `vols` internal volumes and `planes` planes are needed per wavelength
int reg_plan_chunk(const char* who, const RegGeom& G, const RegWorkspace& W, int64_t nlam, double vols, double planes, int64_t* lc_out) {
    size_t free_b = 0, total_b = 0;
    VRT_CUDA(cudaMemGetInfo(&free_b, &total_b));
    free_b += W.bytes();   // what the workspace already holds is ours to reuse
    const double per_lam = 8.0 * (vols * G.vol + planes * G.plane);
    int64_t lc = (int64_t)std::min<double>((double)nlam, floor(0.9 * (double)free_b / per_lam));
    if (const char* e = getenv("VRT_REG_LAM_CHUNK")) lc = std::max<int64_t>(1, std::min<int64_t>(lc, atoll(e)));
    if (lc < 1) {
        set_error("%s: one wavelength of a %lld x %lld x %lld grid needs %.1f GB, %.1f GB free", who, (long long)G.nz,
                  (long long)G.nx, (long long)G.ny, per_lam / 1e9, free_b / 1e9);
        return VRT_E_NOMEM;
    }
    *lc_out = std::min<int64_t>(lc, 65535);
    return VRT_OK;
}

int reg_common_buffers(const RegGeom& G, RegWorkspace& W, int64_t lc) {
    const size_t need_pl = G.plane * lc;
    VRT_TRY(W.cA.ensure(need_pl)); VRT_TRY(W.cB.ensure(need_pl)); VRT_TRY(W.cC.ensure(need_pl));
    VRT_TRY(W.dx.ensure(G.nx)); VRT_TRY(W.dy.ensure(G.ny));
    VRT_CUDA(cudaMemcpy(W.dx.p, G.hx.data(), sizeof(double) * G.nx, cudaMemcpyHostToDevice));
    VRT_CUDA(cudaMemcpy(W.dy.p, G.hy.data(), sizeof(double) * G.ny, cudaMemcpyHostToDevice));
    return VRT_OK;
}

// user array (host or device; nlam wavelengths, nz_ planes) -> internal layout of wavelengths [l0, l0+n_l)
int reg_load(const RegGeom& G, const double* src, bool on_dev, int64_t nz_, int64_t nlam, int64_t l0, int64_t n_l, int par_is_x,
             double* dst, DevBuf<double>& stg, SweepStats* st) {
    const double* s_dev = src;
    int64_t ld = nlam, lo = l0;
    const size_t rows = (size_t)nz_ * G.plane;
    if (!on_dev) {
        VRT_TRY(stg.ensure(rows * n_l));
        if (n_l == nlam) VRT_CUDA(cudaMemcpy(stg.p, src, sizeof(double) * rows * nlam, cudaMemcpyHostToDevice));
        else VRT_CUDA(cudaMemcpy2D(stg.p, sizeof(double) * n_l, src + l0, sizeof(double) * nlam, sizeof(double) * n_l, rows, cudaMemcpyHostToDevice));
        s_dev = stg.p; ld = n_l; lo = 0;
    }
    const int np = (int)(par_is_x ? G.nx : G.ny), ns = (int)(par_is_x ? G.ny : G.nx);
    const dim3 grid((unsigned)((n_l * nz_ + 31) / 32), (unsigned)((np + 31) / 32), (unsigned)ns);
    k_reg_transpose<1><<<grid, dim3(32, 8)>>>(s_dev, dst, ld, lo, (int)n_l, nz_, G.nx, G.ny, par_is_x, 1.0, 0);
    VRT_CUDA(cudaGetLastError());
    if (!on_dev) VRT_CUDA(cudaDeviceSynchronize());   // `stg` may be reused by the next load
    st->kernels += 1;
    return VRT_OK;
}

// internal layout -> user array (host or device): dst = w * I, or dst += w * I
int reg_store(const RegGeom& G, const double* dI, double* dst, bool on_dev, int64_t nlam, int64_t l0, int64_t n_l, int par_is_x,
              double w, int accumulate, DevBuf<double>& stg, SweepStats* st) {
    double* d_dst = dst;
    int64_t ld = nlam, lo = l0;
    if (!on_dev) {
        if (accumulate) {
            set_error("internal: accumulating store needs a device destination");
            return VRT_E_STATE;
        }
        VRT_TRY(stg.ensure(G.vol * n_l));
        d_dst = stg.p; ld = n_l; lo = 0;
    }
    const int np = (int)(par_is_x ? G.nx : G.ny), ns = (int)(par_is_x ? G.ny : G.nx);
    const dim3 grid((unsigned)((n_l * G.nz + 31) / 32), (unsigned)((np + 31) / 32), (unsigned)ns);
    k_reg_transpose<0><<<grid, dim3(32, 8)>>>(dI, d_dst, ld, lo, (int)n_l, G.nz, G.nx, G.ny, par_is_x, w, accumulate);
    VRT_CUDA(cudaGetLastError());
    st->kernels += 1;
    if (!on_dev) {
        if (n_l == nlam) VRT_CUDA(cudaMemcpy(dst, stg.p, sizeof(double) * G.vol * nlam, cudaMemcpyDeviceToHost));
        else VRT_CUDA(cudaMemcpy2D(dst + l0, sizeof(double) * nlam, stg.p, sizeof(double) * n_l, sizeof(double) * n_l, G.vol, cudaMemcpyDeviceToHost));
    }
    return VRT_OK;
}

// The walk over the planes (characteristics.jl:47-92 / :135-177).  dS, dA: internal S and alpha of this chunk in the
// layout of D; dI: internal I with the boundary plane already in place.  branch (optional, host, nz) gets the routine per plane.
int reg_plane_loop(const RegGeom& G, RegWorkspace& W, RegDir D, int n_sweeps, int64_t n_l, const double* dS, const double* dA,
                   double* dI, SweepStats* st, int32_t* branch) {
    RegPlane& P = D.P;
    P.lc = (int)n_l;
    const size_t pst = G.plane * n_l;   // plane stride of the internal arrays
    const double* dcj = P.par_is_x ? W.dx.p : W.dy.p;
    const double* dcs = P.par_is_x ? W.dy.p : W.dx.p;
    const int64_t nz = G.nz;
    const std::vector<double>& hz = G.hz;
    if (branch) branch[D.down ? nz - 1 : 0] = 0;
    for (int64_t step = 1; step < nz; step++) {
        const int64_t idz = D.down ? nz - 1 - step : step;      // 0-based plane being solved
        const int64_t idu = D.down ? idz + 1 : idz - 1;         // upwind plane
        const double dz = D.down ? hz[idz + 1] - hz[idz] : hz[idz] - hz[idz - 1];
        const double r_z = fabs(dz / D.k[0]);
        int cut = 1;
        double best = r_z;
        if (D.r_x < best) { best = D.r_x; cut = 2; }
        if (D.r_y < best) { best = D.r_y; cut = 3; }
        if (branch) branch[idz] = cut;
        double* Ic = dI + pst * idz;
        const double* Iu = dI + pst * idu;
        const dim3 grid((unsigned)((P.np * P.ns + 255) / 256), (unsigned)((n_l + REG_LPT - 1) / REG_LPT));
        if (cut == 1) {
            k_reg_xy<<<grid, 256>>>(P, r_z, dcj, dcs, dS + pst * idz, dA + pst * idz, dS + pst * idu, dA + pst * idu, Iu, Ic);
            st->kernels += 1;
        } else {
            const int up = D.down ? 0 : 1;
            const int64_t izl = up ? idz - 1 : idz, izu = izl + 1;
            const int64_t izc = P.par_is_x ? izu : idz;          // Q13: xz takes the centre from the upper plane
            k_reg_coef<<<grid, 256>>>(P, up, P.par_is_x ? D.r_y : D.r_x, hz[idz], hz[izl], hz[izu], dcj, dS + pst * izl, dS + pst * izu,
                                      dA + pst * izl, dA + pst * izu, dS + pst * izc, dA + pst * izc, Iu, W.cA.p, W.cB.p, W.cC.p);
            const int threads = ((P.np - 2 + 31) / 32) * 32;
            const size_t rec_smem = sizeof(double) * 2 * P.np;
            if (threads <= 512) k_reg_rec<512, REG_PF><<<(unsigned)n_l, threads, rec_smem>>>(P, n_sweeps, W.cA.p, W.cB.p, W.cC.p, Ic);
            else k_reg_rec<1024, REG_PF / 3><<<(unsigned)n_l, threads, rec_smem>>>(P, n_sweeps, W.cA.p, W.cB.p, W.cC.p, Ic);
            st->kernels += 2;
            st->steps += (double)n_sweeps * (P.ns - 2);
        }
        VRT_CUDA(cudaGetLastError());
    }
    return VRT_OK;
}

struct EvPair {
    cudaEvent_t a = nullptr, b = nullptr;
    int create() {
        VRT_CUDA(cudaEventCreate(&a));
        VRT_CUDA(cudaEventCreate(&b));
        return VRT_OK;
    }
    ~EvPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

// I0 plane of one chunk: a user array (nlam x nx x ny), or zero when NULL
int reg_boundary(const RegGeom& G, RegWorkspace& W, const double* I0, int64_t nlam, int64_t l0, int64_t n_l, int par_is_x, int down,
                 SweepStats* st) {
    double* pl = W.dI.p + G.plane * n_l * (down ? G.nz - 1 : 0);
    if (!I0) {
        VRT_CUDA(cudaMemsetAsync(pl, 0, sizeof(double) * G.plane * n_l));
        return VRT_OK;
    }
    return reg_load(G, I0, is_device_ptr(I0), 1, nlam, l0, n_l, par_is_x, pl, W.stage0, st);
}

// J_λ_regular (lambda_continuum.jl:1-24; lambda_iteration.jl:23-55 with a direction-independent alpha): J = Σ w_i I_i over
// the quadrature, up rays (θ > 90) starting from I0_up at z[0], down rays (θ < 90) from I0_down (zero when NULL) at
// z[nz-1].  S, alpha, J: device or host; when `alpha_ready` is given, the internal copies of alpha are kept across calls
// (alpha_ready[layout] says which are valid) — the Λ-iteration's alpha does not change.  Caller holds g_reg_mu.
int reg_mean_intensity(const char* who, const RegGeom& G, RegWorkspace& W, const vrt_quadrature* q, int n_sweeps, int64_t nlam,
                       const double* S, const double* alpha, const double* I0_up, const double* I0_down, double* J,
                       bool* alpha_ready, SweepStats* st) {
    const bool dev_S = is_device_ptr(S), dev_a = is_device_ptr(alpha), dev_J = is_device_ptr(J);
    const bool need_stage = !(dev_S && dev_a);
    int64_t lc = 0;
    VRT_TRY(reg_plan_chunk(who, G, W, nlam, 5.0 + (need_stage ? 1.0 : 0.0) + (dev_J ? 0.0 : 1.0), 4.0, &lc));
    if (alpha_ready && lc < nlam) alpha_ready = nullptr;   // chunked: alpha is re-laid out per chunk
    {
        const size_t need_vol = G.vol * lc;
        bool grow = W.dI.n < need_vol || (need_stage && W.stage.n < need_vol) || (!dev_J && W.dJ.n < need_vol);
        for (int i = 0; i < 2; i++) grow = grow || (W.dS[i].p && W.dS[i].n < need_vol) || (W.dA[i].p && W.dA[i].n < need_vol);
        if (grow) {
            W.release_arrays();
            if (alpha_ready) alpha_ready[0] = alpha_ready[1] = false;
        }
        VRT_TRY(W.dI.ensure(need_vol));
        if (!dev_J) VRT_TRY(W.dJ.ensure(need_vol));
        VRT_TRY(reg_common_buffers(G, W, lc));
    }
    EvPair ev;
    VRT_TRY(ev.create());
    W.token++;   // the S / alpha buffers are about to be overwritten
    for (int64_t l0 = 0; l0 < nlam; l0 += lc) {
        const int64_t n_l = std::min<int64_t>(lc, nlam - l0);
        bool have_S[2] = {false, false}, have_a_local[2] = {false, false};
        bool* have_a = alpha_ready ? alpha_ready : have_a_local;
        // J of this chunk accumulates on the device: in the caller's array, or in dJ (chunk-contiguous) for a host J
        double* Jd = dev_J ? J : W.dJ.p;
        const int64_t J_ld = dev_J ? nlam : n_l, J_l0 = dev_J ? l0 : 0;
        int n_acc = 0;
        for (int64_t i = 0; i < q->n_dirs; i++) {
            const double th = q->theta[i], ph = q->phi[i];
            if (!(th > 90) && !(th < 90)) continue;   // lambda_continuum.jl:15-21: θ = 90 belongs to neither branch
            const int down = th < 90 ? 1 : 0;
            const double k[3] = {cos(th * M_PI / 180), cos(ph * M_PI / 180) * sin(th * M_PI / 180), sin(ph * M_PI / 180) * sin(th * M_PI / 180)};
            RegDir D;
            VRT_TRY(reg_dir_setup(who, G, k, down, &D));
            const int lay = D.P.par_is_x;
            if (!have_S[lay]) {
                VRT_TRY(W.dS[lay].ensure(G.vol * lc));
                VRT_TRY(reg_load(G, S, dev_S, G.nz, nlam, l0, n_l, lay, W.dS[lay].p, W.stage, st));
                have_S[lay] = true;
            }
            if (!have_a[lay]) {
                VRT_TRY(W.dA[lay].ensure(G.vol * lc));
                VRT_TRY(reg_load(G, alpha, dev_a, G.nz, nlam, l0, n_l, lay, W.dA[lay].p, W.stage, st));
                have_a[lay] = true;
            }
            VRT_TRY(reg_boundary(G, W, down ? I0_down : I0_up, nlam, l0, n_l, lay, down, st));
            VRT_CUDA(cudaEventRecord(ev.a));
            VRT_TRY(reg_plane_loop(G, W, D, n_sweeps, n_l, W.dS[lay].p, W.dA[lay].p, W.dI.p, st, nullptr));
            VRT_CUDA(cudaEventRecord(ev.b));
            VRT_TRY(reg_store(G, W.dI.p, Jd, true, J_ld, J_l0, n_l, lay, q->weights[i], n_acc > 0, W.stage, st));
            n_acc++;
            VRT_CUDA(cudaDeviceSynchronize());
            float ms = 0;
            VRT_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
            st->sweep_ms += ms;
            st->visits += (double)(G.nz - 1) * (double)(G.nx - 2) * (double)(G.ny - 2) * (double)n_l;
        }
        if (n_acc == 0) {   // no direction contributes: J = zero(S)
            if (dev_J && n_l == nlam) VRT_CUDA(cudaMemset(J, 0, sizeof(double) * G.vol * nlam));
            else if (dev_J) VRT_CUDA(cudaMemset2D(J + l0, sizeof(double) * nlam, 0, sizeof(double) * n_l, G.vol));
            else VRT_CUDA(cudaMemset(W.dJ.p, 0, sizeof(double) * G.vol * n_l));
        }
        if (!dev_J) {
            if (n_l == nlam) VRT_CUDA(cudaMemcpy(J, W.dJ.p, sizeof(double) * G.vol * nlam, cudaMemcpyDeviceToHost));
            else VRT_CUDA(cudaMemcpy2D(J + l0, sizeof(double) * nlam, W.dJ.p, sizeof(double) * n_l, sizeof(double) * n_l, G.vol, cudaMemcpyDeviceToHost));
        }
    }
    return VRT_OK;
}

// plane iz of a (nz, nx, ny) column-major array -> contiguous (nx, ny)
__global__ void k_reg_take_plane(const double* __restrict__ a, int64_t nz, int64_t iz, int64_t plane, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < plane) out[i] = a[iz + nz * i];
}

// host copy of a quadrature table whose arrays may live on the device
struct RegQuad {
    std::vector<double> w, t, p;
    vrt_quadrature q;
    int init(const vrt_quadrature* in) {
        const size_t n = (size_t)in->n_dirs;
        w.resize(n); t.resize(n); p.resize(n);
        VRT_CUDA(cudaMemcpy(w.data(), in->weights, sizeof(double) * n, cudaMemcpyDefault));
        VRT_CUDA(cudaMemcpy(t.data(), in->theta, sizeof(double) * n, cudaMemcpyDefault));
        VRT_CUDA(cudaMemcpy(p.data(), in->phi, sizeof(double) * n, cudaMemcpyDefault));
        q.n_dirs = in->n_dirs; q.weights = w.data(); q.theta = t.data(); q.phi = p.data();
        return VRT_OK;
    }
};

int reg_read_diff(RegWorkspace& W, double* diff) {
    unsigned long long bits = 0;
    int isn = 0;
    VRT_CUDA(cudaMemcpy(&bits, W.diff_bits.p, sizeof(bits), cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaMemcpy(&isn, W.diff_nan.p, sizeof(isn), cudaMemcpyDeviceToHost));
    double d;
    memcpy(&d, &bits, sizeof(d));
    *diff = isn ? NAN : d;   // Julia's maximum propagates NaN; `NaN > ϵ` is false and ends the loop
    return VRT_OK;
}


int reg_geom_of(const vrt_grid* g, RegGeom* G) {
    if (!g || !g->regular) {
        set_error("internal: not a regular grid");
        return VRT_E_STATE;
    }
    G->nz = g->rnz; G->nx = g->rnx; G->ny = g->rny;
    G->plane = (size_t)g->rnx * g->rny;
    G->vol = G->plane * g->rnz;
    G->hz = g->rz; G->hx = g->rx; G->hy = g->ry;
    return VRT_OK;
}

}  // namespace

int regular_plan_chunk(const vrt_grid* g, int64_t nlam, double extra_vols, int64_t* lc) {
    RegGeom G;
    VRT_TRY(reg_geom_of(g, &G));
    std::lock_guard<std::mutex> lock(g_reg_mu);
    RegWorkspace* Wp = nullptr;
    VRT_TRY(reg_workspace(&Wp));
    // internal: S, alpha and I of the direction being solved (regular_dir_accumulate holds one layout at a time)
    return reg_plan_chunk("regular solver", G, *Wp, nlam, 3.0 + extra_vols, 4.0, lc);
}

int regular_dir_layout(const vrt_grid* g, const double k[3], int* layout) {
    RegGeom G;
    VRT_TRY(reg_geom_of(g, &G));
    RegDir D;
    VRT_TRY(reg_dir_setup("regular solver", G, k, 0, &D));
    *layout = D.P.par_is_x;
    return VRT_OK;
}

int regular_dir_accumulate(const vrt_grid* g, const double k[3], int down, int n_sweeps, int64_t n_l, const double* S, int64_t S_ld,
                           int64_t S_l0, const double* alpha, int64_t a_ld, int64_t a_l0, const double* I0, double* J, int64_t J_ld,
                           int64_t J_l0, double w, int accumulate, bool have_S[2], uint64_t* token, SweepStats* st) {
    const char* who = "regular solver";
    RegGeom G;
    VRT_TRY(reg_geom_of(g, &G));
    RegDir D;
    VRT_TRY(reg_dir_setup(who, G, k, down, &D));
    const int lay = D.P.par_is_x;
    std::lock_guard<std::mutex> lock(g_reg_mu);
    RegWorkspace* Wp = nullptr;
    VRT_TRY(reg_workspace(&Wp));
    RegWorkspace& W = *Wp;
    // one buffer each for S and alpha (slot 0), whatever the layout: the caller groups its directions by layout, so S is
    // laid out twice per wavelength chunk, and the second slot's 2 volumes buy wider chunks instead
    const size_t need_vol = G.vol * n_l;
    // the workspace is shared by every regular-grid entry of the process: S laid out by an earlier call of this function
    // is only trusted when nobody wrote the buffers since (the caller carries the token from direction to direction)
    if (*token != W.token) have_S[0] = have_S[1] = false;
    if (W.dI.n < need_vol || W.dS[0].n < need_vol || W.dA[0].n < need_vol || W.dS[1].p || W.dA[1].p) {
        W.release_arrays();
        have_S[0] = have_S[1] = false;
    }
    VRT_TRY(W.dI.ensure(need_vol));
    VRT_TRY(W.dS[0].ensure(need_vol));
    VRT_TRY(W.dA[0].ensure(need_vol));
    VRT_TRY(reg_common_buffers(G, W, n_l));
    if (!have_S[lay]) {
        VRT_TRY(reg_load(G, S, true, G.nz, S_ld, S_l0, n_l, lay, W.dS[0].p, W.stage, st));
        have_S[lay] = true;
        have_S[1 - lay] = false;
        W.token++;
    }
    *token = W.token;
    VRT_TRY(reg_load(G, alpha, true, G.nz, a_ld, a_l0, n_l, lay, W.dA[0].p, W.stage, st));
    double* pl = W.dI.p + G.plane * n_l * (down ? G.nz - 1 : 0);
    if (I0) VRT_TRY(reg_load(G, I0, true, 1, n_l, 0, n_l, lay, pl, W.stage0, st));
    else VRT_CUDA(cudaMemsetAsync(pl, 0, sizeof(double) * G.plane * n_l));
    EvPair ev;
    VRT_TRY(ev.create());
    VRT_CUDA(cudaEventRecord(ev.a));
    VRT_TRY(reg_plane_loop(G, W, D, n_sweeps, n_l, W.dS[0].p, W.dA[0].p, W.dI.p, st, nullptr));
    VRT_CUDA(cudaEventRecord(ev.b));
    VRT_TRY(reg_store(G, W.dI.p, J, true, J_ld, J_l0, n_l, lay, w, accumulate, W.stage, st));
    VRT_CUDA(cudaDeviceSynchronize());
    float ms = 0;
    VRT_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    st->sweep_ms += ms;
    st->visits += (double)(G.nz - 1) * (double)(G.nx - 2) * (double)(G.ny - 2) * (double)n_l;
    return VRT_OK;
}

}  // namespace vrt

using namespace vrt;

/* a vrt_grid over the regular Cartesian atmosphere: cell c = iz + nz*(ix + nx*iy), the memory order of the reference's
 * (nz, nx, ny) arrays */
extern "C" int vrt_regular_grid_create(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                                       vrt_grid** out) {
    if (!out) return VRT_E_INVALID;
    *out = nullptr;
    RegGeom G;
    VRT_TRY(reg_check_geom("vrt_regular_grid_create", nz, nx, ny, z, x, y, &G));
    if (G.vol >= ((size_t)1 << 31)) {
        set_error("vrt_regular_grid_create: more than 2^31 cells");
        return VRT_E_INVALID;
    }
    vrt_grid* g = new vrt_grid();
    struct Guard { vrt_grid* g; ~Guard() { delete g; } } guard{g};
    g->regular = true;
    g->n = (int64_t)G.vol;
    g->rnz = nz; g->rnx = nx; g->rny = ny;
    g->rz = G.hz; g->rx = G.hx; g->ry = G.hy;
    g->bounds[0] = G.hz.front(); g->bounds[1] = G.hz.back();
    g->bounds[2] = G.hx.front(); g->bounds[3] = G.hx.back();
    g->bounds[4] = G.hy.front(); g->bounds[5] = G.hy.back();
    VRT_CUDA(cudaGetDevice(&g->device));
    // identity maps: the generic host<->internal copies of solver.cu then work unchanged
    std::vector<int32_t> id((size_t)g->n);
    for (int64_t i = 0; i < g->n; i++) id[(size_t)i] = (int32_t)i;
    VRT_TRY(g->site_of.alloc((size_t)g->n));
    VRT_TRY(g->rank_of.alloc((size_t)g->n));
    VRT_CUDA(cudaMemcpy(g->site_of.p, id.data(), sizeof(int32_t) * id.size(), cudaMemcpyHostToDevice));
    VRT_CUDA(cudaMemcpy(g->rank_of.p, id.data(), sizeof(int32_t) * id.size(), cudaMemcpyHostToDevice));
    g->off_up = {1, 1};
    g->off_down = {1, 1};
    guard.g = nullptr;
    *out = g;
    return VRT_OK;
}

extern "C" int vrt_regular_formal_solve(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                                        const double k[3], int32_t down, int32_t n_sweeps, int64_t nlam, const double* S,
                                        const double* alpha, const double* I0, double* I_out, int32_t* plane_branch) {
    const char* who = "vrt_regular_formal_solve";
    if (!k || !S || !alpha || !I0 || !I_out || nlam <= 0 || n_sweeps < 1) {
        set_error("%s: bad arguments", who);
        return VRT_E_INVALID;
    }
    RegGeom G;
    VRT_TRY(reg_check_geom(who, nz, nx, ny, z, x, y, &G));
    RegDir D;
    VRT_TRY(reg_dir_setup(who, G, k, down, &D));
    const int lay = D.P.par_is_x;
    const bool dev_S = is_device_ptr(S), dev_a = is_device_ptr(alpha), dev_out = is_device_ptr(I_out);
    const bool need_stage = !(dev_S && dev_a && dev_out);
    std::lock_guard<std::mutex> lock(g_reg_mu);
    RegWorkspace* Wp = nullptr;
    VRT_TRY(reg_workspace(&Wp));
    RegWorkspace& W = *Wp;
    int64_t lc = 0;
    VRT_TRY(reg_plan_chunk(who, G, W, nlam, 3.0 + (need_stage ? 1.0 : 0.0), 4.0, &lc));
    {
        const size_t need_vol = G.vol * lc;
        // a grow of one buffer must not fail because the others hold stale, larger-than-needed space
        const bool grow = W.dS[lay].n < need_vol || W.dA[lay].n < need_vol || W.dI.n < need_vol || (need_stage && W.stage.n < need_vol);
        if (grow) W.release_arrays();
        VRT_TRY(W.dS[lay].ensure(need_vol)); VRT_TRY(W.dA[lay].ensure(need_vol)); VRT_TRY(W.dI.ensure(need_vol));
        VRT_TRY(reg_common_buffers(G, W, lc));
    }
    SweepStats stats;
    std::vector<int32_t> branch(nz, 0);
    EvPair ev;
    VRT_TRY(ev.create());
    W.token++;   // the S / alpha buffers are about to be overwritten
    for (int64_t l0 = 0; l0 < nlam; l0 += lc) {
        const int64_t n_l = std::min<int64_t>(lc, nlam - l0);
        VRT_TRY(reg_load(G, S, dev_S, nz, nlam, l0, n_l, lay, W.dS[lay].p, W.stage, &stats));
        VRT_TRY(reg_load(G, alpha, dev_a, nz, nlam, l0, n_l, lay, W.dA[lay].p, W.stage, &stats));
        // I[zb, :, :] = I_0 (characteristics.jl:45 / :133), ghost columns as the caller filled them
        VRT_TRY(reg_boundary(G, W, I0, nlam, l0, n_l, lay, D.down, &stats));
        VRT_CUDA(cudaEventRecord(ev.a));
        VRT_TRY(reg_plane_loop(G, W, D, n_sweeps, n_l, W.dS[lay].p, W.dA[lay].p, W.dI.p, &stats, branch.data()));
        VRT_CUDA(cudaEventRecord(ev.b));
        VRT_TRY(reg_store(G, W.dI.p, I_out, dev_out, nlam, l0, n_l, lay, 1.0, 0, W.stage, &stats));
        VRT_CUDA(cudaDeviceSynchronize());
        float ms = 0;
        VRT_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
        stats.sweep_ms += ms;
    }
    stats.visits = (double)(nz - 1) * (double)(nx - 2) * (double)(ny - 2) * (double)nlam;
    g_last_stats = stats;
    if (plane_branch) VRT_CUDA(cudaMemcpy(plane_branch, branch.data(), sizeof(int32_t) * nz, cudaMemcpyDefault));
    return VRT_OK;
}

extern "C" int vrt_regular_mean_intensity(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                                          const vrt_quadrature* quad, int32_t n_sweeps, int64_t nlam, const double* S,
                                          const double* alpha, const double* I0_up, const double* I0_down, double* J) {
    const char* who = "vrt_regular_mean_intensity";
    if (!quad || quad->n_dirs <= 0 || !quad->weights || !quad->theta || !quad->phi || !S || !alpha || !J || nlam <= 0 || n_sweeps < 1) {
        set_error("%s: bad arguments", who);
        return VRT_E_INVALID;
    }
    RegGeom G;
    VRT_TRY(reg_check_geom(who, nz, nx, ny, z, x, y, &G));
    std::lock_guard<std::mutex> lock(g_reg_mu);
    RegWorkspace* Wp = nullptr;
    VRT_TRY(reg_workspace(&Wp));
    SweepStats stats;
    RegQuad hq;
    VRT_TRY(hq.init(quad));
    VRT_TRY(reg_mean_intensity(who, G, *Wp, &hq.q, n_sweeps, nlam, S, alpha, I0_up, I0_down, J, nullptr, &stats));
    g_last_stats = stats;
    return VRT_OK;
}

extern "C" int vrt_regular_lambda_iterate(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                                          const vrt_quadrature* quad, int32_t n_sweeps, const double* alpha, const double* eps_l,
                                          const double* B0, double eps, int32_t maxiter, vrt_iter_cb cb, void* user, double* S_out,
                                          double* J_out, vrt_result* out) {
    const char* who = "vrt_regular_lambda_iterate";
    if (!quad || quad->n_dirs <= 0 || !quad->weights || !quad->theta || !quad->phi || !alpha || !eps_l || !B0 || n_sweeps < 1) {
        set_error("%s: bad arguments", who);
        return VRT_E_INVALID;
    }
    auto t_begin = std::chrono::steady_clock::now();
    RegGeom G;
    VRT_TRY(reg_check_geom(who, nz, nx, ny, z, x, y, &G));
    std::lock_guard<std::mutex> lock(g_reg_mu);
    RegWorkspace* Wp = nullptr;
    VRT_TRY(reg_workspace(&Wp));
    RegWorkspace& W = *Wp;
    RegQuad hq;
    VRT_TRY(hq.init(quad));
    const int64_t n = (int64_t)G.vol;
    // state in the caller's layout, resident for the whole loop
    DevBuf<double> dS, dJ, dal, dep, dB, dI0;
    VRT_TRY(dS.alloc(n)); VRT_TRY(dJ.alloc(n)); VRT_TRY(dal.alloc(n)); VRT_TRY(dep.alloc(n)); VRT_TRY(dB.alloc(n)); VRT_TRY(dI0.alloc(G.plane));
    VRT_TRY(copy_in(dal.p, alpha, sizeof(double) * n)); VRT_TRY(copy_in(dep.p, eps_l, sizeof(double) * n)); VRT_TRY(copy_in(dB.p, B0, sizeof(double) * n));
    VRT_TRY(W.diff_bits.ensure(1)); VRT_TRY(W.diff_nan.ensure(1));
    // S_new = B_0 (lambda_continuum.jl:84-85); bottom boundary blackbody_λ(500 nm, T[1,:,:]) = B_0[1,:,:] (:16)
    VRT_CUDA(cudaMemcpy(dS.p, dB.p, sizeof(double) * n, cudaMemcpyDeviceToDevice));
    VRT_CUDA(cudaMemset(dJ.p, 0, sizeof(double) * n));
    k_reg_take_plane<<<(unsigned)((G.plane + 255) / 256), 256>>>(dB.p, nz, 0, (int64_t)G.plane, dI0.p);
    VRT_CUDA(cudaGetLastError());
    // first criterion: S_old = zero(S_new) (:87), over `thick` = ε > 1e-4 (:81)
    VRT_CUDA(cudaMemset(W.diff_bits.p, 0, sizeof(unsigned long long)));
    VRT_CUDA(cudaMemset(W.diff_nan.p, 0, sizeof(int)));
    VRT_TRY(continuum_criterion(n, dS.p, nullptr, dep.p, W.diff_bits.p, W.diff_nan.p));
    double diff = 0;
    VRT_TRY(reg_read_diff(W, &diff));
    int i = 0;
    bool alpha_ready[2] = {false, false};
    SweepStats total;
    while (diff > eps && i < maxiter) {
        auto t0 = std::chrono::steady_clock::now();
        SweepStats stats;
        vrt_iter_info info;
        memset(&info, 0, sizeof(info));
        info.diff = diff;
        VRT_TRY(reg_mean_intensity(who, G, W, &hq.q, n_sweeps, 1, dS.p, dal.p, dI0.p, nullptr, dJ.p, alpha_ready, &stats));
        info.t_sweep_ms = stats.sweep_ms;
        VRT_CUDA(cudaMemset(W.diff_bits.p, 0, sizeof(unsigned long long)));
        VRT_CUDA(cudaMemset(W.diff_nan.p, 0, sizeof(int)));
        VRT_TRY(continuum_source_update(n, dB.p, dep.p, dJ.p, dS.p, W.diff_bits.p, W.diff_nan.p));
        stats.kernels += 1;
        VRT_TRY(reg_read_diff(W, &diff));
        i++;
        info.iteration = i;
        info.updates = stats.visits;
        info.t_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        total.kernels += stats.kernels; total.visits += stats.visits; total.steps += stats.steps; total.sweep_ms += stats.sweep_ms;
        if (cb && cb(&info, user) != 0) break;
    }
    if (S_out) VRT_TRY(copy_out(S_out, dS.p, sizeof(double) * n));
    if (J_out) VRT_TRY(copy_out(J_out, dJ.p, sizeof(double) * n));
    VRT_CUDA(cudaDeviceSynchronize());
    g_last_stats = total;
    if (out) {
        out->iterations = i;
        out->converged = !(diff > eps);
        out->diff = diff;
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    }
    return VRT_OK;
}

extern "C" int vrt_regular_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_reg_mu);
    if (g_reg_ws) g_reg_ws->release();
    return VRT_OK;
}
