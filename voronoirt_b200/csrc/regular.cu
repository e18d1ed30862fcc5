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

// bilinear (functions.jl:328-355): the first coordinate selects the row of [Q11 Q12; Q21 Q22]
__device__ __forceinline__ double reg_bilinear(double xm, double ym, double x1, double x2, double y1, double y2,
                                               double Q11, double Q12, double Q21, double Q22) {
    const double dx = x2 - x1, dy = y2 - y1;
    const double f1 = ((x2 - xm) * Q11 + (xm - x1) * Q21) / dx;
    const double f2 = ((x2 - xm) * Q12 + (xm - x1) * Q22) / dx;
    return ((y2 - ym) * f1 + (ym - y1) * f2) / dy;
}

// ------------------------------------------------------------------ layout conversion
// src: caller's (nlam_src, nz, nx, ny) column-major, wavelengths [l0, l0+lc) -> dst [iz][l][s][j].
// Tile of 32 (q = l + lc*iz) x 32 (j) per block, one s per blockIdx.z; TO_INTERNAL = 0 is the inverse copy.
template <int TO_INTERNAL>
__global__ void k_reg_transpose(const double* __restrict__ src, double* __restrict__ dst, int64_t nlam_src, int64_t l0,
                                int lc, int64_t nz, int64_t nx, int64_t ny, int par_is_x) {
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
            if (j < np && q < Q) dst[user_off(q, j)] = tile[tx][ty + 8 * r];
        }
    }
}

// ------------------------------------------------------------------ xy branch
// xy_up_ray / xy_down_ray (characteristics.jl:191-280, :290-373).  One thread per (j, s, l) INCLUDING the ghost
// columns: a ghost evaluates the interior point it mirrors (:270-279), which is the same arithmetic, so no second pass.
// Sc/ac: plane idz; Su/au/Iu: upwind plane; cz = z[upwind] - z[idz].
__global__ void k_reg_xy(RegPlane P, double dz, const double* __restrict__ cj, const double* __restrict__ cs,
                         const double* __restrict__ Sc, const double* __restrict__ ac, const double* __restrict__ Su,
                         const double* __restrict__ au, const double* __restrict__ Iu, double* __restrict__ Iout) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // flattened (s, j): np need not be a multiple of the block
    const int l = blockIdx.y;
    if (idx >= P.np * P.ns) return;
    const int s = idx / P.np, j = idx - s * P.np;
    const int gj = reg_wrap(j, P.np), gs = reg_wrap(s, P.ns);
    const double r = fabs(dz / P.kz);
    const double jinc = r * P.kj, sinc = r * P.ks;
    const int jl = gj - (P.sgn_j + 1) / 2, ju = jl + 1;
    const int sl = gs - (P.sgn_s + 1) / 2, su = sl + 1;
    const double jup = cj[gj] + jinc, sup = cs[gs] + sinc;
    const double jb1 = cj[jl], jb2 = cj[ju], sb1 = cs[sl], sb2 = cs[su];
    const size_t base = (size_t)l * P.np * P.ns;
    const size_t oll = base + (size_t)sl * P.np + jl, olu = base + (size_t)su * P.np + jl;   // (j low, s low), (j low, s up)
    const size_t oul = oll + 1, ouu = olu + 1;
    const size_t oc = base + (size_t)gs * P.np + gj;
    double a_u, S_u, I_u;
    if (P.par_is_x) {   // j = x is the first coordinate of bilinear
        a_u = reg_bilinear(jup, sup, jb1, jb2, sb1, sb2, au[oll], au[olu], au[oul], au[ouu]);
        S_u = reg_bilinear(jup, sup, jb1, jb2, sb1, sb2, Su[oll], Su[olu], Su[oul], Su[ouu]);
        I_u = reg_bilinear(jup, sup, jb1, jb2, sb1, sb2, Iu[oll], Iu[olu], Iu[oul], Iu[ouu]);
    } else {            // s = x
        a_u = reg_bilinear(sup, jup, sb1, sb2, jb1, jb2, au[oll], au[oul], au[olu], au[ouu]);
        S_u = reg_bilinear(sup, jup, sb1, sb2, jb1, jb2, Su[oll], Su[oul], Su[olu], Su[ouu]);
        I_u = reg_bilinear(sup, jup, sb1, sb2, jb1, jb2, Iu[oll], Iu[oul], Iu[olu], Iu[ouu]);
    }
    const double dtau = r * (ac[oc] + a_u) / 2;
    double a, b, e;
    reg_weights(dtau, a, b, e);
    Iout[base + (size_t)s * P.np + j] = e * I_u + a * S_u + b * Sc[oc];
}

// ------------------------------------------------------------------ yz / xz branch, parallel part
// Coefficients of the row recurrence for every interior point of plane idz (yz_*_ray :383-604, xz_*_ray :614-835).
// lo/hi: planes izl/izu of S and alpha; prev: the previous (upwind) plane of I; zc = z[idz], zb1 = z[izl], zb2 = z[izu].
// Centre values come from plane idz in the yz branch and from the UPPER plane izu in the xz branch.
__global__ void k_reg_coef(RegPlane P, int up, double zc, double zb1, double zb2, const double* __restrict__ cj,
                           const double* __restrict__ cs, const double* __restrict__ S_lo, const double* __restrict__ S_hi,
                           const double* __restrict__ a_lo, const double* __restrict__ a_hi, const double* __restrict__ S_c,
                           const double* __restrict__ a_c, const double* __restrict__ prev, double* __restrict__ cA,
                           double* __restrict__ cB, double* __restrict__ cC) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (idx >= P.np * P.ns) return;
    const int s = idx / P.np, j = idx - s * P.np;
    if (j < 1 || j > P.np - 2 || s < 1 || s > P.ns - 2) return;
    const double r = fabs((cs[1] - cs[0]) / P.ks);
    const double zup = zc + r * P.kz;
    const double jup = cj[j] + r * P.kj;
    const int jl = j - (P.sgn_j + 1) / 2, ju = jl + 1;
    const int sn = s + P.sgn_s;
    const double jb1 = cj[jl], jb2 = cj[ju];
    const size_t base = (size_t)l * P.np * P.ns;
    const size_t ol = base + (size_t)sn * P.np + jl, ou = ol + 1, oc = base + (size_t)s * P.np + j;
    const double a_u = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, a_lo[ol], a_lo[ou], a_hi[ol], a_hi[ou]);
    const double S_u = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, S_lo[ol], S_lo[ou], S_hi[ol], S_hi[ou]);
    const double dtau = r * (a_c[oc] + a_u) / 2;
    double a, b, e;
    reg_weights(dtau, a, b, e);
    const double p_l = prev[ol], p_u = prev[ou];
    double I_fix, w_l, w_u;   // I_u = I_fix + w_l*car[jl] + w_u*car[ju]
    if (up) {   // I_vals = [I_0 row; carried row]
        I_fix = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, p_l, p_u, 0.0, 0.0);
        w_l = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, 0.0, 0.0, 1.0, 0.0);
        w_u = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, 0.0, 0.0, 0.0, 1.0);
    } else {    // I_vals = [carried row; I_0 row]
        I_fix = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, 0.0, 0.0, p_l, p_u);
        w_l = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, 1.0, 0.0, 0.0, 0.0);
        w_u = reg_bilinear(zup, jup, zb1, zb2, jb1, jb2, 0.0, 1.0, 0.0, 0.0);
    }
    cA[oc] = e * w_l;
    cB[oc] = e * w_u;
    cC[oc] = e * I_fix + a * S_u + b * S_c[oc];
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

// Device workspace of vrt_regular_formal_solve, kept between calls (a Λ-iteration calls it n_directions times per
// iteration with the same shapes; cudaMalloc/cudaFree of tens of GB cost more than the solve).  Grow-only; released by
// vrt_regular_release_workspace().
struct RegWorkspace {
    int device = -1;
    DevBuf<double> dS, dA, dI, stage, stage0, cA, cB, cC, dcj, dcs;
    size_t bytes() const { return 8 * (dS.n + dA.n + dI.n + stage.n + stage0.n + cA.n + cB.n + cC.n + dcj.n + dcs.n); }
    void release() {
        dS.release(); dA.release(); dI.release(); stage.release(); stage0.release();
        cA.release(); cB.release(); cC.release(); dcj.release(); dcs.release();
    }
};
std::mutex g_reg_mu;
RegWorkspace* g_reg_ws = nullptr;   // never destroyed at exit: the CUDA runtime may already be gone by then

inline int host_copy(double* dst, const double* src, int64_t n) {
    VRT_CUDA(cudaMemcpy(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDefault));
    return VRT_OK;
}

}  // namespace
}  // namespace vrt

using namespace vrt;

extern "C" int vrt_regular_formal_solve(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                                        const double k[3], int32_t down, int32_t n_sweeps, int64_t nlam, const double* S,
                                        const double* alpha, const double* I0, double* I_out, int32_t* plane_branch) {
    if (!z || !x || !y || !k || !S || !alpha || !I0 || !I_out || nlam <= 0 || nz < 2 || nx < 3 || ny < 3 || n_sweeps < 1) {
        set_error("vrt_regular_formal_solve: bad arguments");
        return VRT_E_INVALID;
    }
    if (nx > 1026 || ny > 1026) {
        set_error("vrt_regular_formal_solve: nx, ny <= 1026 (one thread per interior point of a row in the recurrence)");
        return VRT_E_INVALID;
    }
    if (!(k[0] == k[0]) || !(k[1] == k[1]) || !(k[2] == k[2]) || k[0] == 0.0) {
        set_error("vrt_regular_formal_solve: direction must be finite with k[0] != 0");
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("vrt_regular_formal_solve: no CUDA device (this library has no CPU path)");
        return VRT_E_CUDA;
    }
    std::vector<double> hz(nz), hx(nx), hy(ny);
    VRT_TRY(host_copy(hz.data(), z, nz));
    VRT_TRY(host_copy(hx.data(), x, nx));
    VRT_TRY(host_copy(hy.data(), y, ny));
    // characteristics.jl:34-40: path lengths to the x and y faces, xy_intersect signs
    const double dx = hx[1] - hx[0], dy = hy[1] - hy[0];
    const double r_x = fabs(dx / k[1]), r_y = fabs(dy / k[2]);
    int sx = 1, sy = 1;
    if (k[1] > 0 && k[2] > 0) { sx = -1; sy = -1; }
    else if (k[1] < 0 && k[2] > 0) { sx = 1; sy = -1; }
    else if (k[1] < 0 && k[2] < 0) { sx = 1; sy = 1; }
    else if (k[1] > 0 && k[2] < 0) { sx = -1; sy = 1; }
    // argmin([r_z, r_x, r_y]) takes the first minimum: a sideways plane is xz only when r_y < r_x strictly
    RegPlane P;
    P.par_is_x = (r_y < r_x) ? 1 : 0;
    P.np = (int)(P.par_is_x ? nx : ny);
    P.ns = (int)(P.par_is_x ? ny : nx);
    P.sgn_j = P.par_is_x ? sx : sy;
    P.sgn_s = P.par_is_x ? sy : sx;
    P.kz = k[0];
    P.kj = P.par_is_x ? k[1] : k[2];
    P.ks = P.par_is_x ? k[2] : k[1];
    const double r_side = P.par_is_x ? r_y : r_x;

    const size_t plane = (size_t)nx * ny;
    const size_t vol = plane * nz;
    const bool dev_S = is_device_ptr(S), dev_a = is_device_ptr(alpha), dev_I0 = is_device_ptr(I0), dev_out = is_device_ptr(I_out);
    const bool need_stage = !(dev_S && dev_a && dev_out);
    std::lock_guard<std::mutex> lock(g_reg_mu);
    if (!g_reg_ws) g_reg_ws = new RegWorkspace();
    RegWorkspace& W = *g_reg_ws;
    int cur_dev = 0;
    VRT_CUDA(cudaGetDevice(&cur_dev));
    if (W.device != cur_dev) { W.release(); W.device = cur_dev; }
    size_t free_b = 0, total_b = 0;
    VRT_CUDA(cudaMemGetInfo(&free_b, &total_b));
    free_b += W.bytes();   // what the workspace already holds is ours to reuse
    const double per_lam = 8.0 * ((3.0 + (need_stage ? 1.0 : 0.0)) * vol + 4.0 * plane);
    int64_t lc = (int64_t)std::min<double>((double)nlam, floor(0.9 * (double)free_b / per_lam));
    if (const char* e = getenv("VRT_REG_LAM_CHUNK")) lc = std::max<int64_t>(1, std::min<int64_t>(lc, atoll(e)));
    if (lc < 1) {
        set_error("vrt_regular_formal_solve: one wavelength of a %lld x %lld x %lld grid needs %.1f GB, %.1f GB free",
                  (long long)nz, (long long)nx, (long long)ny, per_lam / 1e9, free_b / 1e9);
        return VRT_E_NOMEM;
    }
    lc = std::min<int64_t>(lc, 65535);
    DevBuf<double>&dS = W.dS, &dA = W.dA, &dI = W.dI, &stage = W.stage, &stage0 = W.stage0, &cA = W.cA, &cB = W.cB, &cC = W.cC,
                  &dcj = W.dcj, &dcs = W.dcs;
    {
        const size_t need_vol = vol * lc, need_pl = plane * lc;
        // a grow of one buffer must not fail because the others hold stale, larger-than-needed space
        const bool grow = dS.n < need_vol || dA.n < need_vol || dI.n < need_vol || (need_stage && stage.n < need_vol) ||
                          (!dev_I0 && stage0.n < need_pl) || cA.n < need_pl || cB.n < need_pl || cC.n < need_pl;
        if (grow) W.release();
        VRT_TRY(dS.ensure(need_vol)); VRT_TRY(dA.ensure(need_vol)); VRT_TRY(dI.ensure(need_vol));
        if (need_stage) VRT_TRY(stage.ensure(need_vol));
        if (!dev_I0) VRT_TRY(stage0.ensure(need_pl));
        VRT_TRY(cA.ensure(need_pl)); VRT_TRY(cB.ensure(need_pl)); VRT_TRY(cC.ensure(need_pl));
        VRT_TRY(dcj.ensure(P.np)); VRT_TRY(dcs.ensure(P.ns));
    }
    VRT_CUDA(cudaMemcpy(dcj.p, P.par_is_x ? hx.data() : hy.data(), sizeof(double) * P.np, cudaMemcpyHostToDevice));
    VRT_CUDA(cudaMemcpy(dcs.p, P.par_is_x ? hy.data() : hx.data(), sizeof(double) * P.ns, cudaMemcpyHostToDevice));

    SweepStats stats;
    std::vector<int32_t> branch(nz, 0);
    cudaEvent_t ev0, ev1;
    VRT_CUDA(cudaEventCreate(&ev0));
    VRT_CUDA(cudaEventCreate(&ev1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};

    const dim3 tb(32, 8);
    // user array (host or device) -> internal layout of the current chunk
    auto load = [&](const double* src, bool on_dev, int64_t nz_, double* dst, DevBuf<double>& stg, int64_t l0, int64_t n_l) -> int {
        const double* s_dev = src;
        int64_t ld = nlam, lo = l0;
        const size_t rows = (size_t)nz_ * plane;
        if (!on_dev) {
            if (n_l == nlam) VRT_CUDA(cudaMemcpy(stg.p, src, sizeof(double) * rows * nlam, cudaMemcpyHostToDevice));
            else VRT_CUDA(cudaMemcpy2D(stg.p, sizeof(double) * n_l, src + l0, sizeof(double) * nlam, sizeof(double) * n_l, rows, cudaMemcpyHostToDevice));
            s_dev = stg.p; ld = n_l; lo = 0;
        }
        const dim3 grid((unsigned)((n_l * nz_ + 31) / 32), (unsigned)((P.np + 31) / 32), (unsigned)P.ns);
        k_reg_transpose<1><<<grid, tb>>>(s_dev, dst, ld, lo, (int)n_l, nz_, nx, ny, P.par_is_x);
        VRT_CUDA(cudaGetLastError());
        stats.kernels += 1;
        return VRT_OK;
    };

    for (int64_t l0 = 0; l0 < nlam; l0 += lc) {
        const int64_t n_l = std::min<int64_t>(lc, nlam - l0);
        P.lc = (int)n_l;
        const size_t pst = plane * n_l;   // plane stride of the internal arrays
        VRT_TRY(load(S, dev_S, nz, dS.p, stage, l0, n_l));
        VRT_CUDA(cudaDeviceSynchronize());   // `stage` is reused
        VRT_TRY(load(alpha, dev_a, nz, dA.p, stage, l0, n_l));
        const int64_t zb = down ? nz - 1 : 0;
        // I[zb, :, :] = I_0 (characteristics.jl:45 / :133), ghost columns as the caller filled them
        VRT_TRY(load(I0, dev_I0, 1, dI.p + pst * zb, stage0, l0, n_l));
        VRT_CUDA(cudaEventRecord(ev0));
        for (int64_t step = 1; step < nz; step++) {
            const int64_t idz = down ? nz - 1 - step : step;        // 0-based plane being solved
            const int64_t idu = down ? idz + 1 : idz - 1;           // upwind plane
            const double dz = down ? hz[idz + 1] - hz[idz] : hz[idz] - hz[idz - 1];
            const double r_z = fabs(dz / k[0]);
            int cut = 1;
            double best = r_z;
            if (r_x < best) { best = r_x; cut = 2; }
            if (r_y < best) { best = r_y; cut = 3; }
            branch[idz] = cut;
            double* Ic = dI.p + pst * idz;
            const double* Iu = dI.p + pst * idu;
            if (cut == 1) {
                const dim3 grid((unsigned)((P.np * P.ns + 255) / 256), (unsigned)n_l);
                k_reg_xy<<<grid, 256>>>(P, hz[idu] - hz[idz], dcj.p, dcs.p, dS.p + pst * idz, dA.p + pst * idz, dS.p + pst * idu,
                                        dA.p + pst * idu, Iu, Ic);
                stats.kernels += 1;
            } else {
                const int up = down ? 0 : 1;
                const int64_t izl = up ? idz - 1 : idz, izu = izl + 1;
                const int64_t izc = P.par_is_x ? izu : idz;          // Q13: xz takes the centre from the upper plane
                const dim3 grid((unsigned)((P.np * P.ns + 255) / 256), (unsigned)n_l);
                k_reg_coef<<<grid, 256>>>(P, up, hz[idz], hz[izl], hz[izu], dcj.p, dcs.p, dS.p + pst * izl, dS.p + pst * izu,
                                          dA.p + pst * izl, dA.p + pst * izu, dS.p + pst * izc, dA.p + pst * izc, Iu, cA.p, cB.p, cC.p);
                const int threads = ((P.np - 2 + 31) / 32) * 32;
                const size_t rec_smem = sizeof(double) * 2 * P.np;
                if (threads <= 512) k_reg_rec<512, REG_PF><<<(unsigned)n_l, threads, rec_smem>>>(P, n_sweeps, cA.p, cB.p, cC.p, Ic);
                else k_reg_rec<1024, REG_PF / 3><<<(unsigned)n_l, threads, rec_smem>>>(P, n_sweeps, cA.p, cB.p, cC.p, Ic);
                stats.kernels += 2;
                stats.steps += (double)n_sweeps * (P.ns - 2);
            }
            VRT_CUDA(cudaGetLastError());
        }
        VRT_CUDA(cudaEventRecord(ev1));
        // internal -> caller's layout
        {
            double* d_dst = dev_out ? I_out : stage.p;
            const int64_t ld = dev_out ? nlam : n_l, lo = dev_out ? l0 : 0;
            const dim3 grid((unsigned)((n_l * nz + 31) / 32), (unsigned)((P.np + 31) / 32), (unsigned)P.ns);
            k_reg_transpose<0><<<grid, tb>>>(dI.p, d_dst, ld, lo, (int)n_l, nz, nx, ny, P.par_is_x);
            VRT_CUDA(cudaGetLastError());
            stats.kernels += 1;
            if (!dev_out) {
                if (n_l == nlam) VRT_CUDA(cudaMemcpy(I_out, stage.p, sizeof(double) * vol * nlam, cudaMemcpyDeviceToHost));
                else VRT_CUDA(cudaMemcpy2D(I_out + l0, sizeof(double) * nlam, stage.p, sizeof(double) * n_l, sizeof(double) * n_l, vol, cudaMemcpyDeviceToHost));
            }
        }
        VRT_CUDA(cudaDeviceSynchronize());
        float ms = 0;
        VRT_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        stats.sweep_ms += ms;
    }
    (void)r_side;
    stats.visits = (double)(nz - 1) * (double)(nx - 2) * (double)(ny - 2) * (double)nlam;
    g_last_stats = stats;
    if (plane_branch) VRT_CUDA(cudaMemcpy(plane_branch, branch.data(), sizeof(int32_t) * nz, cudaMemcpyDefault));
    return VRT_OK;
}

extern "C" int vrt_regular_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_reg_mu);
    if (g_reg_ws) g_reg_ws->release();
    return VRT_OK;
}
