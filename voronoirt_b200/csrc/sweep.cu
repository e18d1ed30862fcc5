// sweep.cu — K5: the layered short-characteristics sweep as ONE persistent cooperative kernel.
//
// Replaces the hot loop of Delaunay_upII / Delaunay_downII (reference src/irregular_ray_tracing.jl:37-80,
// :118-161).  The kernel interprets the sweep programs built by schedule.cu for up to MAX_DIRS
// directions at once: a global step t runs the t-th dependent step of every direction, then all CTAs
// meet at a grid barrier.  Inside a step the work is flat over (visit, wavelength) with the wavelength
// innermost, so every gather of S, alpha and I of an upwind cell is a contiguous fp64 row read.
// Per item (irregular_ray_tracing.jl:54-77, functions.jl:392-395, :484-500):
//     Δτ_m = r_m (α_c + α_um)/2 ;  (a,b,e) = linear_weights(Δτ_m)
//     I_c  = 0 + w_1 (e_1 I_u1 + a_1 S_u1 + b_1 S_c) + w_2 (e_2 I_u2 + a_2 S_u2 + b_2 S_c)
// Roofline: HBM-bound; algorithmic bytes per (cell,direction,wavelength) update are given in DESIGN.md.
#include <algorithm>
#include "vrt_internal.h"

namespace vrt {

struct DirDev {
    const Visit* visits;
    const double* alpha;
    double* I_main;
    double* scratch[MAX_SWEEPS];
};

struct SweepParams {
    const DirDev* dirs;      // [nd] in global memory
    const int32_t* goffT;    // [(T+1)][nd]: visit offset of direction d at global step t
    const double* S;         // [n][ldS], first wavelength of the chunk
    int64_t ldS;             // row stride of S
    unsigned int* barrier;   // grid barrier counter (zeroed before launch)
    int nd;
    int T;
    int nlam;
    int cpw;                 // visits per warp chunk
};

// linear_weights (functions.jl:484-500)
__device__ __forceinline__ void linear_weights(double dtau, double& a, double& b, double& e) {
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

__device__ __forceinline__ double load_I(uint32_t src, const DirDev* __restrict__ D, int nlam, int l) {
    uint32_t sel = src >> SEL_SHIFT;
    if (sel == SEL_ZERO) return 0.0;
    const double* base = sel == SEL_MAIN ? D->I_main : D->scratch[sel - SEL_SCR0];
    // I is written by other SMs earlier in this launch: read through L2, never a stale L1 line
    return __ldcg(base + (size_t)(src & ROW_MASK) * nlam + l);
}

struct VisitRegs {
    uint4 a, b;
    double2 w, hr;
};

__device__ __forceinline__ VisitRegs load_visit(const Visit* v) {
    const uint4* q = reinterpret_cast<const uint4*>(v);
    VisitRegs r;
    r.a = __ldg(q);
    r.b = __ldg(q + 1);
    r.w = __ldg(reinterpret_cast<const double2*>(q + 2));
    r.hr = __ldg(reinterpret_cast<const double2*>(q + 3));
    return r;
}

__device__ __forceinline__ void do_item(const VisitRegs& v, const DirDev* __restrict__ D, const double* __restrict__ S,
                                        int64_t ldS, int nlam, int l) {
    const uint32_t cell = v.a.x, dst = v.a.y, u1 = v.a.z, u2 = v.a.w, src1 = v.b.x, src2 = v.b.y;
    const double* __restrict__ alpha = D->alpha;
    const size_t oc = (size_t)cell * nlam + l, o1 = (size_t)u1 * nlam + l, o2 = (size_t)u2 * nlam + l;
    // issue all gathers before the math
    double a_c = __ldg(alpha + oc), S_c = __ldg(S + (size_t)cell * ldS + l);
    double a_1 = __ldg(alpha + o1), S_1 = __ldg(S + (size_t)u1 * ldS + l);
    double a_2 = __ldg(alpha + o2), S_2 = __ldg(S + (size_t)u2 * ldS + l);
    double I_1 = load_I(src1, D, nlam, l);
    double I_2 = load_I(src2, D, nlam, l);
    double a, b, e;
    linear_weights(v.hr.x * (a_c + a_1), a, b, e);
    double I = 0.0 + (e * I_1 + a * S_1 + b * S_c) * v.w.x;
    linear_weights(v.hr.y * (a_c + a_2), a, b, e);
    I += (e * I_2 + a * S_2 + b * S_c) * v.w.y;
    uint32_t sel = dst >> SEL_SHIFT;
    double* base = sel == SEL_MAIN ? D->I_main : D->scratch[sel - SEL_SCR0];
    base[(size_t)(dst & ROW_MASK) * nlam + l] = I;
}

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned int*)counter < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

constexpr int SWEEP_BLOCK = 256;

template <bool SINGLE>
__global__ void __launch_bounds__(SWEEP_BLOCK) k_sweep(const SweepParams P) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp_global = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
    const int nlam = P.nlam;
    const double* __restrict__ S = P.S;
    const int64_t ldS = P.ldS;
    for (int t = 0; t < P.T; t++) {
        int beg = 0, end = 0, cnt = 0;
        if (lane < P.nd) {
            beg = __ldg(P.goffT + (size_t)t * P.nd + lane);
            end = __ldg(P.goffT + (size_t)(t + 1) * P.nd + lane);
            cnt = (end - beg + P.cpw - 1) / P.cpw;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(full, incl, 31);
        const int excl = incl - cnt;
        for (int ch = warp_global; ch < total; ch += nwarps) {
            unsigned m = __ballot_sync(full, ch >= excl && ch < incl);
            int d = __ffs(m) - 1;
            int dbeg = __shfl_sync(full, beg, d);
            int dend = __shfl_sync(full, end, d);
            int dexcl = __shfl_sync(full, excl, d);
            const DirDev* __restrict__ D = P.dirs + d;
            const Visit* __restrict__ visits = D->visits;
            int vb = dbeg + (ch - dexcl) * P.cpw;
            if (SINGLE) {
                VisitRegs v = load_visit(visits + vb);
                for (int l = lane; l < nlam; l += 32) do_item(v, D, S, ldS, nlam, l);
            } else {
                int ve = min(vb + P.cpw, dend);
                int nitems = (ve - vb) * nlam;
                for (int i = lane; i < nitems; i += 32) {
                    int vi = i / nlam;
                    int l = i - vi * nlam;
                    VisitRegs v = load_visit(visits + vb + vi);
                    do_item(v, D, S, ldS, nlam, l);
                }
            }
        }
        grid_barrier(P.barrier, (unsigned)(t + 1) * gridDim.x);
    }
}

int sweep_run(vrt_grid* g, int nd, const SweepDir* dirs, const double* S, int64_t ldS, int64_t nlam, cudaStream_t st, SweepStats* stats) {
    if (nd <= 0) return VRT_OK;
    if (nd > MAX_DIRS) {
        set_error("sweep_run: %d directions exceed MAX_DIRS=%d", nd, MAX_DIRS);
        return VRT_E_INVALID;
    }
    // ---- merge the per-direction programs: align steps by (layer, sweep) within the up group and within
    // the down group (keeps all directions on the same layers => S rows are shared in L2), zip up with down
    int T = 0;
    std::vector<std::vector<int32_t>> goff(nd);
    for (int grp = 0; grp < 2; grp++) {
        size_t nls = 0;
        for (int d = 0; d < nd; d++)
            if (dirs[d].sch->down == grp) nls = std::max(nls, dirs[d].sch->nsub.size());
        if (nls == 0) continue;
        std::vector<int32_t> mx(nls, 0);
        int any = 0;
        for (int d = 0; d < nd; d++) {
            const DirSchedule* s = dirs[d].sch;
            if (s->down != grp) continue;
            any = 1;
            if (s->nsub.size() != nls || s->n_sweeps != dirs[0].sch->n_sweeps) {
                set_error("sweep_run: schedules of one launch must share the grid and n_sweeps");
                return VRT_E_INVALID;
            }
            for (size_t i = 0; i < nls; i++) mx[i] = std::max(mx[i], s->nsub[i]);
        }
        if (!any) continue;
        int64_t Tg = 0;
        for (size_t i = 0; i < nls; i++) Tg += mx[i];
        for (int d = 0; d < nd; d++) {
            const DirSchedule* s = dirs[d].sch;
            if (s->down != grp) continue;
            std::vector<int32_t>& o = goff[d];
            o.reserve((size_t)Tg + 1);
            int64_t local = 0;
            for (size_t i = 0; i < nls; i++) {
                for (int32_t j = 0; j < mx[i]; j++) {
                    o.push_back((int32_t)s->step_off[(size_t)local]);
                    if (j < s->nsub[i]) local++;
                }
            }
            o.push_back((int32_t)s->step_off[(size_t)local]);
        }
        T = std::max<int64_t>(T, Tg);
    }
    std::vector<int32_t> goffT((size_t)(T + 1) * nd);
    double visits = 0;
    for (int d = 0; d < nd; d++) {
        std::vector<int32_t>& o = goff[d];
        int32_t last = o.empty() ? 0 : o.back();
        for (int t = 0; t <= T; t++) goffT[(size_t)t * nd + d] = t < (int)o.size() ? o[t] : last;
        visits += (double)dirs[d].sch->n_visits;
    }
    if (T == 0) return VRT_OK;

    std::vector<DirDev> hd(nd);
    for (int d = 0; d < nd; d++) {
        hd[d].visits = dirs[d].sch->visits.p;
        hd[d].alpha = dirs[d].alpha;
        hd[d].I_main = dirs[d].I_main;
        for (int s = 0; s < MAX_SWEEPS; s++) hd[d].scratch[s] = dirs[d].scratch[s];
    }
    DevBuf<DirDev> d_dirs;
    DevBuf<int32_t> d_goffT;
    DevBuf<unsigned int> d_bar;
    VRT_TRY(d_dirs.alloc(nd)); VRT_TRY(d_goffT.alloc(goffT.size())); VRT_TRY(d_bar.alloc(1));
    VRT_CUDA(cudaMemcpyAsync(d_dirs.p, hd.data(), sizeof(DirDev) * nd, cudaMemcpyHostToDevice, st));
    VRT_CUDA(cudaMemcpyAsync(d_goffT.p, goffT.data(), sizeof(int32_t) * goffT.size(), cudaMemcpyHostToDevice, st));
    VRT_CUDA(cudaMemsetAsync(d_bar.p, 0, sizeof(unsigned int), st));

    SweepParams P;
    P.dirs = d_dirs.p;
    P.goffT = d_goffT.p;
    P.S = S;
    P.ldS = ldS;
    P.barrier = d_bar.p;
    P.nd = nd;
    P.T = T;
    P.nlam = (int)nlam;
    P.cpw = nlam >= 48 ? 1 : (int)((64 + nlam - 1) / nlam);

    int dev = 0, sms = 0, per_sm = 0;
    VRT_CUDA(cudaGetDevice(&dev));
    VRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const void* fn = P.cpw == 1 ? (const void*)k_sweep<true> : (const void*)k_sweep<false>;
    VRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, SWEEP_BLOCK, 0));
    if (per_sm < 1) {
        set_error("sweep kernel cannot be resident");
        return VRT_E_CUDA;
    }
    int grid = sms * per_sm;
    if ((double)T * grid >= 4.0e9) {
        set_error("sweep program too long for the 32-bit barrier counter");
        return VRT_E_INVALID;
    }
    cudaEvent_t e0, e1;
    VRT_CUDA(cudaEventCreate(&e0));
    VRT_CUDA(cudaEventCreate(&e1));
    VRT_CUDA(cudaEventRecord(e0, st));
    void* args[] = {(void*)&P};
    cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(SWEEP_BLOCK), args, 0, st);
    if (le != cudaSuccess) {
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return cuda_fail(le, "cudaLaunchCooperativeKernel(k_sweep)", __FILE__, __LINE__);
    }
    VRT_CUDA(cudaEventRecord(e1, st));
    VRT_CUDA(cudaEventSynchronize(e1));  // also keeps the program tables alive until the kernel is done
    float ms = 0;
    VRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    VRT_CUDA(cudaGetLastError());
    if (stats) {
        stats->kernels += 1;
        stats->visits += visits;
        stats->steps += T;
        stats->sweep_ms += ms;
    }
    return VRT_OK;
}

}  // namespace vrt
