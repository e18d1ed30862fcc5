// sweep.cu — K5: the layered short-characteristics sweep as ONE persistent cooperative kernel.
//
// Replaces the hot loop of Delaunay_upII / Delaunay_downII (reference src/irregular_ray_tracing.jl:37-80,
// :118-161).  The kernel interprets the sweep programs built by schedule.cu for up to MAX_DIRS
// directions at once.  There is NO grid barrier: the merged program is one long topologically ordered list
// of chunks (cv visits each); warp w owns chunks w, w+W, ... and a chunk only waits, by spinning on per-chunk
// ready-flags, for the chunks that produce its two upwind intensities (sync-free SpTRSV style).  Inside a
// chunk the work is flat over (visit, wavelength) with the wavelength innermost, so every gather of S, alpha
// and I of an upwind cell is a contiguous fp64 row read.
// Per item (irregular_ray_tracing.jl:54-77, functions.jl:392-395, :484-500):
//     Δτ_m = r_m (α_c + α_um)/2 ;  (a,b,e) = linear_weights(Δτ_m)
//     I_c  = 0 + w_1 (e_1 I_u1 + a_1 S_u1 + b_1 S_c) + w_2 (e_2 I_u2 + a_2 S_u2 + b_2 S_c)
// Roofline: HBM-bound; algorithmic bytes per (cell,direction,wavelength) update are given in DESIGN.md.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include "vrt_internal.h"

namespace vrt {

struct DirDev {
    const Visit* visits;
    const double* alpha;
    double* I_main;
    double* scratch[MAX_SWEEPS];
    int32_t* flags;          // one ready-flag per chunk of this direction's program (== epoch when done)
    const int32_t* step_off; // [nsteps+1] visit-slot offsets of this direction's dependent steps (multiples of cv)
    int32_t nsteps;
    int32_t pad;
};

// the programs of the directions in flight, passed by value (kernel parameter space: no allocation, no copy per launch)
struct DirTable {
    DirDev d[MAX_DIRS];
};

struct SweepParams {
    const double* S;         // [n][ldS], first wavelength of the chunk
    int64_t ldS;             // row stride of S
    int nd;
    int T;                   // max over directions of nsteps; global step g = t*nd + d runs step t of direction d
    int nlam;
    int cv;                  // visits per chunk
    int32_t epoch;
    int32_t run_len;         // visits per producer run (<= 32)
    int32_t pacing;          // 1: global round r runs step r*T_d/T of direction d (same relative progress); 0: step r
    int32_t experiment;      // VRT_EXPERIMENT: timing experiments only (0 in production)
    unsigned long long* prof; // experiment 2: per-role cycle counters
};

// Constants of the sweep's exp, kept in the constant bank so that every DFMA takes its coefficient as a c[bank][offset]
// operand (as immediates they cost two extra issue slots each: the polynomial was 26 of 74 instructions UMOV/MOV).
__constant__ double EXP_C[18] = {
    1.4426950408889634074,        // 0  log2(e)
    6755399441055744.0,           // 1  1.5 * 2^52: rounds to nearest integer
    -6.93147180369123816490e-01,  // 2  -ln2 high part (fdlibm split)
    -1.90821492927058770002e-10,  // 3  -ln2 low part
    1.6059043836821614599e-10,    // 4  1/13!
    2.0876756987868098979e-09,    // 5  1/12!
    2.5052108385441718775e-08,    // 6  1/11!
    2.7557319223985890653e-07,    // 7  1/10!
    2.7557319223985892511e-06,    // 8  1/9!
    2.4801587301587301566e-05,    // 9  1/8!
    1.9841269841269841253e-04,    // 10 1/7!
    1.3888888888888889419e-03,    // 11 1/6!
    8.3333333333333332177e-03,    // 12 1/5!
    4.1666666666666664354e-02,    // 13 1/4!
    1.6666666666666665741e-01,    // 14 1/3!
    -708.0,                       // 15 lower clamp of the argument (exp(-708) = 3e-308 is still a normal number)
    3.3333333333333331483e-01,    // 16 1/3
    1.6666666666666665741e-01,    // 17 1/6
};

// exp(x) for x <= 0: Cody–Waite reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation
// 4e-18), scaling by 2^k through the exponent field.  No fp64<->int conversion, no special-case branches;
// relative error <= ~2 ulp for -708 <= x <= 0 (below, the argument is clamped: linear_weights never uses e for Δτ > 50).
__device__ __forceinline__ double exp_sweep(double x) {
    x = x < EXP_C[15] ? EXP_C[15] : x;   // (not fmax: its NaN handling costs five more instructions)
    const double t = fma(x, EXP_C[0], EXP_C[1]);
    const double kf = t - EXP_C[1];
    const int k = __double2loint(t);
    double r = fma(kf, EXP_C[2], x);
    r = fma(kf, EXP_C[3], r);
    double p = EXP_C[4];
    p = fma(p, r, EXP_C[5]);
    p = fma(p, r, EXP_C[6]);
    p = fma(p, r, EXP_C[7]);
    p = fma(p, r, EXP_C[8]);
    p = fma(p, r, EXP_C[9]);
    p = fma(p, r, EXP_C[10]);
    p = fma(p, r, EXP_C[11]);
    p = fma(p, r, EXP_C[12]);
    p = fma(p, r, EXP_C[13]);
    p = fma(p, r, EXP_C[14]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// 1/a for a > 0: hardware seed (MUFU.RCP64H) + two Newton steps -> full double precision (<= 1 ulp)
__device__ __forceinline__ double rcp_sweep(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// linear_weights (functions.jl:484-500), branch-free: a warp's lanes hold different wavelengths (line core: large Δτ, far
// wing and corona: Δτ < 5e-4), so branches would serialise.  The three cases are folded as follows:
//   Δτ > 50:  the reference sets e = 0, a = 1/Δτ, b = 1 - a.  The general expressions give exactly these a and b there
//             (exp(-Δτ) < 2e-22 is below half an ulp of 1, of 1/Δτ and of 1 - 1/Δτ), so only e needs a select;
//   Δτ < 5e-4: Taylor forms (the two divisions by 3 and 6 as multiplications: x/3 and x*(1/3) differ by at most an ulp of a
//             term that is < 2e-4 of the 0.5 it is subtracted from).
__device__ __forceinline__ void linear_weights(double dtau, double& a, double& b, double& e) {
    const double ex = exp_sweep(-dtau);
    const double inv = rcp_sweep(dtau);
    const double a_mid = (1 - ex) * inv - ex;
    const double b_mid = 1 - a_mid - ex;
    const bool small = dtau < 5e-4;
    e = small ? (1 - dtau + 0.5 * (dtau * dtau)) : (dtau > 50 ? 0.0 : ex);
    a = small ? dtau * (0.5 - dtau * EXP_C[16]) : a_mid;
    b = small ? dtau * (0.5 - dtau * EXP_C[17]) : b_mid;
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

// dataflow synchronisation (no grid barrier): a chunk spins on the ready-flags of the chunks that produce its
// upwind intensities.  Chunks are claimed in a fixed global topological order (warp w owns chunks w, w+W, ...),
// every warp of the cooperative launch is resident, so by induction on the chunk index no wait can deadlock.
// Polling uses relaxed loads (an acquire load costs an L1 invalidate per poll); `acquire` adds the one acquire that
// orders the subsequent generic-proxy reads of the intensities.  The TMA producer polls with wait_flags2, which ends with
// an acquire fence, and then crosses to the async proxy with fence.proxy.async before its bulk copies.
__device__ __forceinline__ void wait_flag(const int32_t* f, int32_t epoch, bool acquire) {
    int v;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v == epoch) break;
        __nanosleep(100);
    }
    if (acquire) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
}
// both producer flags of a visit, polled together (one L2 round trip when they are already set)
__device__ __forceinline__ void wait_flags2(const int32_t* flags, uint32_t d1, uint32_t d2, int32_t epoch, bool acquire) {
    const int32_t* f1 = flags + (d1 == DEP_NONE ? 0 : d1);
    const int32_t* f2 = flags + (d2 == DEP_NONE ? 0 : d2);
    for (;;) {
        int v1, v2;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v1) : "l"(f1) : "memory");
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v2) : "l"(f2) : "memory");
        if ((d1 == DEP_NONE || v1 == epoch) && (d2 == DEP_NONE || v2 == epoch)) break;
        __nanosleep(100);
    }
    // the relaxed polls alone order nothing: this fence makes them an acquire (it pairs with the consumers'
    // fence.acq_rel + relaxed flag store), so the producers' result rows are visible before anything issued after it
    if (acquire && (d1 != DEP_NONE || d2 != DEP_NONE)) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// lane 0 publishes the chunk after __syncwarp(): the release is cumulative over the other lanes' stores, which are
// ordered before it by the warp barrier (PTX memory model: causality order through bar.warp.sync)
__device__ __forceinline__ void set_flag(int32_t* f, int32_t epoch) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
// several chunks at once: ONE release fence (it waits for the outstanding result stores), then relaxed flag stores.
// Back-to-back st.release would each wait a full store round trip.
__device__ __forceinline__ void set_flags(int32_t* const* f, int n, int32_t epoch) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    for (int i = 0; i < n; i++) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(f[i]), "r"(epoch) : "memory");
}

constexpr int SWEEP_BLOCK = 256;
#ifndef SWEEP_MIN_BLOCKS
#define SWEEP_MIN_BLOCKS 3
#endif

template <bool SINGLE>
__global__ void __launch_bounds__(SWEEP_BLOCK, SWEEP_MIN_BLOCKS) k_sweep(const SweepParams P, const __grid_constant__ DirTable DT) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)((gridDim.x * (unsigned)blockDim.x) >> 5);
    long long gnext = (long long)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);  // next global chunk of this warp
    long long base = 0;                                                                     // global index of step t's first chunk
    const int nlam = P.nlam, cv = P.cv;
    const int32_t epoch = P.epoch;
    const double* __restrict__ S = P.S;
    const int64_t ldS = P.ldS;
    // global step g = t*nd + d: step t of direction d.  Consecutive global steps belong to different directions, so
    // the dependent steps t and t+1 of one direction are separated by the independent work of all the others and the
    // producers' flags are (almost) always set by the time a chunk looks at them.
    const int G = P.T * P.nd;
    for (int g = 0; g < G; g++) {
        const int d = g % P.nd, r = g / P.nd;
        const DirDev* __restrict__ D = DT.d + d;
        int t = r;
        if (P.pacing) {   // proportional pacing: every direction advances through its program at the same relative speed
            const long long Td = D->nsteps;
            t = (int)(((long long)r * Td) / P.T);
            if ((int)(((long long)(r + 1) * Td) / P.T) == t) continue;
        } else if (t >= D->nsteps)
            continue;
        const int beg = __ldg(D->step_off + t);
        const int total = (__ldg(D->step_off + t + 1) - beg) / cv;
        const Visit* __restrict__ visits = D->visits;
        int32_t* flags = D->flags;
        while (gnext < base + total) {
            const int cj = beg / cv + (int)(gnext - base);   // chunk index inside direction d's program
            if (SINGLE) {
                VisitRegs v = load_visit(visits + cj);
                if (lane < 2) {
                    uint32_t dep = lane ? v.b.w : v.b.z;
                    if (dep != DEP_NONE) wait_flag(flags + dep, epoch, true);
                }
                __syncwarp();
                for (int l = lane; l < nlam; l += 32) do_item(v, D, S, ldS, nlam, l);
            } else {
                const int vb = cj * cv;
                for (int q = lane; q < 2 * cv; q += 32) {
                    uint32_t dep = __ldg(reinterpret_cast<const uint32_t*>(visits + vb + (q >> 1)) + 6 + (q & 1));
                    if (dep != DEP_NONE) wait_flag(flags + dep, epoch, true);
                }
                __syncwarp();
                const int nitems = cv * nlam;
                for (int i = lane; i < nitems; i += 32) {
                    int vi = i / nlam;
                    int l = i - vi * nlam;
                    VisitRegs v = load_visit(visits + vb + vi);
                    if (v.a.x != CELL_DUMMY) do_item(v, D, S, ldS, nlam, l);
                }
            }
            __syncwarp();
            if (lane == 0) set_flag(flags + cj, epoch);
            gnext += nwarps;
        }
        base += total;
    }
}

// ---------------------------------------------------------------- TMA-pipelined variant (wide rows)
// One producer warp per CTA walks the CTA's chunks (CTA b owns global chunks b, b+B, ...), waits for the two
// producer flags of a visit and then issues EIGHT 1-D bulk-tensor copies (cp.async.bulk, completion on an
// mbarrier) that land the visit's rows  α_c S_c | α_u1 S_u1 I_u1 | α_u2 S_u2 I_u2  in a shared-memory stage.
// NC consumer warps drain the stages: wavelengths across lanes, math in registers, result row written
// straight to HBM, flag released, stage handed back.  The stage ring keeps tens of KB of gathers in flight
// per SM independent of register pressure, which is what a latency-bound irregular gather needs.
struct __align__(16) StageHdr {
    double* dst;
    int32_t* flag;
    double w1, w2, hr1, hr2;
    uint32_t offs;    // bit r: row r starts one double into its 16-byte aligned copy; bit 8+r: row r is all zero
    uint32_t valid;
    uint32_t seq;     // index of the chunk in this CTA's sequence: disambiguates mbarrier phases two rounds apart
    uint32_t pad;
};
static_assert(sizeof(StageHdr) == 64, "StageHdr must be 64 bytes");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    // the suspend-time hint keeps a waiting warp parked in hardware instead of re-issuing the try_wait every few cycles
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(2000u) : "memory");
    } while (!done);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


constexpr int TMA_MAX_STAGES = 32;
#ifndef TMA_DEFER
#define TMA_DEFER 12
#endif

// TMA_NP producer warps (independent issue chains) + TMA_NC consumer warps per CTA
template <int TMA_NP, int TMA_NC>
__global__ void __launch_bounds__(32 * (TMA_NP + TMA_NC), 3) k_sweep_tma(const SweepParams P, const __grid_constant__ DirTable DT, int ns, int rowb) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    // stage hand-back: consumed[s] counts the rounds of stage s that have been drained.  A counter (not an mbarrier
    // parity) lets several rounds of one stage be pending, so a producer run can be a full warp of 32 visits.
    volatile uint32_t* consumed = reinterpret_cast<volatile uint32_t*>(full + TMA_MAX_STAGES);
    unsigned int* next_chunk = const_cast<unsigned int*>(reinterpret_cast<volatile unsigned int*>(consumed + TMA_MAX_STAGES));   // consumer dispatch counter
    StageHdr* hdr = reinterpret_cast<StageHdr*>(smem + 2 * TMA_MAX_STAGES * sizeof(uint64_t));
    DirDev* sdirs = reinterpret_cast<DirDev*>(smem + 2 * TMA_MAX_STAGES * sizeof(uint64_t) + TMA_MAX_STAGES * sizeof(StageHdr));
    unsigned char* data = smem + 2 * TMA_MAX_STAGES * sizeof(uint64_t) + TMA_MAX_STAGES * sizeof(StageHdr) + MAX_DIRS * sizeof(DirDev);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nlam = P.nlam;
    const int32_t epoch = P.epoch;
    for (int i = threadIdx.x; i < P.nd * (int)(sizeof(DirDev) / 8); i += blockDim.x)
        reinterpret_cast<unsigned long long*>(sdirs)[i] = reinterpret_cast<const unsigned long long*>(DT.d)[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; s++) {
            mbar_init(full + s, 1);
            consumed[s] = 0;
            if (s == 0) *next_chunk = 0;
            hdr[s].seq = 0xffffffffu;   // shared memory may still hold the headers of an earlier launch
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp < TMA_NP) {
        // ===== producers: every lane owns one visit of a run of up to 32 consecutive chunks; the CTA's runs are dealt
        // round-robin to its TMA_NP producer warps so that the latency chains (visit record -> flags -> stage ->
        // bulk copies) of consecutive runs overlap =====
        // (the 32 visit records of a run are one coalesced 2 KB read; flags, stage hand-back and the eight bulk
        //  copies of a visit are then handled by its lane, so up to 32 visits are being issued at once)
        long long rnext = blockIdx.x;     // next global run of this CTA (CTA b owns runs b, b+B, ...)
        long long rbase = 0;              // global index of step t's first run
        const long long nblk = gridDim.x;
        unsigned it = 0;                  // chunks handed to the consumers so far
        unsigned rc = 0;                  // runs of this CTA so far
        const int RL = P.run_len;
        const double* __restrict__ S = P.S;
        const int64_t ldS = P.ldS;
        const int G = P.T * P.nd;
        for (int g = 0; g < G; g++) {
            const int d = g % P.nd, r = g / P.nd;
            const DirDev* __restrict__ D = sdirs + d;
            int t = r;
            if (P.pacing) {   // proportional pacing: every direction advances through its program at the same relative speed
                const long long Td = D->nsteps;
                t = (int)(((long long)r * Td) / P.T);
                if ((int)(((long long)(r + 1) * Td) / P.T) == t) continue;
            } else if (t >= D->nsteps)
                continue;
            const int dbeg = __ldg(D->step_off + t);
            const int dlen = __ldg(D->step_off + t + 1) - dbeg;
            const int total = (dlen + RL - 1) / RL;           // runs of this step
            while (rnext < rbase + total) {
                const int first = (int)(rnext - rbase) * RL;
                const int nrun = min(RL, dlen - first);       // chunks in this run
                long long r0 = 0;
                if (P.experiment == 2) r0 = clock64();
                if (lane < nrun && (int)(rc % TMA_NP) == warp) {
                    const int cj = dbeg + first + lane;
                    const VisitRegs v = load_visit(D->visits + cj);
                    const unsigned my = it + (unsigned)lane;
                    const int stage = (int)(my % (unsigned)ns);
                    const uint32_t round = my / (unsigned)ns;
                    wait_flags2(D->flags, v.b.z, v.b.w, epoch, P.experiment != 4);   // ends with an acquire fence (generic proxy; experiment 4 times its cost)
                    asm volatile("fence.proxy.async;" ::: "memory");   // ... which the bulk copies (async proxy) are ordered after
                    // rows of the stage: 0 α_c, 1 S_c, 2 α_u1, 3 S_u1, 4 I_u1, 5 α_u2, 6 S_u2, 7 I_u2
                    const double* src[8];
                    const double* alpha = D->alpha;
                    src[0] = alpha + (size_t)v.a.x * nlam;
                    src[1] = S + (size_t)v.a.x * ldS;
                    src[2] = alpha + (size_t)v.a.z * nlam;
                    src[3] = S + (size_t)v.a.z * ldS;
                    src[5] = alpha + (size_t)v.a.w * nlam;
                    src[6] = S + (size_t)v.a.w * ldS;
                    {
                        const uint32_t s1 = v.b.x >> SEL_SHIFT, s2 = v.b.y >> SEL_SHIFT;
                        src[4] = s1 == SEL_ZERO ? nullptr : (s1 == SEL_MAIN ? D->I_main : D->scratch[s1 - SEL_SCR0]) + (size_t)(v.b.x & ROW_MASK) * nlam;
                        src[7] = s2 == SEL_ZERO ? nullptr : (s2 == SEL_MAIN ? D->I_main : D->scratch[s2 - SEL_SCR0]) + (size_t)(v.b.y & ROW_MASK) * nlam;
                    }
                    uint32_t offmask = 0, zeromask = 0, tot = 0;
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        if (src[r]) {
                            const uint32_t off = (uint32_t)((reinterpret_cast<uintptr_t>(src[r]) >> 3) & 1u);
                            offmask |= off << r;
                            tot += (uint32_t)(((nlam + off) * 8 + 15) & ~15);
                        } else
                            zeromask |= 1u << r;
                    }
                    StageHdr h;
                    const uint32_t dsel = v.a.y >> SEL_SHIFT;
                    h.dst = (dsel == SEL_MAIN ? D->I_main : D->scratch[dsel - SEL_SCR0]) + (size_t)(v.a.y & ROW_MASK) * nlam;
                    h.flag = D->flags + cj;
                    h.w1 = v.w.x; h.w2 = v.w.y; h.hr1 = v.hr.x; h.hr2 = v.hr.y;
                    h.offs = offmask | (zeromask << 8);
                    h.valid = 1;
                    h.seq = my;
                    h.pad = 0;
                    // take the stage as late as possible: its lifetime bounds the throughput of the ring
                    while (consumed[stage] != round) __nanosleep(100);
                    // the consumer's generic-proxy reads of this stage happen before its hand-back (fence + store below);
                    // order our async-proxy overwrite after having observed it
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    hdr[stage] = h;
                    mbar_arrive_expect_tx(full + stage, tot);
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        if (src[r]) {
                            const uint32_t off = (offmask >> r) & 1u;
                            const uint32_t nb = (uint32_t)(((nlam + off) * 8 + 15) & ~15);
                            if (P.experiment == 1) {   // timing experiment: same bytes in twice as many bulk copies
                                const uint32_t h1 = (nb / 2) & ~15u;
                                bulk_g2s(data + (size_t)(stage * 8 + r) * rowb, src[r] - off, h1, full + stage);
                                bulk_g2s(data + (size_t)(stage * 8 + r) * rowb + h1, reinterpret_cast<const unsigned char*>(src[r] - off) + h1, nb - h1, full + stage);
                            } else
                                bulk_g2s(data + (size_t)(stage * 8 + r) * rowb, src[r] - off, nb, full + stage);
                        }
                    }
                }
                __syncwarp();
                if (P.experiment == 2 && lane == 0) {
                    atomicAdd(P.prof + 3, (unsigned long long)(clock64() - r0));   // producer: whole run
                    atomicAdd(P.prof + 4, 1ull);
                }
                it += (unsigned)nrun;
                rc++;
                rnext += nblk;
            }
            rbase += total;
        }
        // one termination token per consumer warp (sent by producer 0)
        for (int k = 0; k < TMA_NC && warp == 0; k++) {
            const int stage = (int)(it % (unsigned)ns);
            if (lane == 0) {
                while (consumed[stage] != it / (unsigned)ns) __nanosleep(32);
                __threadfence_block();
                hdr[stage].valid = 0;
                hdr[stage].seq = it;
                mbar_arrive(full + stage);
            }
            __syncwarp();
            it++;
        }
    } else {
        // ===== consumers =====
        // The result row is stored straight to HBM; publishing its ready-flag needs a release fence that waits for
        // those stores, which under a saturated memory system takes microseconds.  The stage is therefore handed
        // back immediately and the flag publication is deferred: up to TMA_DEFER finished visits share one fence.
        // A consumer never blocks while it holds unpublished flags (the data it waits for may depend on them).
        int32_t* pend[TMA_DEFER];
        int np = 0;
        // chunks are dispatched dynamically: whichever consumer warp is free takes the next chunk of the CTA's sequence
        for (;;) {
            unsigned it = 0;
            if (lane == 0) it = atomicAdd(next_chunk, 1u);
            it = __shfl_sync(0xffffffffu, it, 0);
            const int stage = (int)(it % (unsigned)ns);
            const uint32_t par = (it / (unsigned)ns) & 1u;
            // producer lanes fill stages out of order, so a parity wait alone could be satisfied by the fill of two
            // rounds ago; the sequence number in the header tells the rounds apart
            long long k0 = 0, k1 = 0;
            if (P.experiment == 2) k0 = clock64();
            bool ready = mbar_test(full + stage, par) && reinterpret_cast<volatile StageHdr*>(hdr)[stage].seq == it;
            if (!ready) {
                if (np) {
                    __syncwarp();
                    if (lane == 0) set_flags(pend, np, epoch);
                    np = 0;
                }
                for (;;) {
                    // experiment 8: plain timed polling instead of try_wait (whose SYNCS wake-up fires on every mbarrier
                    // event of the CTA: ~100 wake-ups per visit in the ncu source view)
                    if (P.experiment == 8) {
                        while (!mbar_test(full + stage, par)) __nanosleep(256);
                    } else
                        mbar_wait(full + stage, par);
                    if (reinterpret_cast<volatile StageHdr*>(hdr)[stage].seq == it) break;
                    __nanosleep(64);
                }
            }
            const StageHdr h = hdr[stage];
            if (!h.valid) break;
            if (P.experiment == 2) k1 = clock64();
            // rows of the stage, already shifted by the row's alignment offset and by this lane's first wavelength: inside
            // the wavelength loop every operand is then one LDS with an immediate offset
            const unsigned char* rb = data + (size_t)stage * 8 * rowb;
            const double* r0 = reinterpret_cast<const double*>(rb) + (h.offs & 1u) + lane;
            const double* r1 = reinterpret_cast<const double*>(rb + rowb) + ((h.offs >> 1) & 1u) + lane;
            const double* r2 = reinterpret_cast<const double*>(rb + 2 * rowb) + ((h.offs >> 2) & 1u) + lane;
            const double* r3 = reinterpret_cast<const double*>(rb + 3 * rowb) + ((h.offs >> 3) & 1u) + lane;
            const double* r4 = reinterpret_cast<const double*>(rb + 4 * rowb) + ((h.offs >> 4) & 1u) + lane;
            const double* r5 = reinterpret_cast<const double*>(rb + 5 * rowb) + ((h.offs >> 5) & 1u) + lane;
            const double* r6 = reinterpret_cast<const double*>(rb + 6 * rowb) + ((h.offs >> 6) & 1u) + lane;
            const double* r7 = reinterpret_cast<const double*>(rb + 7 * rowb) + ((h.offs >> 7) & 1u) + lane;
            double* out = h.dst + lane;
            const bool z1 = (h.offs >> 12) & 1u, z2 = (h.offs >> 15) & 1u;
            auto item = [&](int o) {
                const double a_c = r0[o], S_c = r1[o];
                const double I_1 = z1 ? 0.0 : r4[o];
                const double I_2 = z2 ? 0.0 : r7[o];
                double a, b, e;
                linear_weights(h.hr1 * (a_c + r2[o]), a, b, e);
                double I = 0.0 + (e * I_1 + a * r3[o] + b * S_c) * h.w1;
                linear_weights(h.hr2 * (a_c + r5[o]), a, b, e);
                I += (e * I_2 + a * r6[o] + b * S_c) * h.w2;
                out[o] = I;
            };
            for (int o = 0; lane + o < nlam; o += 32) item(o);
            __syncwarp();   // every lane has read its part of the stage and issued its stores
            if (lane == 0) {
                __threadfence_block();   // release: the warp's reads of the stage are ordered before the hand-back
                consumed[stage] = it / (unsigned)ns + 1u;
            }
            pend[np++] = h.flag;
            long long k2 = 0;
            if (P.experiment == 2) k2 = clock64();
            if (np == TMA_DEFER) {
                if (lane == 0) set_flags(pend, TMA_DEFER, epoch);
                np = 0;
            }
            if (P.experiment == 2 && lane == 0) {
                atomicAdd(P.prof + 5, (unsigned long long)(k1 - k0));            // consumer: waiting for data
                atomicAdd(P.prof + 6, (unsigned long long)(k2 - k1));            // consumer: compute + stores
                atomicAdd(P.prof + 7, (unsigned long long)(clock64() - k2));     // consumer: flag release
                atomicAdd(P.prof + 8, 1ull);
            }
        }
        __syncwarp();
        if (lane == 0) set_flags(pend, np, epoch);
    }
}

// events of the launches since the last sweep_collect(): (start, stop) pairs from a pool owned by the grid
struct SweepTimers {
    std::vector<cudaEvent_t> pool;   // 2 per launch
    size_t used = 0;
    ~SweepTimers() {
        for (auto e : pool) cudaEventDestroy(e);
    }
};
static SweepTimers* timers_of(vrt_grid* g) {
    if (!g->sweep_timers) g->sweep_timers = new SweepTimers();
    return static_cast<SweepTimers*>(g->sweep_timers);
}
void sweep_timers_free(void* p) { delete static_cast<SweepTimers*>(p); }

// Adds the kernel time of the launches recorded since the last call to stats->sweep_ms.  The stream must have been
// synchronised by the caller (no launch of this grid in flight).
int sweep_collect(vrt_grid* g, SweepStats* stats) {
    std::lock_guard<std::recursive_mutex> lock(g->mu);
    SweepTimers* T = timers_of(g);
    for (size_t i = 0; i + 1 < T->used; i += 2) {
        float ms = 0;
        VRT_CUDA(cudaEventElapsedTime(&ms, T->pool[i], T->pool[i + 1]));
        if (stats) stats->sweep_ms += ms;
    }
    T->used = 0;
    return VRT_OK;
}

// Launches the merged sweep program asynchronously on `st` (no host synchronisation, no allocation in steady state).
// The kernel time is accounted for by sweep_collect() after the caller has synchronised.
int sweep_run(vrt_grid* g, int nd, const SweepDir* dirs, const double* S, int64_t ldS, int64_t nlam, cudaStream_t st, SweepStats* stats) {
    if (nd <= 0) return VRT_OK;
    if (nd > MAX_DIRS) {
        set_error("sweep_run: %d directions exceed MAX_DIRS=%d", nd, MAX_DIRS);
        return VRT_E_INVALID;
    }
    int T = 0;
    double visits = 0;
    for (int d = 0; d < nd; d++) {
        T = std::max<int>(T, (int)dirs[d].sch->step_off.size() - 1);
        visits += (double)dirs[d].sch->n_visits;
        if (dirs[d].sch->n_sweeps != dirs[0].sch->n_sweeps) {
            set_error("sweep_run: schedules of one launch must share n_sweeps");
            return VRT_E_INVALID;
        }
    }
    if (T == 0) return VRT_OK;

    // the flag pool and the epoch belong to the grid, which solvers and host threads may share: one launch at a time
    // carves its flags and takes its epoch (launches of one grid are also serialised on the device: the flags of a
    // launch must not be reused before it has finished, so all launches of a grid go through one stream order)
    std::lock_guard<std::recursive_mutex> lock(g->mu);
    if (g->sweep_stream_set && g->sweep_stream != st) VRT_CUDA(cudaStreamSynchronize(g->sweep_stream));
    g->sweep_stream = st;
    g->sweep_stream_set = true;

    // ready-flags: one int per chunk, carved from a pool that is never cleared (flag == epoch <=> done now)
    const int cv = dirs[0].sch->cv;
    size_t nflags = 0;
    for (int d = 0; d < nd; d++) {
        if (dirs[d].sch->cv != cv) {
            set_error("sweep_run: schedules of one launch must share the chunk size");
            return VRT_E_INVALID;
        }
        nflags += (size_t)dirs[d].sch->n_chunks;
    }
    if (g->flag_pool.n < nflags || g->epoch >= INT32_MAX - 1) {
        VRT_CUDA(cudaStreamSynchronize(st));   // an earlier launch may still be using the old pool
        VRT_TRY(g->flag_pool.alloc(nflags));
        VRT_CUDA(cudaMemsetAsync(g->flag_pool.p, 0, sizeof(int32_t) * g->flag_pool.n, st));
        g->epoch = 0;
    }
    const int32_t epoch = ++g->epoch;
    DirTable DT;
    memset(&DT, 0, sizeof(DT));
    size_t fo = 0;
    for (int d = 0; d < nd; d++) {
        DirDev& h = DT.d[d];
        h.flags = g->flag_pool.p + fo;
        fo += (size_t)dirs[d].sch->n_chunks;
        h.step_off = dirs[d].sch->step_off_dev.p;
        h.nsteps = (int32_t)dirs[d].sch->step_off.size() - 1;
        h.pad = 0;
        h.visits = dirs[d].sch->visits.p;
        h.alpha = dirs[d].alpha;
        h.I_main = dirs[d].I_main;
        for (int s = 0; s < MAX_SWEEPS; s++) h.scratch[s] = dirs[d].scratch[s];
    }

    SweepParams P;
    P.S = S;
    P.ldS = ldS;
    P.nd = nd;
    P.T = T;
    P.nlam = (int)nlam;
    P.cv = cv;
    P.epoch = epoch;
    P.run_len = getenv("VRT_RUN_LEN") ? std::max(1, std::min(32, atoi(getenv("VRT_RUN_LEN")))) : 32;
    P.experiment = getenv("VRT_EXPERIMENT") ? atoi(getenv("VRT_EXPERIMENT")) : 0;
    P.pacing = getenv("VRT_PACING") ? atoi(getenv("VRT_PACING")) : 0;
    P.prof = nullptr;
    DevBuf<unsigned long long> d_prof;
    if (P.experiment == 2) {
        VRT_TRY(d_prof.alloc(16));
        VRT_CUDA(cudaMemsetAsync(d_prof.p, 0, 16 * sizeof(unsigned long long), st));
        P.prof = d_prof.p;
    }

    int dev = 0, sms = 0, per_sm = 0, smem_max = 0;
    VRT_CUDA(cudaGetDevice(&dev));
    VRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VRT_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const char* env_tma = getenv("VRT_TMA");
    bool use_tma = P.cv == 1 && nlam >= 16 && !(env_tma && atoi(env_tma) == 0);
    const void* fn = nullptr;
    int block = 0, ns = 0, rowb = 0;
    size_t shmem = 0;
    if (use_tma) {
        const char* env_cfg = getenv("VRT_TMA_CFG");
        const int cfg = env_cfg ? atoi(env_cfg) : 27;
        switch (cfg) {
            case 26: fn = (const void*)k_sweep_tma<2, 6>; block = 32 * 8; break;
            case 36: fn = (const void*)k_sweep_tma<3, 6>; block = 32 * 9; break;
            case 35: fn = (const void*)k_sweep_tma<3, 5>; block = 32 * 8; break;
            case 44: fn = (const void*)k_sweep_tma<4, 4>; block = 32 * 8; break;
            case 17: fn = (const void*)k_sweep_tma<1, 7>; block = 32 * 8; break;
            default: fn = (const void*)k_sweep_tma<2, 7>; block = 32 * 9; break;
        }
        rowb = (int)(((nlam + 1) * 8 + 15) & ~15);
        const size_t fixed = 2 * TMA_MAX_STAGES * sizeof(uint64_t) + TMA_MAX_STAGES * sizeof(StageHdr) + MAX_DIRS * sizeof(DirDev);
        const char* env_ns = getenv("VRT_TMA_STAGES");
        const char* env_kb = getenv("VRT_TMA_SMEM_KB");
        const size_t budget = (size_t)(env_kb && atoi(env_kb) > 0 ? atoi(env_kb) : 72) * 1024;
        ns = (int)((budget - fixed) / ((size_t)8 * rowb));
        if (env_ns && atoi(env_ns) > 0) ns = atoi(env_ns);
        ns = std::max(2, std::min(ns, TMA_MAX_STAGES));
        shmem = fixed + (size_t)ns * 8 * rowb;
        // very wide rows (vrt_formal_solve with hundreds of wavelengths): two stages no longer fit one CTA -> register path
        if (shmem > (size_t)smem_max) use_tma = false;
        else VRT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    }
    if (!use_tma) {
        fn = P.cv == 1 ? (const void*)k_sweep<true> : (const void*)k_sweep<false>;
        block = SWEEP_BLOCK;
        shmem = 0;
    }
    VRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, shmem));
    if (per_sm < 1) {
        set_error("sweep kernel cannot be resident");
        return VRT_E_CUDA;
    }
    int grid = sms * per_sm;
    SweepTimers* TM = timers_of(g);
    if (TM->used + 2 > TM->pool.size()) {
        for (int i = 0; i < 2; i++) {
            cudaEvent_t e;
            VRT_CUDA(cudaEventCreate(&e));
            TM->pool.push_back(e);
        }
    }
    cudaEvent_t e0 = TM->pool[TM->used], e1 = TM->pool[TM->used + 1];
    VRT_CUDA(cudaEventRecord(e0, st));
    void* args[] = {(void*)&P, (void*)&DT, (void*)&ns, (void*)&rowb};
    cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(block), args, shmem, st);
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchCooperativeKernel(k_sweep)", __FILE__, __LINE__);
    VRT_CUDA(cudaEventRecord(e1, st));
    TM->used += 2;
    if (P.experiment == 2) {
        VRT_CUDA(cudaStreamSynchronize(st));
        unsigned long long h[16];
        VRT_CUDA(cudaMemcpy(h, d_prof.p, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[sweep prof] producer lane: flags %.0f cyc, stage %.0f cyc per visit (%llu visits) | run %.0f cyc (%llu runs) | "
                        "consumer per visit: wait %.0f, compute %.0f, release %.0f cyc (%llu)\n",
                (double)h[0] / (h[2] + 1e-9), (double)h[1] / (h[2] + 1e-9), h[2], (double)h[3] / (h[4] + 1e-9), h[4],
                (double)h[5] / (h[8] + 1e-9), (double)h[6] / (h[8] + 1e-9), (double)h[7] / (h[8] + 1e-9), h[8]);
    }
    if (stats) {
        stats->kernels += 1;
        stats->visits += visits;
        stats->steps += (double)T * nd;
    }
    return VRT_OK;
}

}  // namespace vrt
