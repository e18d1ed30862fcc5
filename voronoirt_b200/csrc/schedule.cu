// schedule.cu — K3: turns the sequential in-layer Gauss–Seidel sweep of Delaunay_upII/downII
// (reference src/irregular_ray_tracing.jl:37-80, :118-161) into a parallel program that produces the same
// numbers:
//
//   1. classify each upwind reference of each processed cell: FINAL (lower layer), THIS (same layer,
//      earlier in the loop order: reads this sweep's value), LAG (same layer, later: reads the previous
//      sweep's value, 0 in sweep 1), ZERO (higher layer or the never-processed last-rank site, Q1/Q5);
//   2. chg[s][c] = "the value of c computed in sweep s can differ from sweep s-1" (structural, exact):
//      chg[1] = processed; chg[s] = OR_LAG chg[s-1][u]  OR_THIS chg[s][u].  stab(c) = #s with chg[s][c].
//      With prune=0 every processed cell is visited n_sweeps times like the reference does;
//   3. a VISIT per (c, s <= stab(c)); every operand is resolved to (buffer,row): sweep-s values of cells with
//      stab > s live in scratch buffer s, the final value in the main intensity array, so no value is ever
//      overwritten and ANY topological order of the visits reproduces the sequential result;
//   4. the visits are ordered by their level in the dependency DAG (level = 1 + deepest producer read) and
//      chunked; each chunk records the chunks that produce its two upwind intensities (dataflow sweep);
//   5. (wide rows, optional) BLOCKED order: the visits are keyed (level slab, column block in upwind order, level) and
//      every key is pushed behind the keys of its producers, key(v) = max(base(v), key(p) + 1) — still a topological
//      order, so the values do not change by a bit, but what one visit writes is read again while it is still in L2.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <string.h>
#include "vrt_internal.h"

namespace vrt {

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

struct DirCtx {
    const int32_t* layer;  // 1-based layer of each internal cell
    const int32_t* ord;    // rank in the direction's perm (nullptr: up, ord = c)
    int down;
    int32_t X;             // never-processed cell (last rank)
    int64_t n;
};

__device__ __forceinline__ bool processed(const DirCtx& d, int64_t c) { return d.layer[c] >= 2 && c != d.X; }
// true when u is visited before c inside one sweep of their common layer (up ascending rank, down descending)
__device__ __forceinline__ bool before(const DirCtx& d, int32_t u, int32_t c) {
    if (!d.down) return u < c;
    return d.ord[u] > d.ord[c];
}

__global__ void k_classify(DirCtx d, const int32_t* __restrict__ up, int32_t* __restrict__ cls, int* __restrict__ bad) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= d.n) return;
    if (!processed(d, c)) {
        cls[2 * c] = cls[2 * c + 1] = -1;
        return;
    }
    int32_t lc = d.layer[c];
    for (int m = 0; m < 2; m++) {
        int32_t u = up[2 * c + m];
        int32_t k;
        if (u < 0) k = CLS_ZERO;  // no positive neighbour at all (the reference would index garbage)
        else if (u == (int32_t)c) { k = CLS_ZERO; atomicExch(bad, 1); }
        else if (u == d.X) k = CLS_ZERO;
        else if (d.layer[u] < lc) k = CLS_FINAL;
        else if (d.layer[u] > lc) k = CLS_ZERO;
        else k = before(d, u, (int32_t)c) ? CLS_THIS : CLS_LAG;
        cls[2 * c + m] = k;
    }
}

__global__ void k_chg_init(DirCtx d, uint8_t* __restrict__ chg) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c < d.n) chg[c] = processed(d, c) ? 1 : 0;
}

// base of sweep s: OR over LAG references of chg[s-1]
__global__ void k_chg_base(int64_t n, const int32_t* __restrict__ up, const int32_t* __restrict__ cls,
                           const uint8_t* __restrict__ prev, uint8_t* __restrict__ cur) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    uint8_t v = 0;
    for (int m = 0; m < 2; m++)
        if (cls[2 * c + m] == CLS_LAG && prev[up[2 * c + m]]) v = 1;
    cur[c] = v;
}

// propagate along THIS references until a fixed point (monotone, so in-place Jacobi is safe)
__global__ void k_chg_relax(int64_t n, const int32_t* __restrict__ up, const int32_t* __restrict__ cls,
                            uint8_t* cur, int* __restrict__ changed) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n || cur[c]) return;
    for (int m = 0; m < 2; m++)
        if (cls[2 * c + m] == CLS_THIS && cur[up[2 * c + m]]) {
            cur[c] = 1;
            *changed = 1;
            return;
        }
}

__global__ void k_sub_init(int64_t n, const uint8_t* __restrict__ chg, int32_t* __restrict__ sub) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c < n) sub[c] = chg[c] ? 1 : 0;
}

__global__ void k_sub_relax(int64_t n, const int32_t* __restrict__ up, const int32_t* __restrict__ cls,
                            const uint8_t* __restrict__ chg, int32_t* sub, int* __restrict__ changed) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n || !chg[c]) return;
    int32_t s = sub[c], want = 1;
    for (int m = 0; m < 2; m++)
        if (cls[2 * c + m] == CLS_THIS) {
            int32_t u = up[2 * c + m];
            if (chg[u]) {
                int32_t su = ((volatile int32_t*)sub)[u] + 1;
                want = su > want ? su : want;
            }
        }
    if (want > s) {
        sub[c] = want;
        *changed = 1;
    }
}

__global__ void k_nsub(int64_t n, const int32_t* __restrict__ layer, const int32_t* __restrict__ sub, int n_sweeps, int s,
                       int32_t* __restrict__ nsub) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int32_t v = sub[c];
    if (v > 0) atomicMax(&nsub[(layer[c] - 2) * n_sweeps + (s - 1)], v);
}

// DAG level of every visit (c, s <= stab(c)): 1 + the deepest producer it reads.  Jacobi relaxation to the
// longest-path fixed point (levels only grow).  lev is [S][n]; the producer rules are those of producer() below.
__global__ void k_level_relax(int64_t n, int S, const int32_t* __restrict__ up, const int32_t* __restrict__ cls,
                              const int32_t* __restrict__ stab, int32_t* lev, int* __restrict__ changed) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int st = stab[c];
    bool ch = false;
    for (int s = 1; s <= st; s++) {
        int32_t want = 1;
        for (int m = 0; m < 2; m++) {
            const int32_t k = cls[2 * c + m];
            if (k != CLS_FINAL && k != CLS_THIS && k != CLS_LAG) continue;
            const int32_t u = up[2 * c + m];
            const int su = stab[u];
            int t = k == CLS_FINAL ? su : (k == CLS_THIS ? s : s - 1);
            if (su <= 0 || t <= 0) continue;
            if (t > su) t = su;
            const int32_t lu = ((volatile int32_t*)lev)[(int64_t)(t - 1) * n + u] + 1;
            want = lu > want ? lu : want;
        }
        int32_t* p = lev + (int64_t)(s - 1) * n + c;
        if (want > *p) {
            *p = want;
            ch = true;
        }
    }
    if (ch) *changed = 1;
}

__global__ void k_level_init(int64_t n, int S, const int32_t* __restrict__ stab, int32_t* __restrict__ lev) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int s = 1; s <= S; s++) lev[(int64_t)(s - 1) * n + c] = s <= stab[c] ? 1 : 0;
}

__global__ void k_level_max(int64_t n, int S, const int32_t* __restrict__ lev, int32_t* __restrict__ out) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int32_t m = 0;
    if (c < n)
        for (int s = 0; s < S; s++) m = max(m, lev[(int64_t)s * n + c]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

__global__ void k_stab_acc(int64_t n, const uint8_t* __restrict__ chg, int32_t* __restrict__ stab, int first) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c < n) stab[c] = (first ? 0 : stab[c]) + (chg[c] ? 1 : 0);
}

__global__ void k_flag_gt(int64_t n, const int32_t* __restrict__ stab, int s, int32_t* __restrict__ flag) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c < n) flag[c] = stab[c] > s ? 1 : 0;
}

// emit (key = local step, val = sweep<<29 | cell) for every visit, in cell order
__global__ void k_emit(int64_t n, const int32_t* __restrict__ stab, const int64_t* __restrict__ voff,
                       const int32_t* __restrict__ lev /* [s][n] */, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int32_t st = stab[c];
    int64_t o = voff[c];
    for (int s = 1; s <= st; s++) {
        int32_t t = lev[(int64_t)(s - 1) * n + c] - 1;
        key[o + s - 1] = (uint32_t)t;
        val[o + s - 1] = ((uint32_t)s << SEL_SHIFT) | (uint32_t)c;
    }
}

// ---------------------------------------------------------------- blocked (cache-aware) order
// key of a visit: [63:48] level slab  [47:24] rank of the cell's column block in upwind order  [23:0] level
constexpr int KEY_LEV_BITS = 24, KEY_BLK_BITS = 24;

// rank (in upwind order) of the column block that holds each cell; pos is [c][3] = (z, x, y)
__global__ void k_block_rank(int64_t n, const double* __restrict__ pos, double x0, double sx, int bx, double y0, double sy, int by,
                             const int32_t* __restrict__ rank_of_block, int32_t* __restrict__ brank) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int ix = (int)floor((pos[3 * c + 1] - x0) * sx), iy = (int)floor((pos[3 * c + 2] - y0) * sy);
    ix = min(max(ix, 0), bx - 1);
    iy = min(max(iy, 0), by - 1);
    brank[c] = rank_of_block[ix * by + iy];
}

__global__ void k_key_init(int64_t n, int S, const int32_t* __restrict__ stab, const int32_t* __restrict__ lev,
                           const int32_t* __restrict__ brank, int slab, unsigned long long* __restrict__ key) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int st = stab[c];
    for (int s = 1; s <= S; s++) {
        unsigned long long k = 0;
        if (s <= st) {
            const unsigned long long l = (unsigned long long)lev[(int64_t)(s - 1) * n + c];
            const unsigned long long tau = slab > 0 ? (l - 1) / (unsigned long long)slab : 0ull;
            k = (tau << (KEY_LEV_BITS + KEY_BLK_BITS)) | ((unsigned long long)brank[c] << KEY_LEV_BITS) | l;
        }
        key[(int64_t)(s - 1) * n + c] = k;
    }
}

// key(v) = max(key(v), key(producer) + 1): monotone, so the in-place (chaotic) iteration reaches the least fixed point.
// The producers are those of k_level_relax / producer() below.
__global__ void k_key_relax(int64_t n, const int32_t* __restrict__ up, const int32_t* __restrict__ cls,
                            const int32_t* __restrict__ stab, unsigned long long* key, int* __restrict__ changed) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int st = stab[c];
    bool ch = false;
    for (int s = 1; s <= st; s++) {
        unsigned long long* p = key + (int64_t)(s - 1) * n + c;
        const unsigned long long cur = *p;
        unsigned long long want = cur;
        for (int m = 0; m < 2; m++) {
            const int32_t k = cls[2 * c + m];
            if (k != CLS_FINAL && k != CLS_THIS && k != CLS_LAG) continue;
            const int32_t u = up[2 * c + m];
            const int su = stab[u];
            int t = k == CLS_FINAL ? su : (k == CLS_THIS ? s : s - 1);
            if (su <= 0 || t <= 0) continue;
            if (t > su) t = su;
            const unsigned long long ku = ((volatile unsigned long long*)key)[(int64_t)(t - 1) * n + u] + 1ull;
            want = ku > want ? ku : want;
        }
        if (want > cur) {
            *p = want;
            ch = true;
        }
    }
    if (ch) *changed = 1;
}

__global__ void k_emit64(int64_t n, const int32_t* __restrict__ stab, const int64_t* __restrict__ voff,
                         const unsigned long long* __restrict__ vkey /* [s][n] */, unsigned long long* __restrict__ key,
                         uint32_t* __restrict__ val) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n) return;
    int32_t st = stab[c];
    int64_t o = voff[c];
    for (int s = 1; s <= st; s++) {
        key[o + s - 1] = vkey[(int64_t)(s - 1) * n + c];
        val[o + s - 1] = ((uint32_t)s << SEL_SHIFT) | (uint32_t)c;
    }
}

__global__ void k_key_heads(int64_t V, const unsigned long long* __restrict__ key, uint8_t* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < V) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

__global__ void k_step_bounds(int64_t V, const uint32_t* __restrict__ key, int64_t* __restrict__ off) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= V) return;
    if (i == 0 || key[i] != key[i - 1]) off[key[i]] = i;
}

__device__ __forceinline__ uint32_t resolve(int32_t u, int t, const int32_t* __restrict__ stab,
                                            const int32_t* __restrict__ slots /* [s][n] */, int64_t n) {
    if (t <= 0) return SEL_ZERO << SEL_SHIFT;
    if (t >= stab[u]) return (SEL_MAIN << SEL_SHIFT) | (uint32_t)u;
    return ((SEL_SCR0 + (uint32_t)(t - 1)) << SEL_SHIFT) | (uint32_t)slots[(int64_t)(t - 1) * n + u];
}

// padded position of sorted visit i: every step is padded to a multiple of cv visits
__device__ __forceinline__ int64_t padded_pos(int64_t i, uint32_t step, const int64_t* __restrict__ step_off,
                                              const int64_t* __restrict__ pstep_off) {
    return pstep_off[step] + (i - step_off[step]);
}

// vpos[voff[c] + s - 1] = padded position of visit (c, s)
__global__ void k_visit_pos(int64_t V, const uint32_t* __restrict__ key, const uint32_t* __restrict__ val,
                            const int64_t* __restrict__ voff, const int64_t* __restrict__ step_off,
                            const int64_t* __restrict__ pstep_off, int32_t* __restrict__ vpos) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= V) return;
    uint32_t v = val[i];
    int s = (int)(v >> SEL_SHIFT);
    int32_t c = (int32_t)(v & ROW_MASK);
    vpos[voff[c] + s - 1] = (int32_t)padded_pos(i, key[i], step_off, pstep_off);
}

// chunk that produces the value of cell u at sweep t (t clipped to stab(u)); DEP_NONE for boundary / unprocessed cells
__device__ __forceinline__ uint32_t producer(int32_t u, int t, const int32_t* __restrict__ stab, const int64_t* __restrict__ voff,
                                             const int32_t* __restrict__ vpos, int cv) {
    int su = stab[u];
    if (su <= 0 || t <= 0) return DEP_NONE;
    if (t > su) t = su;
    return (uint32_t)(vpos[voff[u] + t - 1] / cv);
}

__global__ void k_build_visits(int64_t V, int64_t n, const uint32_t* __restrict__ key, const uint32_t* __restrict__ val,
                               const int32_t* __restrict__ up,
                               const int32_t* __restrict__ cls, const int32_t* __restrict__ stab,
                               const int32_t* __restrict__ slots, const double* __restrict__ w, const double* __restrict__ r,
                               const int64_t* __restrict__ voff, const int32_t* __restrict__ vpos,
                               const int64_t* __restrict__ step_off, const int64_t* __restrict__ pstep_off, int cv,
                               Visit* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= V) return;
    uint32_t v = val[i];
    int s = (int)(v >> SEL_SHIFT);
    int32_t c = (int32_t)(v & ROW_MASK);
    Visit o;
    o.cell = (uint32_t)c;
    o.dst = (s == stab[c]) ? ((SEL_MAIN << SEL_SHIFT) | (uint32_t)c)
                           : (((SEL_SCR0 + (uint32_t)(s - 1)) << SEL_SHIFT) | (uint32_t)slots[(int64_t)(s - 1) * n + c]);
    uint32_t src[2], uu[2], dep[2];
    for (int m = 0; m < 2; m++) {
        int32_t u = up[2 * c + m];
        int32_t k = cls[2 * c + m];
        uu[m] = u >= 0 ? (uint32_t)u : (uint32_t)c;
        dep[m] = DEP_NONE;
        if (k == CLS_FINAL) {
            src[m] = (SEL_MAIN << SEL_SHIFT) | (uint32_t)u;
            dep[m] = producer(u, MAX_SWEEPS, stab, voff, vpos, cv);
        } else if (k == CLS_THIS) {
            src[m] = resolve(u, s, stab, slots, n);
            dep[m] = producer(u, s, stab, voff, vpos, cv);
        } else if (k == CLS_LAG) {
            src[m] = resolve(u, s - 1, stab, slots, n);
            dep[m] = producer(u, s - 1, stab, voff, vpos, cv);
        } else
            src[m] = SEL_ZERO << SEL_SHIFT;
    }
    o.u1 = uu[0]; o.u2 = uu[1];
    o.src1 = src[0]; o.src2 = src[1];
    o.dep1 = dep[0]; o.dep2 = dep[1];
    o.w1 = w[2 * c]; o.w2 = w[2 * c + 1];
    o.hr1 = 0.5 * r[2 * c]; o.hr2 = 0.5 * r[2 * c + 1];
    out[padded_pos(i, key[i], step_off, pstep_off)] = o;
}

// every visit must come after the chunks that produce its operands: the dataflow sweep spins on their flags and would
// never return otherwise
__global__ void k_check_topological(int64_t Vp, int cv, const Visit* __restrict__ v, int* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Vp || v[i].cell == CELL_DUMMY) return;
    const uint32_t me = (uint32_t)(i / cv);
    if ((v[i].dep1 != DEP_NONE && v[i].dep1 >= me) || (v[i].dep2 != DEP_NONE && v[i].dep2 >= me)) atomicExch(bad, 1);
}

// runs `launch` until a whole batch of `batch` relaxation sweeps changes nothing (one flag read per batch, not per sweep)
static int relax_loop(void (*launch)(void*), void* ctx, int* d_flag, int max_iter, int batch = 8) {
    for (int it = 0; it < max_iter; it += batch) {
        VRT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int)));
        for (int b = 0; b < batch; b++) launch(ctx);
        VRT_CUDA(cudaGetLastError());
        int h = 0;
        VRT_CUDA(cudaMemcpy(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (!h) return VRT_OK;
    }
    set_error("schedule relaxation did not converge");
    return VRT_E_STATE;
}

// order of the visits of wide-row programs (cv == 1): kx x ky column blocks in x, y taken in upwind order, slabs of
// `slab` levels (0: one slab), steps of at least `step_min` visits.  kx * ky == 1 and slab == 0: plain level order.
OrderCfg order_config(int64_t n, int cv) {
    OrderCfg oc;
    if (cv != 1) return oc;
    const char* e = getenv("VRT_BLOCKS");
    if (e) {
        int a = 1, b = 1;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && b >= 1 && a * b < (1 << KEY_BLK_BITS)) {
            oc.bx = a;
            oc.by = b;
        }
    }
    e = getenv("VRT_SLAB");
    if (e && atoi(e) > 0) oc.slab = atoi(e);
    e = getenv("VRT_STEP_MIN");
    if (e && atoi(e) > 0) oc.step_min = atoi(e);
    (void)n;
    return oc;
}

int schedule_build(vrt_grid* g, const double k[3], int down, int n_sweeps, double p, int prune, int cv, const OrderCfg& oc,
                   bool keep_introspection, DirSchedule** out) {
    *out = nullptr;
    if (n_sweeps < 1 || n_sweeps > MAX_SWEEPS - 1) {
        set_error("n_sweeps=%d unsupported (1..%d)", n_sweeps, MAX_SWEEPS - 1);
        return VRT_E_INVALID;
    }
    const int64_t n = g->n;
    const int bs = 256;
    const int nb = nblocks(n, bs);
    DirSchedule* sch = new DirSchedule();
    struct Guard { DirSchedule* s; ~Guard() { delete s; } } guard{sch};
    memcpy(sch->k, k, sizeof(double) * 3);
    sch->down = down; sch->n_sweeps = n_sweeps; sch->p = p; sch->prune = prune; sch->cv = cv;
    sch->order = oc;
    const bool blocked = cv == 1 && oc.blocked();

    Stencil st;
    VRT_TRY(grid_stencil(g, k, p, &st));

    DirCtx d;
    d.layer = down ? g->layer_dn.p : g->layer_up.p;
    d.ord = down ? g->rank_dn.p : nullptr;
    d.down = down;
    d.n = n;
    if (down) {
        VRT_CUDA(cudaMemcpy(&d.X, g->perm_dn_int.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    } else
        d.X = (int32_t)(n - 1);

    DevBuf<int> flag;
    VRT_TRY(flag.alloc(1));
    VRT_CUDA(cudaMemset(flag.p, 0, sizeof(int)));
    VRT_TRY(sch->cls.alloc(2 * n)); VRT_TRY(sch->sublevel.alloc(n)); VRT_TRY(sch->stab.alloc(n));
    k_classify<<<nb, bs>>>(d, st.up.p, sch->cls.p, flag.p);
    VRT_CUDA(cudaGetLastError());
    {
        int h = 0;
        VRT_CUDA(cudaMemcpy(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (h) {
            set_error("a site is its own upwind neighbour (periodic self-image); not supported");
            return VRT_E_GRID;
        }
    }

    const int S = n_sweeps;
    DevBuf<uint8_t> chg;      // [S][n]
    DevBuf<int32_t> lev;      // [S][n] DAG level of visit (c, s)
    VRT_TRY(chg.alloc((size_t)S * n)); VRT_TRY(lev.alloc((size_t)S * n));

    for (int s = 1; s <= S; s++) {
        uint8_t* cur = chg.p + (size_t)(s - 1) * n;
        if (s == 1 || !prune) {
            k_chg_init<<<nb, bs>>>(d, cur);
        } else {
            const uint8_t* prev = chg.p + (size_t)(s - 2) * n;
            k_chg_base<<<nb, bs>>>(n, st.up.p, sch->cls.p, prev, cur);
            struct C { int nb, bs; int64_t n; const int32_t *up, *cls; uint8_t* cur; int* flag; } cx{nb, bs, n, st.up.p, sch->cls.p, cur, flag.p};
            VRT_TRY(relax_loop([](void* v) { C* c = (C*)v; k_chg_relax<<<c->nb, c->bs>>>(c->n, c->up, c->cls, c->cur, c->flag); }, &cx, flag.p, 100000));
        }
        k_stab_acc<<<nb, bs>>>(n, cur, sch->stab.p, s == 1);
        if (s == 1) {
            // in-layer sub-levels of sweep 1 (SURVEY App. G rule 3): kept for introspection / tests only
            k_sub_init<<<nb, bs>>>(n, cur, sch->sublevel.p);
            struct C2 { int nb, bs; int64_t n; const int32_t *up, *cls; const uint8_t* chg; int32_t* sub; int* flag; } c2{nb, bs, n, st.up.p, sch->cls.p, cur, sch->sublevel.p, flag.p};
            VRT_TRY(relax_loop([](void* v) { C2* c = (C2*)v; k_sub_relax<<<c->nb, c->bs>>>(c->n, c->up, c->cls, c->chg, c->sub, c->flag); }, &c2, flag.p, 100000));
        }
        VRT_CUDA(cudaGetLastError());
    }
    // global topological levels over all (cell, sweep) visits: the program order of the dataflow sweep.  This is
    // 2.5-3.5x shallower than (layer, sweep, sub-level) because a visit only waits for what it really reads.
    k_level_init<<<nb, bs>>>(n, S, sch->stab.p, lev.p);
    {
        struct C3 { int nb, bs; int64_t n; int S; const int32_t *up, *cls, *stab; int32_t* lev; int* flag; } c3{nb, bs, n, S, st.up.p, sch->cls.p, sch->stab.p, lev.p, flag.p};
        VRT_TRY(relax_loop([](void* v) { C3* c = (C3*)v; k_level_relax<<<c->nb, c->bs>>>(c->n, c->S, c->up, c->cls, c->stab, c->lev, c->flag); }, &c3, flag.p, 1000000, 32));
    }
    int64_t T = 0;
    {
        VRT_CUDA(cudaMemset(flag.p, 0, sizeof(int)));
        k_level_max<<<nb, bs>>>(n, S, lev.p, flag.p);
        int h = 0;
        VRT_CUDA(cudaMemcpy(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
        T = h;
    }

    // visit offsets = exclusive scan of stab; scratch slots = exclusive scan of [stab > s]
    DevBuf<int64_t> voff;
    DevBuf<int32_t> slots, flags;
    DevBuf<char> tmp;
    VRT_TRY(voff.alloc(n + 1)); VRT_TRY(slots.alloc((size_t)(S > 1 ? S - 1 : 1) * n)); VRT_TRY(flags.alloc(n));
    size_t tb1 = 0, tb2 = 0;
    VRT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb1, sch->stab.p, voff.p, (int)n));
    VRT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, flags.p, slots.p, (int)n));
    VRT_TRY(tmp.alloc(tb1 > tb2 ? tb1 : tb2));
    size_t tb = tmp.n;
    VRT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, sch->stab.p, voff.p, (int)n));
    int64_t V = 0;
    {
        int64_t lastoff = 0;
        int32_t laststab = 0;
        VRT_CUDA(cudaMemcpy(&lastoff, voff.p + (n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost));
        VRT_CUDA(cudaMemcpy(&laststab, sch->stab.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
        V = lastoff + laststab;
    }
    for (int s = 1; s < S; s++) {
        k_flag_gt<<<nb, bs>>>(n, sch->stab.p, s, flags.p);
        tb = tmp.n;
        VRT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags.p, slots.p + (size_t)(s - 1) * n, (int)n));
        int32_t lo = 0, lf = 0;
        VRT_CUDA(cudaMemcpy(&lo, slots.p + (size_t)(s - 1) * n + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
        VRT_CUDA(cudaMemcpy(&lf, flags.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
        sch->scr_rows[s - 1] = (int64_t)lo + lf;
    }
    sch->n_visits = V;
    sch->step_off.assign((size_t)T + 1, 0);
    if (V > 0) {
        if (V >= (int64_t)INT32_MAX) {
            set_error("too many visits (%lld) for one direction", (long long)V);
            return VRT_E_INVALID;
        }
        DevBuf<uint32_t> key, val, key2, val2;
        VRT_TRY(key.alloc(V)); VRT_TRY(val.alloc(V)); VRT_TRY(key2.alloc(V)); VRT_TRY(val2.alloc(V));
        DevBuf<int64_t> off, poff;
        if (!blocked) {
            k_emit<<<nb, bs>>>(n, sch->stab.p, voff.p, lev.p, key.p, val.p);
            VRT_CUDA(cudaGetLastError());
            int bits = 1;
            while ((1ll << bits) < T + 1 && bits < 32) bits++;
            size_t sb = 0;
            VRT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, key.p, key2.p, val.p, val2.p, (int)V, 0, bits));
            DevBuf<char> stmp;
            VRT_TRY(stmp.alloc(sb));
            VRT_CUDA(cub::DeviceRadixSort::SortPairs(stmp.p, sb, key.p, key2.p, val.p, val2.p, (int)V, 0, bits));
            VRT_TRY(off.alloc(T + 1)); VRT_TRY(poff.alloc(T + 1));
            k_step_bounds<<<nblocks(V, bs), bs>>>(V, key2.p, off.p);
            std::vector<int64_t> soff((size_t)T + 1, 0);
            VRT_CUDA(cudaMemcpy(soff.data(), off.p, sizeof(int64_t) * T, cudaMemcpyDeviceToHost));
            soff[T] = V;
            VRT_CUDA(cudaMemcpy(off.p + T, &V, sizeof(int64_t), cudaMemcpyHostToDevice));
            // pad every step to whole chunks of cv visits: a chunk never straddles two dependent steps
            for (int64_t t = 0; t < T; t++) sch->step_off[t + 1] = sch->step_off[t] + (soff[t + 1] - soff[t] + cv - 1) / cv * cv;
        } else {
            // ---- blocked order (rule 5).  Column blocks ranked by the projection of their centre on the horizontal part
            // of -k (k points upwind): upwind blocks first.
            if (T + 1 >= (1ll << KEY_LEV_BITS) - (1 << 20)) {
                set_error("blocked order: too many dependency levels (%lld)", (long long)T);
                return VRT_E_INVALID;
            }
            const int nblk = oc.bx * oc.by;
            std::vector<double> proj(nblk);
            const double x0 = g->bounds[2], x1 = g->bounds[3], y0 = g->bounds[4], y1 = g->bounds[5];
            for (int ix = 0; ix < oc.bx; ix++)
                for (int iy = 0; iy < oc.by; iy++)
                    proj[ix * oc.by + iy] = -(k[1] * (x0 + (ix + 0.5) * (x1 - x0) / oc.bx) + k[2] * (y0 + (iy + 0.5) * (y1 - y0) / oc.by));
            std::vector<int32_t> ord(nblk), rank(nblk);
            std::iota(ord.begin(), ord.end(), 0);
            std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return proj[a] < proj[b]; });
            for (int i = 0; i < nblk; i++) rank[ord[i]] = i;
            DevBuf<int32_t> d_rank, brank;
            VRT_TRY(d_rank.alloc(nblk)); VRT_TRY(brank.alloc(n));
            VRT_CUDA(cudaMemcpy(d_rank.p, rank.data(), sizeof(int32_t) * nblk, cudaMemcpyHostToDevice));
            k_block_rank<<<nb, bs>>>(n, g->pos.p, x0, oc.bx / (x1 - x0), oc.bx, y0, oc.by / (y1 - y0), oc.by, d_rank.p, brank.p);
            DevBuf<unsigned long long> vkey;
            VRT_TRY(vkey.alloc((size_t)S * n));
            k_key_init<<<nb, bs>>>(n, S, sch->stab.p, lev.p, brank.p, oc.slab, vkey.p);
            VRT_CUDA(cudaGetLastError());
            {
                struct C4 { int nb, bs; int64_t n; const int32_t *up, *cls, *stab; unsigned long long* key; int* flag; } c4{nb, bs, n, st.up.p, sch->cls.p, sch->stab.p, vkey.p, flag.p};
                VRT_TRY(relax_loop([](void* v) { C4* c = (C4*)v; k_key_relax<<<c->nb, c->bs>>>(c->n, c->up, c->cls, c->stab, c->key, c->flag); }, &c4, flag.p, 1000000, 32));
            }
            DevBuf<unsigned long long> key64, key64b;
            VRT_TRY(key64.alloc(V)); VRT_TRY(key64b.alloc(V));
            k_emit64<<<nb, bs>>>(n, sch->stab.p, voff.p, vkey.p, key64.p, val.p);
            VRT_CUDA(cudaGetLastError());
            vkey.release();
            size_t sb = 0;
            VRT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, key64.p, key64b.p, val.p, val2.p, (int)V, 0, 64));
            DevBuf<char> stmp;
            VRT_TRY(stmp.alloc(sb));
            VRT_CUDA(cub::DeviceRadixSort::SortPairs(stmp.p, sb, key64.p, key64b.p, val.p, val2.p, (int)V, 0, 64));
            // steps: runs of equal keys (mutually independent visits), merged to at least step_min visits.  The sweep kernel
            // does not need independence inside a step (every visit waits for its own producers), steps only pace the
            // interleaving of the directions in flight.
            DevBuf<uint8_t> head;
            DevBuf<int32_t> hpos, nsel;
            VRT_TRY(head.alloc(V)); VRT_TRY(hpos.alloc(V)); VRT_TRY(nsel.alloc(1));
            k_key_heads<<<nblocks(V, bs), bs>>>(V, key64b.p, head.p);
            size_t selb = 0;
            thrust::counting_iterator<int32_t> cnt(0);
            VRT_CUDA(cub::DeviceSelect::Flagged(nullptr, selb, cnt, head.p, hpos.p, nsel.p, (int)V));
            VRT_TRY(stmp.ensure(selb));
            VRT_CUDA(cub::DeviceSelect::Flagged(stmp.p, selb, cnt, head.p, hpos.p, nsel.p, (int)V));
            int32_t nh = 0;
            VRT_CUDA(cudaMemcpy(&nh, nsel.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
            std::vector<int32_t> hp((size_t)nh);
            VRT_CUDA(cudaMemcpy(hp.data(), hpos.p, sizeof(int32_t) * (size_t)nh, cudaMemcpyDeviceToHost));
            sch->step_off.assign(1, 0);
            for (int32_t i = 1; i < nh; i++)
                if (hp[i] - sch->step_off.back() >= oc.step_min) sch->step_off.push_back(hp[i]);
            if (V - sch->step_off.back() < oc.step_min / 2 && sch->step_off.size() > 1) sch->step_off.pop_back();
            sch->step_off.push_back(V);
            T = (int64_t)sch->step_off.size() - 1;
            // visits are unpadded (cv == 1): position = index in the sorted order; k_visit_pos / k_build_visits see one step
            VRT_CUDA(cudaMemset(key2.p, 0, sizeof(uint32_t) * (size_t)V));
            VRT_TRY(off.alloc(2)); VRT_TRY(poff.alloc(2));
            const int64_t two[2] = {0, V};
            VRT_CUDA(cudaMemcpy(off.p, two, sizeof(two), cudaMemcpyHostToDevice));
        }
        const int64_t Vp = sch->step_off.back();
        if (Vp >= (int64_t)INT32_MAX) {
            set_error("too many visit slots (%lld) for one direction", (long long)Vp);
            return VRT_E_INVALID;
        }
        if (blocked) {
            const int64_t two[2] = {0, Vp};
            VRT_CUDA(cudaMemcpy(poff.p, two, sizeof(two), cudaMemcpyHostToDevice));
        } else
            VRT_CUDA(cudaMemcpy(poff.p, sch->step_off.data(), sizeof(int64_t) * (T + 1), cudaMemcpyHostToDevice));
        sch->n_slots = Vp;
        sch->n_chunks = Vp / cv;
        DevBuf<int32_t> vpos;
        VRT_TRY(vpos.alloc(V));
        k_visit_pos<<<nblocks(V, bs), bs>>>(V, key2.p, val2.p, voff.p, off.p, poff.p, vpos.p);
        VRT_TRY(sch->visits.alloc(Vp));
        VRT_CUDA(cudaMemset(sch->visits.p, 0xff, sizeof(Visit) * (size_t)Vp));  // cell = CELL_DUMMY everywhere, then fill
        k_build_visits<<<nblocks(V, bs), bs>>>(V, n, key2.p, val2.p, st.up.p, sch->cls.p, sch->stab.p, slots.p, st.w.p, st.r.p,
                                               voff.p, vpos.p, off.p, poff.p, cv, sch->visits.p);
        VRT_CUDA(cudaGetLastError());
        VRT_CUDA(cudaMemset(flag.p, 0, sizeof(int)));
        k_check_topological<<<nblocks(Vp, bs), bs>>>(Vp, cv, sch->visits.p, flag.p);
        int h = 0;
        VRT_CUDA(cudaMemcpy(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (h) {
            set_error("internal error: the sweep program is not in topological order");
            return VRT_E_STATE;
        }
    }
    {
        std::vector<int32_t> so(sch->step_off.size());
        for (size_t i = 0; i < so.size(); i++) so[i] = (int32_t)sch->step_off[i];
        VRT_TRY(sch->step_off_dev.alloc(so.size()));
        VRT_CUDA(cudaMemcpy(sch->step_off_dev.p, so.data(), sizeof(int32_t) * so.size(), cudaMemcpyHostToDevice));
    }
    VRT_CUDA(cudaDeviceSynchronize());
    if (!keep_introspection) {
        sch->cls.release();
        sch->sublevel.release();
        sch->stab.release();
    }
    guard.s = nullptr;
    *out = sch;
    return VRT_OK;
}

DirSchedule* schedule_get(vrt_grid* g, const double k[3], int down, int n_sweeps, double p, int prune, int cv, int* rc) {
    std::lock_guard<std::recursive_mutex> lock(g->mu);
    const OrderCfg oc = order_config(g->n, cv);
    for (auto* s : g->cache)
        if (s->k[0] == k[0] && s->k[1] == k[1] && s->k[2] == k[2] && s->down == down && s->n_sweeps == n_sweeps &&
            s->p == p && s->prune == prune && s->cv == cv && s->order == oc) {
            *rc = VRT_OK;
            return s;
        }
    DirSchedule* s = nullptr;
    *rc = schedule_build(g, k, down, n_sweeps, p, prune, cv, oc, false, &s);
    if (*rc != VRT_OK) return nullptr;
    g->cache.push_back(s);
    return s;
}

void schedule_cache_clear(vrt_grid* g) {
    std::lock_guard<std::recursive_mutex> lock(g->mu);
    for (auto* s : g->cache) delete s;
    g->cache.clear();
}

}  // namespace vrt
