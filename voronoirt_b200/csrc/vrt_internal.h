// vrt_internal.h — internal structures of libvrt.so (not part of the ABI).
//
// Data layout in HBM:
//   * internal cell id c = 0-based rank of the site in perm_up (sites of one BFS layer are contiguous);
//     every per-site array is stored in that order, every nlam x n array as [c][l] with l fastest.
//   * per direction: a list of VISITS sorted by dependent step; a visit is one (cell, sweep) evaluation
//     of irregular_ray_tracing.jl:41-77 with all its operands resolved to (buffer, row) pairs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/vrt.h"

namespace vrt {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define VRT_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return ::vrt::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)
#define VRT_TRY(call)            \
    do {                         \
        int _r = (call);         \
        if (_r != VRT_OK) return _r; \
    } while (0)

// ---------------------------------------------------------------- device buffers
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            return VRT_E_NOMEM;
        }
        n = count;
        return VRT_OK;
    }
    int ensure(size_t count) { return (count <= n && p) ? VRT_OK : alloc(count); }
};

bool is_device_ptr(const void* p);
// copy `bytes` from a host-or-device pointer to device memory (and back)
int copy_in(void* dst_dev, const void* src, size_t bytes, cudaStream_t st = 0);
int copy_out(void* dst, const void* src_dev, size_t bytes, cudaStream_t st = 0);

// ---------------------------------------------------------------- sweep program
// operand selector: top 3 bits of a 32-bit word = buffer, low 29 bits = row
enum : uint32_t { SEL_ZERO = 0u, SEL_MAIN = 1u, SEL_SCR0 = 2u /* SEL_SCR0 + (sweep-1) */ };
constexpr uint32_t SEL_SHIFT = 29;
constexpr uint32_t ROW_MASK = (1u << SEL_SHIFT) - 1u;
constexpr int MAX_SWEEPS = 6;            // scratch selectors 2..6 -> sweeps 1..5 non-final
constexpr int MAX_DIRS = 32;             // directions merged into one sweep launch
constexpr uint32_t DEP_NONE = 0xffffffffu;   // operand needs no producer (boundary value, zero, or never-processed site)
constexpr uint32_t CELL_DUMMY = 0xffffffffu; // padding visit (steps are padded to whole chunks)

// reference classes (SURVEY App. G rule 2)
enum : int32_t { CLS_FINAL = 0, CLS_THIS = 1, CLS_LAG = 2, CLS_ZERO = 3 };

struct __align__(16) Visit {
    uint32_t cell;   // row of S / alpha of the cell itself
    uint32_t dst;    // selector|row the result is written to
    uint32_t u1, u2; // rows of S / alpha of the two upwind cells
    uint32_t src1, src2;  // selector|row the upwind intensities are read from
    uint32_t dep1, dep2;  // chunk (of this direction's program) that produces src1 / src2, or DEP_NONE
    double w1, w2;   // dot_weights (irregular_ray_tracing.jl:51)
    double hr1, hr2; // r/2 (euclidean, :66; the /2 of trapezoidal, functions.jl:393)
};
static_assert(sizeof(Visit) == 64, "Visit must be 64 bytes");

// per-direction stencil in internal order
struct Stencil {
    DevBuf<int32_t> up;     // 2*n internal ids (or -1)
    DevBuf<double> dots;    // 2*n
    DevBuf<double> w;       // 2*n
    DevBuf<double> r;       // 2*n
};

// order of the visits of a wide-row program (schedule.cu, rule 5)
struct OrderCfg {
    int bx = 1, by = 1;     // column blocks in x and y
    int slab = 0;           // levels per slab (0: one slab)
    int step_min = 2048;    // visits per step of the blocked program (steps only interleave the directions in flight)
    bool blocked() const { return bx * by > 1 || slab > 0; }
    bool operator==(const OrderCfg& o) const { return bx == o.bx && by == o.by && slab == o.slab && step_min == o.step_min; }
};

struct DirSchedule {
    double k[3];
    OrderCfg order;
    int down = 0;
    int n_sweeps = 3;
    int prune = 1;
    double p = 7.0;
    int cv = 1;                     // visits per chunk (work unit of one warp; one ready-flag per chunk)
    DevBuf<Visit> visits;           // n_slots records sorted by local step, every step padded to a multiple of cv
    int64_t n_visits = 0;           // real (non-padding) visits
    int64_t n_slots = 0;            // padded length = n_chunks * cv
    int64_t n_chunks = 0;
    std::vector<int64_t> step_off;  // local steps: T_local+1 offsets into visits (padded positions)
    DevBuf<int32_t> step_off_dev;   // the same on the device (int32)
    // (layer, sweep) -> number of sub-levels; index (layer-2)*n_sweeps + (sweep-1)
    std::vector<int32_t> nsub;
    int64_t scr_rows[MAX_SWEEPS] = {0};  // rows needed in scratch buffer s (sweep s+1 non-final writers)
    int64_t n_pushed = 0;           // blocked order: visits whose key was pushed behind a producer of a later block
    // introspection (internal order); released after the build unless asked for (vrt_grid_get_schedule)
    DevBuf<int32_t> cls;       // 2*n
    DevBuf<int32_t> sublevel;  // n (sweep 1)
    DevBuf<int32_t> stab;      // n
};

}  // namespace vrt

// ---------------------------------------------------------------- opaque handles
struct vrt_grid {
    int64_t n = 0, ld = 0, max_nb = 0;
    double bounds[6];
    int device = 0;
    // host copies (ABI order, 1-based)
    std::vector<int64_t> perm_up, perm_down, off_up, off_down;
    int64_t L_up = 0, L_down = 0;
    // device, internal order
    vrt::DevBuf<double> pos;         // 3*n, [c][3] (z,x,y)
    vrt::DevBuf<int32_t> nbr;        // max_nb*n, [j][c]: internal id, or negative wall code, or INT_MIN (empty)
    vrt::DevBuf<int32_t> nnb;        // n
    vrt::DevBuf<int32_t> site_of;    // n: internal id -> 0-based host site
    vrt::DevBuf<int32_t> rank_of;    // n: 0-based host site -> internal id
    vrt::DevBuf<int32_t> layer_up;   // n: 1-based up layer of internal cell
    vrt::DevBuf<int32_t> layer_dn;   // n
    vrt::DevBuf<int32_t> rank_dn;    // n: 0-based rank in perm_down of internal cell
    vrt::DevBuf<int32_t> perm_dn_int;// n: rank in perm_down -> internal id
    // cached schedules (keyed by direction, down, n_sweeps, p, prune, cv)
    std::vector<vrt::DirSchedule*> cache;
    // ready-flags of the dataflow sweep: flag == epoch <=> chunk done in the current launch (no memset per launch)
    vrt::DevBuf<int32_t> flag_pool;
    int32_t epoch = 0;
    // a vrt_grid may be shared by several solvers / host threads: the schedule cache, the flag pool and the epoch are
    // guarded by this mutex (held from the flag carve-out to the end of the sweep launch)
    std::recursive_mutex mu;
    // stream order all sweep launches of this grid go through, and the event pool that times them (sweep.cu)
    cudaStream_t sweep_stream = 0;
    bool sweep_stream_set = false;
    void* sweep_timers = nullptr;
    // regular Cartesian grid (vrt_regular_grid_create): cells in the Julia (nz, nx, ny) order, identity permutation,
    // no layers / stencils / schedules; the formal solver is regular.cu's plane walk
    bool regular = false;
    int64_t rnz = 0, rnx = 0, rny = 0;
    std::vector<double> rz, rx, ry;
    ~vrt_grid();
};

namespace vrt {

// grid.cu
int grid_build(vrt_grid* g, const double* positions, const int64_t* nbr, int64_t ld);
int grid_stencil(vrt_grid* g, const double k[3], double p, Stencil* st);

// schedule.cu
OrderCfg order_config(int64_t n, int cv);   // from the environment (VRT_BLOCKS, VRT_SLAB, VRT_STEP_MIN) or the defaults for n sites
int schedule_build(vrt_grid* g, const double k[3], int down, int n_sweeps, double p, int prune, int cv, const OrderCfg& oc,
                   bool keep_introspection, DirSchedule** out);
// cached per grid (no introspection arrays); serialised by the grid's mutex
DirSchedule* schedule_get(vrt_grid* g, const double k[3], int down, int n_sweeps, double p, int prune, int cv, int* rc);
void schedule_cache_clear(vrt_grid* g);
// visits per chunk for rows of nlam wavelengths (wide rows: one visit per chunk, TMA pipeline)
inline int chunk_visits(int64_t nlam) { return nlam >= 16 ? 1 : (nlam >= 12 ? 4 : (nlam >= 6 ? 8 : (nlam >= 3 ? 16 : 32))); }

// sweep.cu
struct SweepDir {
    const DirSchedule* sch;
    const double* alpha;     // [n][nlam]
    double* I_main;          // [n][nlam]
    double* scratch[MAX_SWEEPS];
};
struct SweepStats {
    double kernels = 0, visits = 0, steps = 0, sweep_ms = 0;
};
// runs the merged sweep program of `nd` directions over nlam wavelengths; S is [n][ldS] (pointer at the chunk's first λ)
// asynchronous launch on `st`; sweep_collect() adds the kernel times of the launches since the last call once the caller has synchronised
int sweep_run(vrt_grid* g, int nd, const SweepDir* dirs, const double* S, int64_t ldS, int64_t nlam, cudaStream_t st, SweepStats* stats);
int sweep_collect(vrt_grid* g, SweepStats* stats);
void sweep_timers_free(void* p);
int sweep_scratch_rows(const DirSchedule* sch, int s);

// comm.cu: in-library NCCL communicators of a solver and the exchange steps of the Λ-iteration
int comm_create(const char* dir_id, int dir_rank, int dir_size, const char* lam_id, int lam_rank, int lam_size, void** out);
void comm_free(void* comm);
cudaStream_t comm_stream(void* comm);
int comm_op(void* comm, double* buf, int64_t count, int op, cudaStream_t st);

// outfile.cu
int outfile_write_at(vrt_outfile* f, const char* name, uint64_t offset, const void* host, size_t nbytes);
int outfile_shape(const vrt_outfile* f, int64_t* n_sites, int64_t* nlam);

// misc kernels (physics.cu)
int permute_rows(const double* src, double* dst, const int32_t* map, int64_t n, int64_t nlam, int gather, cudaStream_t st);
extern thread_local SweepStats g_last_stats;
// regular.cu: one direction of J_λ_regular on DEVICE arrays in the caller's [cell][λ] layout (ld = wavelengths per cell,
// l0 = first wavelength of the chunk): J[.., l0:l0+n_l] (+)= w * I.  I0: [nx*ny][n_l] boundary plane or nullptr (zero).
// have_S[layout] says whether S of this chunk is already laid out (cleared by the caller per chunk / per new S); token is the
// caller's copy of the workspace generation: S is laid out again when another entry point wrote the shared buffers in between.
int regular_dir_accumulate(const vrt_grid* g, const double k[3], int down, int n_sweeps, int64_t n_l, const double* S, int64_t S_ld,
                           int64_t S_l0, const double* alpha, int64_t a_ld, int64_t a_l0, const double* I0, double* J, int64_t J_ld,
                           int64_t J_l0, double w, int accumulate, bool have_S[2], uint64_t* token, SweepStats* st);
// internal layout (0: j = y, 1: j = x) a direction is solved in; callers group their directions by it
int regular_dir_layout(const vrt_grid* g, const double k[3], int* layout);
// wavelengths per chunk that fit next to `extra_vols` more volumes per wavelength held by the caller
int regular_plan_chunk(const vrt_grid* g, int64_t nlam, double extra_vols, int64_t* lc);
// S_new = (1-ε)J + εB with the criterion fused / the criterion alone, one wavelength (solver.cu; used by regular.cu)
int continuum_criterion(int64_t n, const double* S_new, const double* S_old, const double* eps, unsigned long long* diff_bits, int* diff_nan);
int continuum_source_update(int64_t n, const double* B0, const double* eps, const double* J, double* S, unsigned long long* diff_bits,
                            int* diff_nan);

}  // namespace vrt
