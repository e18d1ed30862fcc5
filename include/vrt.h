/*
 * vrt.h — C ABI of libvrt.so, the B200-native (sm_100a) short-characteristics formal solver and
 * Lambda-iteration engine for VoronoiRT's irregular-grid path, plus the rows either side of it that SURVEY.md §8(f)
 * ranks next: the regular-grid comparison solver, the Voronoi neighbour generation, site initialisation and the
 * nearest-site query of the raster resampling.
 *
 * Every entry point replaces one Julia function of the reference (cited as file:line into the
 * reference tree).  The reference has no FFI of its own: its boundary is the set of Julia functions
 * the entry scripts call (compare_searchlight.jl, compare_continuum.jl, compare_line.jl); the Julia
 * host `ccall`s these symbols through julia/VoronoiRTB200.jl.
 *
 * Conventions
 *  - plain C types only; all array arguments are caller-owned.  Every array pointer may be a HOST
 *    pointer or a CUDA DEVICE pointer: the library asks the driver (cudaPointerGetAttributes) and
 *    copies as needed.  Host pointers are never retained after the call returns.
 *  - array layouts are exactly what the Julia arrays hold (column-major): positions 3 x n with rows
 *    (z,x,y); neighbours n x ld with column 0 = count, then 1-based ids, walls -5 (z_min) / -6
 *    (z_max); S, J, alpha, damping nlam x n (wavelength fastest); populations n x 3; R, C 3 x 3 x n.
 *  - ids and permutations are 1-based int64 at the ABI (what Julia holds).
 *  - units are what `ustrip` yields in the reference: m, m^-1, kW m^-2 nm^-1, K, m^-3, m s^-1, nm, s^-1.
 *  - every function returns 0 (VRT_OK) or a negative VRT_E_* code; vrt_last_error() gives the
 *    thread-local message.  No exception crosses the boundary.  There is NO CPU fallback: without a
 *    usable CUDA device every compute entry point returns VRT_E_CUDA.
 */
#ifndef VRT_H
#define VRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRT_ABI_VERSION 4   /* 3: regular-grid entries, native tessellation, trilinear, nearest site; 4: schedule release, in-library
                               collectives, cell-sliced state, output file (additions only) */

enum {
    VRT_OK = 0,
    VRT_E_INVALID = -1,     /* bad argument */
    VRT_E_CUDA = -2,        /* CUDA runtime/driver error, or no device */
    VRT_E_NOMEM = -3,       /* host or device allocation failed */
    VRT_E_GRID = -4,        /* neighbour graph not connected to the wall (the reference would loop forever) */
    VRT_E_STATE = -5,       /* call order / handle state */
    VRT_E_IO = -6           /* file parsing */
};

typedef struct vrt_grid vrt_grid;       /* replaces VoronoiSites' grid part (voronoi_utils.jl:7-28)   */
typedef struct vrt_solver vrt_solver;   /* state of Λ_voronoi (lambda_iteration.jl:207, lambda_continuum.jl:109) */
typedef struct vrt_outfile vrt_outfile; /* the HDF5 output / checkpoint file of create_output_file (io.jl:159-225)          */

/* Two-level + continuum hydrogenic atom: the scalar fields of HydrogenicLine (line.jl:14-72) plus the
 * three broadening constants γ_constant evaluates once per call (broadening.jl:63-82).
 * γ_i = c_unsold*T^0.3*n_HI + gamma_natural + c_linear_stark*n_e^(2/3) + c_quadratic_stark*T^(1/6)*n_e
 * with n_HI, n_e in m^-3 (the host folds Transparency.jl's unit factors into the constants). */
typedef struct vrt_line {
    int64_t nlam;            /* total number of wavelengths, λ = vcat(λbb, λbf_l, λbf_u) (line.jl:59) */
    int64_t lidx[4];         /* line.λidx = [0, nbb, nbb+nbf, nbb+2nbf] (line.jl:61)                  */
    double lambda0;          /* nm                                                                     */
    double Aji, Bji, Bij;    /* s^-1, m^3 J^-1, m^3 J^-1 (line.jl:63-65)                               */
    double chi_i, chi_j, chi_inf;  /* J                                                                */
    int64_t gi, gj, Z;
    double atom_weight;      /* kg */
    double c_unsold;         /* const_unsold(line)           (broadening.jl:24-35)  */
    double gamma_natural;    /* 4.702e8 s^-1                 (broadening.jl:76)     */
    double c_linear_stark;   /* γ_linear_stark = c*n_e[m^-3]^(2/3)  (broadening.jl:77) */
    double c_quadratic_stark;/* const_quadratic_stark(line)  (broadening.jl:52-61)  */
} vrt_line;

/* Per-site inputs, all length n unless noted.  The one-time Transparency.jl quantities (α_cont, C, ε,
 * LTE populations, ΔD) are computed by the host exactly as Λ_voronoi does before its loop
 * (lambda_iteration.jl:216-247) and passed in. */
typedef struct vrt_site_data {
    const double* temperature;        /* K      */
    const double* electron_density;   /* m^-3   */
    const double* hydrogen_density;   /* m^-3, sites.hydrogen_populations (N_H) */
    const double* velocity_z;         /* m/s    */
    const double* velocity_x;
    const double* velocity_y;
    const double* doppler_width;      /* line.ΔD, nm (line.jl:67) */
    const double* alpha_cont;         /* m^-1  (lambda_iteration.jl:223-230) */
    const double* destruction;        /* ελ    (lambda_iteration.jl:244)     */
    const double* C;                  /* 3 x 3 x n collisional rates, s^-1 (rates.jl:52-85) */
    const double* lte_pops;           /* n x 3 LTE populations, m^-3 (populations.jl:112-138) */
} vrt_site_data;

/* Angular quadrature table: the three columns of the quadratures .dat files (functions.jl:33-63), degrees. */
typedef struct vrt_quadrature {
    int64_t n_dirs;
    const double* weights;
    const double* theta;
    const double* phi;
} vrt_quadrature;

typedef struct vrt_config {
    int32_t n_sweeps;        /* in-layer Gauss–Seidel sweeps; the reference hard-wires 3 (lambda_iteration.jl:82) */
    int32_t dir_begin;       /* direction shard [dir_begin, dir_end) of the quadrature table owned by this process; */
    double p;                /* upwind weighting exponent, `const p = 7.0` (irregular_ray_tracing.jl:1) */
    int64_t lam_begin;       /* wavelength shard [lam_begin, lam_end) owned by this process, 0-based;   */
    int64_t lam_end;         /*   0,0 = all wavelengths                                                  */
    int64_t lam_chunk;       /* wavelengths swept together per pass (0 = choose from free HBM)          */
    int32_t prune;           /* 1 = skip re-sweeps of cells whose value cannot change (exact), 0 = visit every cell n_sweeps times */
    int32_t dir_end;         /*   0,0 = all directions (J is then this shard's partial sum unless an all-reduce hook is set) */
    int32_t cell_shard_rank; /* with direction shards: position of this process among the cell_shard_count processes that share */
    int32_t cell_shard_count;/*   its wavelength shard.  When > 1, vrt_lambda_iterate reduce-scatters J over cells (hook op 3), runs the
                              *   source update, rates and statistical equilibrium on its own cell slice only, and all-gathers S and the
                              *   populations (hook op 4).  0 = every process does these for all cells after an all-reduce of J. */
} vrt_config;

/* all-reduce hook of the multi-GPU path: called with a DEVICE buffer of `count` doubles that must be reduced in place.
 *   op 0: sum over the processes that own the same direction shard (i.e. over the wavelength shards): radiative rates;
 *   op 1: max over all processes: convergence criterion;
 *   op 2: sum over the processes that own the same wavelength shard (i.e. over the direction shards): mean intensity J;
 *   op 3: reduce-scatter (sum) over the same group as op 2: the buffer is cell_shard_count equal slices, on return slice
 *         cell_shard_rank holds the sum;   op 4: all-gather over that group: slice r is valid on rank r, all slices on return. */
typedef int (*vrt_allreduce_fn)(void* dev_buf, int64_t count, int32_t op, void* user);

/* per-iteration report handed to the host callback (replaces the println/HDF5 hooks at
 * lambda_iteration.jl:245,280-281,315-317,346) */
typedef struct vrt_iter_info {
    int32_t iteration;       /* 1-based index of the iteration that just finished            */
    int32_t reserved0;
    double diff;             /* criterion value evaluated BEFORE this iteration (lambda_iteration.jl:325-349) */
    double t_opacity_ms, t_sweep_ms, t_source_ms, t_rates_ms, t_stateq_ms, t_total_ms;
    double updates;          /* n * n_dirs * nlam_local */
} vrt_iter_info;
typedef int (*vrt_iter_cb)(const vrt_iter_info* info, void* user);   /* nonzero return stops the loop */

typedef struct vrt_result {
    int32_t iterations;
    int32_t converged;
    double diff;             /* last criterion value */
    double seconds;
} vrt_result;

/* ---------------------------------------------------------------- misc */
int vrt_abi_version(void);
const char* vrt_last_error(void);
int vrt_device_count(int32_t* count);
int vrt_set_device(int32_t device);

/* ---------------------------------------------------------------- grid: read_cell (voronoi_utils.jl:36-85) */

/* Parse the voro++ neighbour text file "id nb1 nb2 ..." (voronoi_utils.jl:42-70; written by
 * rt_preprocessing/output_sites.cc:49).  Call once with nbr == NULL to get ld (= max neighbours + 1),
 * then again with an n x ld int64 buffer (column-major, zero-filled by the library). */
int vrt_read_neighbours(const char* fname, int64_t n, int64_t* nbr, int64_t ld, int64_t* ld_needed);

/* Native Voronoi neighbour generation on the GPU: what the voro++ driver
 * (rt_preprocessing/output_sites.cc:35-49: container periodic in x and y, walls in z, `print_custom("%i %n")`), its text
 * files (io.jl:8-40) and vrt_read_neighbours produce together.  positions: 3 x n rows (z, x, y); bounds: z_min, z_max,
 * x_min, x_max, y_min, y_max.  nbr: n x ld column-major, column 0 = number of faces, then the 1-based ids of the face
 * neighbours, -5 / -6 for the z_min / z_max walls; the SETS equal voro++'s, the order inside a row is not voro++'s (the
 * reference does not define it).  ld_needed (optional) receives max faces + 1; call with nbr == NULL to get it first, or
 * pass ld = 64 (the per-cell capacity) and trim.  VRT_E_GRID when a cell exceeds the capacity. */
int vrt_voronoi_neighbours(int64_t n, const double* positions, const double bounds[6], int64_t* nbr, int64_t ld, int64_t* ld_needed);

/* trilinear (functions.jl:207-248), broadcast over the sites: vals is a (nz, nx, ny) field on the atmosphere axes z, x, y,
 * positions 3 x n rows (z, x, y), out n.  initialise (voronoi_utils.jl:687-708) is six such calls (temperature, electron
 * density, hydrogen density, velocity z, x, y).  Corner search as searchsortedfirst - 1; the lerps in the reference's order
 * with every operation rounded separately (bit-exact against an IEEE evaluation of the Julia expressions).  Sites
 * outside the axes, where the reference throws a BoundsError, get NaN and the call returns VRT_E_INVALID. */
int vrt_trilinear(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                  const double* vals, int64_t n, const double* positions, double* out);

/* rejection_sampling(n_sites, atmos, quantity) (functions.jl:79-121): candidates uniform in the box of the axes, accepted
 * when trilinear(quantity) > U(min quantity, max quantity); positions 3 x n_sites rows (z, x, y).  The candidate stream is
 * Philox4x32-10 with counters (site, trial) and key `seed` — reproducible and independent of the thread mapping, but not
 * Julia's Xoshiro stream: the acceptance rule, and therefore the distribution of the sites, is what is kept.  A candidate
 * exactly on a lower face of the box (where the reference would throw) counts as rejected.  mean_trials (optional)
 * receives the average number of candidates per site. */
int vrt_rejection_sampling(int64_t n_sites, int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                           const double* quantity, uint64_t seed, double* positions, double* mean_trials);

/* Nearest site of each of m points (3 x m rows (z, x, y)): the `nn(KDTree(ustrip.(sites.positions)), p)` of the
 * Voronoi -> raster resampling (voronoi_utils.jl:441-444 and its variants :479-771): plain Euclidean distance, no periodic
 * wrap.  idx: 1-based site ids (ties go to the smaller id), dist (optional): the distances.  The gathers that follow in the
 * reference (temperature[k,i,j] = sites.temperature[idx] ...) stay with the host. */
int vrt_nearest_site(int64_t n, const double* positions, const double bounds[6], int64_t m, const double* points,
                     int64_t* idx, double* dist);

/* The k nearest sites (1 <= k <= 8) of each point, ascending by distance: `knn(tree, p, n_k)` of
 * Voronoi_to_Raster_inv_dist (voronoi_utils.jl:797-805, n_k = 2; inv_dist_itp :848-860 stays with the host).
 * idx, dist: k x m column-major. */
int vrt_nearest_sites(int64_t n, const double* positions, const double bounds[6], int64_t m, const double* points, int32_t k,
                      int64_t* idx, double* dist);

/* initialiseII (voronoi_utils.jl:716-770) for one field: the value at the nearest of the eight corners of the cell of the
 * atmosphere grid that holds the site (corners in the reference's order, first minimum wins).  Arguments as vrt_trilinear. */
int vrt_nearest_corner(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                       const double* vals, int64_t n, const double* positions, double* out);

/* Build the grid: layers from the bottom/top wall (_sort_by_layer_up/_down, voronoi_utils.jl:93-174),
 * stable sort permutations and reduce_layers offsets (:71-79,:253-269), unit Delaunay edge vectors
 * (calc_Delaunay_lines, :186-245).  bounds = {z_min,z_max,x_min,x_max,y_min,y_max}. */
int vrt_grid_create(int64_t n, const double* positions, const int64_t* nbr, int64_t ld,
                    const double bounds[6], vrt_grid** out);
void vrt_grid_destroy(vrt_grid* g);
int vrt_grid_size(const vrt_grid* g, int64_t* n, int64_t* max_nb);
/* number of layers L (offsets has L+1 entries) */
int vrt_grid_num_layers(const vrt_grid* g, int32_t down, int64_t* L);
/* perm (n, 1-based site ids) and offsets (L+1, 1-based, reference semantics: last entry = n, voronoi_utils.jl:266) */
int vrt_grid_get_layers(const vrt_grid* g, int32_t down, int64_t* perm, int64_t* offsets);
/* Delaunay_lines 3 x max_nb x n (wall slots are written as 0; the reference leaves them undefined) */
int vrt_grid_get_delaunay_lines(const vrt_grid* g, double* lines);

/* Upwind stencil for direction k (smallest_angle, voronoi_utils.jl:360-396; weights
 * irregular_ray_tracing.jl:51; path length :66), in host site order.  upwind 2 x n (1-based ids),
 * dots/weights/r 2 x n.  Any output may be NULL. */
int vrt_grid_get_stencil(vrt_grid* g, const double k[3], double p,
                         int64_t* upwind, double* dots, double* weights, double* r);

/* Sweep schedule for direction k: per site, host order:
 *   cls 2 x n: class of each upwind reference (0 FINAL, 1 THIS, 2 LAG, 3 ZERO; -1 for unprocessed sites),
 *   sublevel n: 1-based in-layer dependency level of sweep 1 (0 for boundary / unprocessed sites),
 *   stab n: sweep after which the site's value no longer changes (1..n_sweeps; 0 boundary/unprocessed).
 * n_steps = number of grid-wide dependent steps of the sweep program actually executed. */
int vrt_grid_get_schedule(vrt_grid* g, const double k[3], int32_t down, int32_t n_sweeps, int32_t prune,
                          int32_t* cls, int32_t* sublevel, int32_t* stab, int64_t* n_steps, int64_t* n_visits);

/* Frees the sweep programs the grid caches per (direction, n_sweeps, p, prune) — tens of bytes per cell and direction.
 * Only when no solver created on this grid is alive (their programs are the cached ones); they are rebuilt on demand. */
int vrt_grid_release_schedules(vrt_grid* g);

/* ---------------------------------------------------------------- formal solver */

/* Delaunay_upII (down=0) / Delaunay_downII (down=1) (irregular_ray_tracing.jl:15-82, :96-163), batched
 * over nlam independent wavelengths: S, alpha, I_out are nlam x n; I0 is nlam x n1 in perm order
 * (n1 = offsets[1]-1 boundary sites).  p is the exponent the stale 7-argument call sites pass
 * (compare_searchlight.jl:309,429). */
int vrt_formal_solve(vrt_grid* g, const double k[3], int32_t down, double p, int32_t n_sweeps,
                     int64_t nlam, const double* S, const double* alpha, const double* I0, double* I_out);

/* ---------------------------------------------------------------- regular-grid formal solver */

/* short_characteristics_up (down=0, characteristics.jl:19-95) / short_characteristics_down (down=1, :110-180) with
 * their six ray routines (:191-835), batched over nlam independent wavelengths.  z, x, y: the Atmosphere axes
 * (atmosphere.jl:22-31), x and y INCLUDING the periodic ghost columns, as the reference holds them.  S, alpha,
 * I_out: nlam x nz x nx x ny column-major (wavelength fastest; for nlam = 1 exactly the Julia (nz, nx, ny) array);
 * I0: nlam x nx x ny, the boundary plane (z[0] for up, z[nz-1] for down).  plane_branch (optional, nz int32) receives the
 * ray routine taken per plane: 1 xy, 2 yz, 3 xz, 0 for the boundary plane (argmin at :52-56).  nx, ny <= 1026. */
int vrt_regular_formal_solve(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                             const double k[3], int32_t down, int32_t n_sweeps, int64_t nlam, const double* S,
                             const double* alpha, const double* I0, double* I_out, int32_t* plane_branch);

/* J_λ_regular for a direction-independent opacity (lambda_continuum.jl:1-24; the line form lambda_iteration.jl:23-55
 * differs only in the per-direction alpha): J = Σ_i weights[i] * I_i over the quadrature, rays with θ > 90 solved upwards
 * from I0_up at z[0], rays with θ < 90 downwards from I0_down at z[nz-1] (NULL = zero, as the reference has it); θ = 90
 * belongs to neither branch.  S, alpha, J: nlam x nz x nx x ny; I0_up, I0_down: nlam x nx x ny.  S and alpha are laid out
 * once per internal layout and reused by all directions. */
int vrt_regular_mean_intensity(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                               const vrt_quadrature* quad, int32_t n_sweeps, int64_t nlam, const double* S,
                               const double* alpha, const double* I0_up, const double* I0_down, double* J);

/* Λ_regular, 500 nm continuum (lambda_continuum.jl:58-107) with its criterion (:162-179, maximum over ε > 1e-4 only):
 * S = B_0; while criterion: J = J_λ_regular(S_old); S_new = (1-ε)J + εB_0.  alpha, eps_l, B0: nz x nx x ny as the host
 * computes them at :66-85 (Transparency.jl quantities stay on the host); the bottom boundary is B0[0,:,:] (:16).
 * S_out, J_out (optional): nz x nx x ny.  cb as in vrt_lambda_iterate. */
int vrt_regular_lambda_iterate(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                               const vrt_quadrature* quad, int32_t n_sweeps, const double* alpha, const double* eps_l,
                               const double* B0, double eps, int32_t maxiter, vrt_iter_cb cb, void* user, double* S_out,
                               double* J_out, vrt_result* out);

/* A vrt_grid over the regular Cartesian atmosphere (atmosphere.jl:22-31): cell c = iz + nz*(ix + nx*iy), i.e. the memory
 * order of the reference's (nz, nx, ny) arrays, x and y with their ghost columns.  The handle feeds the same Λ-iteration
 * engine as a Voronoi grid — vrt_solver_create_line / _continuum, vrt_mean_intensity (J_λ_regular, lambda_iteration.jl:1-58),
 * vrt_calculate_R (rates.jl:96-143), vrt_lambda_iterate (Λ_regular, lambda_iteration.jl:116-205 with criterion :299-323),
 * vrt_get_state / vrt_set_state — with every per-site array being the flattened (nz, nx, ny) array, S and J
 * nlam x nz x nx x ny, populations nz x nx x ny x 3.  The formal solutions run through the plane walk of
 * vrt_regular_formal_solve.  Layer / stencil / schedule queries and cell shards return VRT_E_STATE /
 * VRT_E_INVALID on such a handle; direction and wavelength shards work as on a Voronoi grid (J is all-reduced over the
 * direction group).  Destroy with vrt_grid_destroy. */
int vrt_regular_grid_create(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, vrt_grid** out);

/* the regular-grid entries keep their device workspace (3-6 internal copies of one wavelength chunk) between calls;
 * this frees it, e.g. before handing the GPU to an irregular-grid solver. */
int vrt_regular_release_workspace(void);

/* ---------------------------------------------------------------- Λ-iteration engine */

/* NLTE line solver state (Λ_voronoi, lambda_iteration.jl:207-297). lambda: nlam wavelengths, nm. */
int vrt_solver_create_line(vrt_grid* g, const vrt_line* line, const double* lambda,
                           const vrt_site_data* sites, const vrt_quadrature* quad,
                           const vrt_config* cfg, vrt_solver** out);
/* 500 nm continuum solver state (Λ_voronoi, lambda_continuum.jl:109-160): alpha_cont, eps (ε_λ), B0 length n. */
int vrt_solver_create_continuum(vrt_grid* g, const double* alpha_cont, const double* eps, const double* B0,
                                const vrt_quadrature* quad, const vrt_config* cfg, vrt_solver** out);
void vrt_solver_destroy(vrt_solver* s);
int vrt_solver_set_allreduce(vrt_solver* s, vrt_allreduce_fn fn, void* user);
/* Collectives inside the library (preferred over the hook above; north star (5): "NCCL allreduce over NVLink").  One process
 * per GPU.  The host creates a unique id with vrt_nccl_unique_id on ONE process of a group, ferries its 128 bytes to the
 * others with whatever it has (MPI, sockets, a file, torch.distributed — no NCCL binding needed on the host side) and every
 * member calls vrt_solver_comm_init; the library then does ncclCommInitRank itself and issues the reduce-scatter of J, the
 * all-gather of S and the populations, the all-reduce of the rates and the max of the criterion on its own stream.
 *   dir group: the processes that share a wavelength shard and differ in direction shard (ops 2, 3, 4 of the hook);
 *              dir_rank / dir_size must equal vrt_config.cell_shard_rank / cell_shard_count when cell shards are used;
 *   lam group: the processes that share a direction shard and differ in wavelength shard (op 0).
 * A group of size <= 1 takes a NULL id.  NCCL is bound at run time (libnccl.so.2, or $VRT_NCCL_LIB); vrt_nccl_available() is 0
 * when it cannot be loaded, and these calls then return VRT_E_STATE.  The independence being exploited is
 * lambda_iteration.jl:84-110 (every direction and wavelength is a separate formal solution). */
int vrt_nccl_available(void);
int vrt_nccl_version(int32_t* version);
int vrt_nccl_unique_id(char id[128]);
int vrt_solver_comm_init(vrt_solver* s, const char* dir_id, int32_t dir_rank, int32_t dir_size,
                         const char* lam_id, int32_t lam_rank, int32_t lam_size);

/* Work of each direction this solver holds, as the number of (cell, sweep) visits of its sweep program, in the order of the
 * solver's quadrature table without the θ = 90 rows.  visits may be NULL to query n_dirs.  A host that shards the directions
 * over processes balances them with it (longest processing time first) instead of dealing them round-robin. */
int vrt_solver_direction_visits(const vrt_solver* s, int64_t* n_dirs, double* visits, int64_t capacity);

/* Reduction of J through peer memory instead of a reduce-scatter (one node, NVLink / NVSwitch).  Every process of the direction
 * group exports its J buffer (vrt_solver_peer_handle: a 64-byte CUDA IPC handle), the host gathers the handles of the group in
 * rank order (dir_size x 64 bytes) and hands them to every member (vrt_solver_peer_attach).  From then on vrt_lambda_iterate
 * replaces "reduce-scatter J, then update S on the own cell slice" by ONE kernel that reads the slice from all peers' buffers
 * over NVLink, adds them in rank order, keeps the sum as its J and updates S.  Needs vrt_solver_comm_init (the barrier in front
 * of the kernel is a one-element all-reduce) and cell shards.  If the handles cannot be opened the call fails and the solver keeps
 * using the reduce-scatter. */
int vrt_solver_peer_handle(vrt_solver* s, char handle[64]);
int vrt_solver_peer_attach(vrt_solver* s, const char* handles, int32_t count);
/* Unmaps the peers' buffers (the solver goes back to the reduce-scatter).  CUDA requires that every importer has closed a handle
 * before the exporter frees the memory: call this on every process, then synchronise the processes, then destroy the solvers. */
int vrt_solver_peer_detach(vrt_solver* s);

/* Restricts one direction of this solver (index into its quadrature table without the θ = 90 rows) to the local wavelengths
 * [lam_begin, lam_end): the direction is then shared with another process that takes the remaining wavelengths (both add
 * their part into J before the reduction over the direction group).  This is how 20 directions balance on 8 processes: two
 * whole directions each plus one half of a ninth.  lam_begin == lam_end == 0 or the full range restores the default.  Line
 * solver on a Voronoi grid only; a sub-range needs at least 16 wavelengths (it runs the wide-row sweep program). */
int vrt_solver_set_direction_lambda(vrt_solver* s, int32_t direction, int64_t lam_begin, int64_t lam_end);

/* (Re)upload one per-site input of the line solver.  In the reference these are plain function arguments
 * (α_cont of J_λ_voronoi, LTE_pops of calculate_R, C of get_revised_populations), so a drop-in caller may
 * hand them over late or change them between calls.  Shapes as in vrt_site_data. */
enum { VRT_FIELD_ALPHA_CONT = 0, VRT_FIELD_DESTRUCTION = 1, VRT_FIELD_C = 2, VRT_FIELD_LTE_POPS = 3 };
int vrt_solver_set_field(vrt_solver* s, int32_t field, const double* data);
/* local wavelength count of this shard */
int vrt_solver_nlam_local(const vrt_solver* s, int64_t* nlam_local);

/* J_λ_voronoi (lambda_iteration.jl:60-113 / lambda_continuum.jl:27-56).  S, J: nlam_local x n;
 * populations n x 3 (NULL for the continuum solver); damping (optional) nlam_local x n. */
int vrt_mean_intensity(vrt_solver* s, const double* S, const double* populations, double* J, double* damping);

/* calculate_R (rates.jl:154-201): J nlam_local x n -> R 3 x 3 x n.  damping NULL = use the γ of the last
 * vrt_mean_intensity call.  With a wavelength shard the result is this shard's partial sum unless an
 * all-reduce hook is set. */
int vrt_calculate_R(vrt_solver* s, const double* J, const double* damping, double* R);

/* get_revised_populations (populations.jl:191-221): R, C 3 x 3 x n, N_H n -> populations n x 3. */
int vrt_get_revised_populations(int64_t n, const double* R, const double* C, const double* N_H, double* populations);

/* Λ_voronoi loop (lambda_iteration.jl:253-285 / lambda_continuum.jl:145-149).  Starts from S = B_0 and
 * LTE populations unless vrt_set_state was called. */
int vrt_lambda_iterate(vrt_solver* s, double eps, int32_t maxiter, vrt_iter_cb cb, void* user, vrt_result* out);

/* checkpoint hooks (replace write_to_file / recover_simulation.jl): any pointer may be NULL.
 * S, J nlam_local x n; populations n x 3. */
int vrt_get_state(vrt_solver* s, double* S, double* J, double* populations);
int vrt_set_state(vrt_solver* s, const double* S, const double* populations);

/* Cell-sliced checkpoint hooks for the multi-GPU solve.  With cell shards every process owns the source function, the mean
 * intensity and the populations of the cells [first, last) in the library's INTERNAL cell order — internal cell c is site
 * perm_up[c] (vrt_grid_get_layers, down = 0) — and only that slice crosses its PCIe link: S, J are nlam_local x (last-first),
 * populations (last-first) x 3.  vrt_set_state_slice is a collective call: the other processes' slices of S and of the
 * populations arrive through the all-gather over NVLink.  Without cell shards the slice is all cells (in internal order). */
int vrt_solver_cell_slice(const vrt_solver* s, int64_t* first, int64_t* last);
int vrt_get_state_slice(vrt_solver* s, double* S, double* J, double* populations);
int vrt_set_state_slice(vrt_solver* s, const double* S, const double* populations);

/* ---------------------------------------------------------------- output / checkpoint file
 * create_output_file + write_to_file (io.jl:57-225) without the HDF5 library: a version-0-superblock HDF5 file with one flat
 * root group of contiguous little-endian datasets carrying the reference's names and shapes, so that
 * recover_simulation.jl:213-277 and the scripts under python/ read it like a file written by HDF5.jl.  Dataset names (Voronoi,
 * io.jl:196-225): source_function (nλ, n_sites), populations (n_sites, 3), positions (3, n_sites), temperature,
 * hydrogen_populations, electron_density, velocity_z, velocity_x, velocity_y (n_sites), boundaries (6), convergence
 * (maxiter+1, created as zeros), n_bb, n_bf (1, Int64), wavelength (nλ), line_center (1), time (1); the regular-grid form
 * (io.jl:159-190) has z, x, y and (nz, nx, ny)-shaped fields instead.  Shapes are Julia shapes: the file stores them with the
 * dimensions reversed over the same bytes, as HDF5.jl does.  All values Float64 unless noted. */
int vrt_output_create(const char* path, int64_t nlam, int64_t n_sites, int64_t maxiter, vrt_outfile** out);
int vrt_output_create_regular(const char* path, int64_t nlam, int64_t nz, int64_t nx, int64_t ny, int64_t maxiter, vrt_outfile** out);
int vrt_output_dataset_size(const vrt_outfile* f, const char* name, int64_t* nbytes);
/* write_to_file(array, output_path): the whole dataset; data is a host or device pointer of exactly the dataset's size */
int vrt_output_write(vrt_outfile* f, const char* name, const void* data, int64_t nbytes);
/* write_to_file(difference, iteration, output_path) (io.jl:129-135): convergence[iteration] = difference, 1-based */
int vrt_output_write_convergence(vrt_outfile* f, int64_t iteration, double difference);
/* write_to_file(S_λ, …) and write_to_file(populations, …) straight from the solver's device state (what Λ_voronoi does after
 * every iteration, lambda_iteration.jl:280-281); call it from the vrt_iter_cb of vrt_lambda_iterate */
int vrt_output_write_state(vrt_outfile* f, vrt_solver* s);
int vrt_output_close(vrt_outfile* f);

/* Fingerprint of the device-resident state, reduced on the device (nothing the size of S crosses the bus):
 * out[0] = Σ S, out[1] = max |S|, out[2] = Σ populations (0 for the continuum solver), out[3] = Σ J over the cells this
 * process owns (all cells unless the post-J stages are cell-sharded).  Runs of the same problem on 1, 2, 4, 8 GPUs must agree
 * in out[0..2] to rounding (the shards add the directions in another order). */
int vrt_state_checksum(vrt_solver* s, double out[4]);

/* counters of the last vrt_mean_intensity / vrt_formal_solve call on this thread's device:
 * out[0] kernels launched, out[1] cell visits, out[2] dependent steps, out[3] sweep-kernel ms (CUDA events) */
int vrt_last_stats(double out[8]);

#ifdef __cplusplus
}
#endif
#endif /* VRT_H */
