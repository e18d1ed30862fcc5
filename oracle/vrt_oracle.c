/*
 * vrt_oracle.c — CPU restatement of VoronoiRT's irregular-grid formal solver and Λ-iteration.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the timed CPU baseline
 * (bench.py cpu_baseline / --impl reference).  Nothing under voronoirt_b200/ may import, link or call
 * it; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
 *
 * PARITY UNPINNED: the reference is Julia and no Julia toolchain exists in this image, and the
 * reference ships no test or golden vector for the irregular path.  What pins this file:
 *   - the bound-bound wavelength table printed in python/plot_line.py:16-34 (tests/test_oracle_golden.py),
 *   - the searchlight beam centroid of data/searchlight_data/I_160_45_voronoi.npy (statistic only),
 *   - an independent pure-Python restatement (tests/pyref.py) and the values of SURVEY.md App. F.
 * Third-party arithmetic (Transparency.jl, unpinned, un-vendored): voigt_profile is restated from the
 * published Humlíček (1982, JQSRT 27, 437) w4 algorithm that Transparency.jl implements.
 * Standard-library arithmetic: `norm(p_d)` (voronoi_utils.jl:238) is LinearAlgebra.generic_norm2, which takes its unscaled
 * branch whenever n·max|x|² is finite and non-zero — sqrt of the left-to-right sum of squares, exactly what orc_calc_delaunay_lines
 * computes; `inv(A)*b` (populations.jl:214) goes through LAPACK getrf/getri of whatever BLAS Julia ships and is restated as an LU
 * solve (last-bit differences of the 2 x 2 solve are inside the 1e-6 population tolerance and cannot be pinned from here).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 * All arrays use the Julia (column-major, 1-based ids) layouts.  Float64 throughout; compile with
 * -ffp-contract=off so that comparisons of dot products are reproducible bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/vrt.h"

/* CODATA 2018 (PhysicalConstants.CODATA2018, src/atmosphere.jl:1-8) */
#define H_PLANCK 6.62607015e-34
#define K_B 1.380649e-23
#define C_0 299792458.0
#define E_CHARGE 1.602176634e-19
#define M_ELECTRON 9.1093837015e-31
#define EPS_0 8.8541878128e-12
#define R_INF 10973731.568160
#define PI 3.14159265358979323846

typedef struct orc_sites {
    int64_t n, ld, max_nb;
    const double* pos;      /* 3 x n (z,x,y) */
    const int64_t* nbr;     /* n x ld, col 0 = count */
    double* lines;          /* 3 x max_nb x n */
    int64_t* layers_up;     /* L_up+1 */
    int64_t* layers_down;
    int64_t L_up, L_down;
    int64_t* perm_up;       /* n, 1-based */
    int64_t* perm_down;
    double x_min, x_max, y_min, y_max, z_min, z_max;
} orc_sites;

#define NBR(s, i, j) ((s)->nbr[(i) + (s)->n * (j)])  /* 0-based row i, column j */

/* ------------------------------------------------------------------ read_cell: parsing
 * src/voronoi_utils.jl:42-70.  Returns the number of columns needed (max neighbours + 1) or <0. */
int64_t orc_read_neighbours(const char* fname, int64_t n, int64_t* nbr, int64_t ld) {
    FILE* f = fopen(fname, "r");
    if (!f) return -1;
    size_t cap = 1 << 16;
    char* buf = (char*)malloc(cap);
    int64_t maxn = 0;
    if (nbr) memset(nbr, 0, sizeof(int64_t) * (size_t)n * (size_t)ld);
    while (fgets(buf, (int)cap, f)) {
        char* p = buf;
        char* e;
        long long id = strtoll(p, &e, 10);
        if (e == p) continue;
        p = e;
        int64_t cnt = 0;
        for (;;) {
            long long v = strtoll(p, &e, 10);
            if (e == p) break;
            p = e;
            cnt++;
            if (nbr && id >= 1 && id <= n && cnt < ld) nbr[(id - 1) + n * cnt] = v;
        }
        if (nbr && id >= 1 && id <= n) nbr[(id - 1)] = cnt;
        if (cnt > maxn) maxn = cnt;
    }
    free(buf);
    fclose(f);
    return maxn + 1;
}

/* ------------------------------------------------------------------ _sort_by_layer_up/_down
 * src/voronoi_utils.jl:93-130 (wall = -5) and :138-174 (wall = -6).  layers: n, 1-based layer numbers.
 * Returns 0, or -1 when a pass makes no progress (the reference would spin forever). */
int orc_sort_by_layer(const int64_t* nbr, int64_t n, int64_t wall, int64_t* layers) {
    for (int64_t i = 0; i < n; i++) layers[i] = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t cnt = nbr[i];
        for (int64_t j = 1; j <= cnt; j++)
            if (nbr[i + n * j] == wall) layers[i] = 1;
    }
    int64_t lower = 1;
    for (;;) {
        int64_t assigned = 0;
        /* a pass tests `== lower` and writes `lower + 1` only (voronoi_utils.jl:114): its result does not depend on the
         * order of the sites, so the scan may run in parallel (set-up of the 16 M-site baseline would take minutes otherwise) */
#pragma omp parallel for reduction(+ : assigned) schedule(static)
        for (int64_t i = 0; i < n; i++) {
            if (layers[i] == 0) {
                int64_t cnt = nbr[i];
                for (int64_t j = 1; j <= cnt; j++) {
                    int64_t nb = nbr[i + n * j];
                    if (nb > 0 && layers[nb - 1] == lower) {
                        layers[i] = lower + 1;
                        assigned++;
                        break;
                    }
                }
            }
        }
        int any0 = 0;
#pragma omp parallel for reduction(| : any0) schedule(static)
        for (int64_t i = 0; i < n; i++)
            if (layers[i] == 0) any0 |= 1;
        if (!any0) break;
        if (assigned == 0) return -1;
        lower++;
    }
    return 0;
}

/* sortperm (stable; Julia guarantees stability for sortperm's default algorithm), src/voronoi_utils.jl:72 */
static void orc_sortperm(const int64_t* layers, int64_t n, int64_t* perm, int64_t* sorted) {
    int64_t maxl = 0;
    for (int64_t i = 0; i < n; i++)
        if (layers[i] > maxl) maxl = layers[i];
    int64_t* cnt = (int64_t*)calloc((size_t)maxl + 2, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) cnt[layers[i] + 1]++;
    for (int64_t l = 1; l <= maxl + 1; l++) cnt[l] += cnt[l - 1];
    for (int64_t i = 0; i < n; i++) {
        int64_t r = cnt[layers[i]]++;
        perm[r] = i + 1;
        sorted[r] = layers[i];
    }
    free(cnt);
}

/* reduce_layers, src/voronoi_utils.jl:253-269.  out has max(layers)+1 entries; returns that count. */
static int64_t orc_reduce_layers(const int64_t* sorted, int64_t n, int64_t* out) {
    int64_t maxl = 0;
    for (int64_t i = 0; i < n; i++)
        if (sorted[i] > maxl) maxl = sorted[i];
    out[0] = 1;
    int64_t layer = 2;
    for (int64_t i = 0; i < n; i++) {
        if (sorted[i] == layer) {
            out[layer - 1] = i + 1;
            layer++;
        }
    }
    out[maxl] = n; /* reduced_layers[end] = length(layers): n, not n+1 (Q1) */
    return maxl + 1;
}

/* ------------------------------------------------------------------ calc_Delaunay_lines
 * src/voronoi_utils.jl:186-245, including the mirror quirk at :221-222/:231-232 (Q3). */
static void orc_calc_delaunay_lines(orc_sites* s) {
    int64_t n = s->n, mnb = s->max_nb;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double* P = s->pos + 3 * i;
        double x_r_r = s->x_max - P[1];
        double x_r_l = P[1] - s->x_min;
        double y_r_r = s->y_max - P[2];
        double y_r_l = P[2] - s->y_min;
        int64_t cnt = NBR(s, i, 0);
        for (int64_t j = 0; j < mnb; j++) {
            double* out = s->lines + 3 * (j + mnb * i);
            out[0] = out[1] = out[2] = 0.0;
            if (j >= cnt) continue;
            int64_t nb = NBR(s, i, j + 1);
            if (nb <= 0) continue;
            double pn0 = s->pos[3 * (nb - 1)], pn1 = s->pos[3 * (nb - 1) + 1], pn2 = s->pos[3 * (nb - 1) + 2];
            double x_i_r = fabs(s->x_max - pn1);
            double x_i_l = fabs(pn1 - s->x_min);
            if (x_r_r + x_i_l < P[1] - pn1)
                pn1 = s->x_max + pn1 - s->x_min;
            else if (x_r_l + x_i_r < pn1 - P[1])
                pn1 = s->x_min + s->x_max - pn1;
            double y_i_r = fabs(s->y_max - pn2);
            double y_i_l = fabs(pn2 - s->y_min);
            if (y_r_r + y_i_l < P[2] - pn2)
                pn2 = s->y_max + pn2 - s->y_min;
            else if (y_r_l + y_i_r < pn2 - P[2])
                pn2 = s->y_min + s->y_max - pn2;
            double d0 = pn0 - P[0], d1 = pn1 - P[1], d2 = pn2 - P[2];
            double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            out[0] = d0 / nrm;
            out[1] = d1 / nrm;
            out[2] = d2 / nrm;
        }
    }
}

/* ------------------------------------------------------------------ read_cell (grid part)
 * src/voronoi_utils.jl:65-84.  bounds = {z_min,z_max,x_min,x_max,y_min,y_max}. */
orc_sites* orc_sites_create(int64_t n, const double* pos, const int64_t* nbr, int64_t ld, const double* bounds) {
    orc_sites* s = (orc_sites*)calloc(1, sizeof(orc_sites));
    s->n = n;
    s->ld = ld;
    s->pos = pos;
    s->nbr = nbr;
    s->z_min = bounds[0]; s->z_max = bounds[1];
    s->x_min = bounds[2]; s->x_max = bounds[3];
    s->y_min = bounds[4]; s->y_max = bounds[5];
    int64_t mnb = 0;
    for (int64_t i = 0; i < n; i++)
        if (nbr[i] > mnb) mnb = nbr[i];
    s->max_nb = mnb;
    int64_t* layers = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    int64_t* sorted = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    s->perm_up = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    s->perm_down = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    s->layers_up = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 2));
    s->layers_down = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 2));
    int bad = 0;
    if (orc_sort_by_layer(nbr, n, -5, layers)) bad = 1;
    if (!bad) {
        orc_sortperm(layers, n, s->perm_up, sorted);
        s->L_up = orc_reduce_layers(sorted, n, s->layers_up) - 1;
    }
    if (!bad && orc_sort_by_layer(nbr, n, -6, layers)) bad = 1;
    if (!bad) {
        orc_sortperm(layers, n, s->perm_down, sorted);
        s->L_down = orc_reduce_layers(sorted, n, s->layers_down) - 1;
    }
    free(layers);
    free(sorted);
    if (bad) {
        free(s->perm_up); free(s->perm_down); free(s->layers_up); free(s->layers_down); free(s);
        return NULL;
    }
    s->lines = (double*)malloc(sizeof(double) * 3 * (size_t)(mnb > 0 ? mnb : 1) * (size_t)n);
    orc_calc_delaunay_lines(s);
    return s;
}

void orc_sites_destroy(orc_sites* s) {
    if (!s) return;
    free(s->lines); free(s->perm_up); free(s->perm_down); free(s->layers_up); free(s->layers_down);
    free(s);
}

int64_t orc_sites_num_layers(const orc_sites* s, int down) { return down ? s->L_down : s->L_up; }
int64_t orc_sites_max_nb(const orc_sites* s) { return s->max_nb; }
void orc_sites_get_layers(const orc_sites* s, int down, int64_t* perm, int64_t* offsets) {
    const int64_t* p = down ? s->perm_down : s->perm_up;
    const int64_t* o = down ? s->layers_down : s->layers_up;
    int64_t L = down ? s->L_down : s->L_up;
    memcpy(perm, p, sizeof(int64_t) * (size_t)s->n);
    memcpy(offsets, o, sizeof(int64_t) * (size_t)(L + 1));
}
void orc_sites_get_lines(const orc_sites* s, double* lines) {
    memcpy(lines, s->lines, sizeof(double) * 3 * (size_t)s->max_nb * (size_t)s->n);
}

/* ------------------------------------------------------------------ smallest_angle(n::Int, ...)
 * src/voronoi_utils.jl:360-396.  idx0: 0-based site.  Order-dependent "top-2" (Q2). */
static inline void orc_smallest_angle(const orc_sites* s, int64_t idx0, const double* k, double* dots, int64_t* ind) {
    dots[0] = -1.0;
    dots[1] = -1.0;
    ind[0] = 0;
    ind[1] = 0;
    int64_t cnt = NBR(s, idx0, 0);
    for (int64_t j = 0; j < cnt; j++) {
        int64_t nb = NBR(s, idx0, j + 1);
        if (nb > 0) {
            const double* l = s->lines + 3 * (j + s->max_nb * idx0);
            double d = k[0] * l[0] + k[1] * l[1] + k[2] * l[2];
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
}

void orc_stencil(const orc_sites* s, const double* k, double p, int64_t* upwind, double* dots, double* weights, double* r) {
    for (int64_t i = 0; i < s->n; i++) {
        double d[2];
        int64_t ind[2];
        orc_smallest_angle(s, i, k, d, ind);
        double p1 = pow(d[0], p), p2 = pow(d[1], p);
        double sum = p1 + p2;
        for (int m = 0; m < 2; m++) {
            if (upwind) upwind[2 * i + m] = ind[m];
            if (dots) dots[2 * i + m] = d[m];
            if (weights) weights[2 * i + m] = (m == 0 ? p1 : p2) / sum;
            if (r) {
                if (ind[m] > 0) {
                    const double* a = s->pos + 3 * i;
                    const double* b = s->pos + 3 * (ind[m] - 1);
                    double e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
                    r[2 * i + m] = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
                } else
                    r[2 * i + m] = 0.0;
            }
        }
    }
}

/* linear_weights, src/functions.jl:484-500 (Q7) */
static inline void orc_linear_weights(double dtau, double* a, double* b, double* e) {
    if (dtau < 5e-4) {
        *e = 1 - dtau + 0.5 * (dtau * dtau);
        *a = dtau * (1.0 / 2 - dtau / 3);
        *b = dtau * (1.0 / 2 - dtau / 6);
    } else if (dtau > 50) {
        *e = 0.0;
        *a = 1 / dtau;
        *b = 1.0 - *a;
    } else {
        *e = exp(-dtau);
        *a = (1 - *e) / dtau - *e;
        *b = 1 - *a - *e;
    }
}

/* ------------------------------------------------------------------ Delaunay_upII / Delaunay_downII
 * src/irregular_ray_tracing.jl:15-82 and :96-163.  One wavelength: S, alpha, I (n), I_0 (n1).
 * hoist == 0: smallest_angle is recomputed at every visit as the reference does (:50);
 * hoist != 0: `st_*` hold the per-direction stencil (identical numbers, just not recomputed). */
void orc_delaunay(const orc_sites* s, int down, const double* k, const double* S, const double* I_0,
                  const double* alpha, int n_sweeps, double p, double* I,
                  int hoist, const int64_t* st_up, const double* st_w, const double* st_r) {
    int64_t n = s->n;
    const int64_t* perm = down ? s->perm_down : s->perm_up;
    const int64_t* lay = down ? s->layers_down : s->layers_up;
    int64_t max_layer = (down ? s->L_down : s->L_up) + 1; /* length(sites.layers_up) */
    for (int64_t i = 0; i < n; i++) I[i] = 0.0;
    int64_t lower_idx = lay[1] - 1;
    for (int64_t i = 0; i < lower_idx; i++) I[perm[i] - 1] = I_0[i];
    for (int64_t layer = 2; layer <= max_layer - 1; layer++) {
        int64_t lo = lay[layer - 1];
        int64_t up = lay[layer];
        for (int sweep = 1; sweep <= n_sweeps; sweep++) {
            int64_t cntc = up - lo; /* i in lo : up-1 */
            for (int64_t t = 0; t < cntc; t++) {
                int64_t i = down ? (up - 1 - t) : (lo + t);
                int64_t idx = perm[i - 1] - 1;
                int64_t ind[2];
                double w[2], rr[2];
                if (!hoist) {
                    double d[2];
                    orc_smallest_angle(s, idx, k, d, ind);
                    double p1 = pow(d[0], p), p2 = pow(d[1], p);
                    double sum = p1 + p2;
                    w[0] = p1 / sum;
                    w[1] = p2 / sum;
                    for (int m = 0; m < 2; m++) {
                        const double* a = s->pos + 3 * idx;
                        const double* b = s->pos + 3 * (ind[m] - 1);
                        double e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
                        rr[m] = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
                    }
                } else {
                    ind[0] = st_up[2 * idx]; ind[1] = st_up[2 * idx + 1];
                    w[0] = st_w[2 * idx]; w[1] = st_w[2 * idx + 1];
                    rr[0] = st_r[2 * idx]; rr[1] = st_r[2 * idx + 1];
                }
                I[idx] = 0.0;
                for (int m = 0; m < 2; m++) {
                    int64_t u = ind[m] - 1;
                    double dtau = rr[m] * (alpha[idx] + alpha[u]) / 2; /* trapezoidal, functions.jl:392-395 */
                    double a, b, e;
                    orc_linear_weights(dtau, &a, &b, &e);
                    I[idx] += (e * I[u] + a * S[u] + b * S[idx]) * w[m];
                }
            }
        }
    }
}

/* Batched over wavelengths with the reference's decomposition (threads over λ, lambda_iteration.jl:91).
 * S, alpha, I: nlam x n (λ fastest); I0: nlam x n1.  Row copies mirror `S_λ[l,:]`. */
void orc_formal_solve(const orc_sites* s, const double* k, int down, double p, int n_sweeps, int64_t nlam,
                      const double* S, const double* alpha, const double* I0, double* I_out, int hoist) {
    int64_t n = s->n;
    int64_t n1 = (down ? s->layers_down : s->layers_up)[1] - 1;
    int64_t* st_up = NULL;
    double *st_w = NULL, *st_r = NULL;
    if (hoist) {
        st_up = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)n);
        st_w = (double*)malloc(sizeof(double) * 2 * (size_t)n);
        st_r = (double*)malloc(sizeof(double) * 2 * (size_t)n);
        orc_stencil(s, k, p, st_up, NULL, st_w, st_r);
    }
#pragma omp parallel
    {
        double* Sv = (double*)malloc(sizeof(double) * (size_t)n);
        double* av = (double*)malloc(sizeof(double) * (size_t)n);
        double* Iv = (double*)malloc(sizeof(double) * (size_t)n);
        double* I0v = (double*)malloc(sizeof(double) * (size_t)(n1 > 0 ? n1 : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t l = 0; l < nlam; l++) {
            for (int64_t i = 0; i < n; i++) { Sv[i] = S[l + nlam * i]; av[i] = alpha[l + nlam * i]; }
            for (int64_t i = 0; i < n1; i++) I0v[i] = I0[l + nlam * i];
            orc_delaunay(s, down, k, Sv, I0v, av, n_sweeps, p, Iv, hoist, st_up, st_w, st_r);
            for (int64_t i = 0; i < n; i++) I_out[l + nlam * i] = Iv[i];
        }
        free(Sv); free(av); free(Iv); free(I0v);
    }
    free(st_up); free(st_w); free(st_r);
}

/* ------------------------------------------------------------------ Voigt profile
 * Transparency.jl `voigt_profile(a, v, ΔD) = Re(humlicek(a, v)) / (sqrt(π) ΔD)` (call sites
 * src/line.jl:133, src/rates.jl:408).  Humlíček (1982) w4, four regions; complex arithmetic written
 * out so that the CUDA kernel can repeat it operation for operation. */
typedef struct { double re, im; } cplx;
static inline cplx c_mul(cplx a, cplx b) { cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cplx c_add_r(double x, cplx a) { cplx r = {x + a.re, a.im}; return r; }
static inline cplx c_rsub(double x, cplx a) { cplx r = {x - a.re, -a.im}; return r; }
static inline cplx c_div(cplx a, cplx b) {
    double den = b.re * b.re + b.im * b.im;
    cplx r = {(a.re * b.re + a.im * b.im) / den, (a.im * b.re - a.re * b.im) / den};
    return r;
}
static inline cplx c_scale(double x, cplx a) { cplx r = {x * a.re, x * a.im}; return r; }

#define INV_SQRT_PI 0.5641895835477563

double orc_humlicek_re(double a, double v) {
    cplx z = {v, a};
    double s = fabs(v) + a;
    cplx w;
    if (s > 15.0) { /* region I: w = i/√π z / (z² - 0.5) */
        cplx zz = c_mul(z, z);
        cplx den = {zz.re - 0.5, zz.im};
        cplx num = {-INV_SQRT_PI * z.im, INV_SQRT_PI * z.re};
        w = c_div(num, den);
    } else if (s > 5.5) { /* region II: w = i z (z²/√π - 1.4104739589) / (0.75 + z²(z² - 3)) */
        cplx zz = c_mul(z, z);
        cplx t1 = {zz.re * INV_SQRT_PI - 1.4104739589, zz.im * INV_SQRT_PI};
        cplx zt = c_mul(z, t1);
        cplx num = {-zt.im, zt.re};
        cplx zz3 = {zz.re - 3.0, zz.im};
        cplx den = c_add_r(0.75, c_mul(zz, zz3));
        w = c_div(num, den);
    } else {
        double x = v, y = a;
        cplx t = {y, -x};
        if (y >= 0.195 * fabs(x) - 0.176) { /* region III */
            cplx num = c_add_r(3.778987, c_scale(0.5642236, t));
            num = c_add_r(11.96482, c_mul(t, num));
            num = c_add_r(20.20933, c_mul(t, num));
            num = c_add_r(16.4955, c_mul(t, num));
            cplx den = c_add_r(6.699398, t);
            den = c_add_r(21.69274, c_mul(t, den));
            den = c_add_r(39.27121, c_mul(t, den));
            den = c_add_r(38.82363, c_mul(t, den));
            den = c_add_r(16.4955, c_mul(t, den));
            w = c_div(num, den);
        } else { /* region IV */
            cplx u = c_mul(t, t);
            cplx num = c_rsub(1.320522, c_scale(0.56419, u));
            num = c_rsub(35.7668, c_mul(u, num));
            num = c_rsub(219.031, c_mul(u, num));
            num = c_rsub(1540.787, c_mul(u, num));
            num = c_rsub(3321.99, c_mul(u, num));
            num = c_rsub(36183.31, c_mul(u, num));
            num = c_mul(t, num);
            cplx den = c_rsub(1.84144, u);
            den = c_rsub(61.5704, c_mul(u, den));
            den = c_rsub(364.219, c_mul(u, den));
            den = c_rsub(2186.18, c_mul(u, den));
            den = c_rsub(9022.23, c_mul(u, den));
            den = c_rsub(24322.8, c_mul(u, den));
            den = c_rsub(32066.6, c_mul(u, den));
            cplx q = c_div(num, den);
            double eu = exp(u.re);
            w.re = eu * cos(u.im) - q.re;
            w.im = eu * sin(u.im) - q.im;
        }
    }
    return w.re;
}

double orc_voigt_profile(double a, double v, double dD) { return orc_humlicek_re(a, v) / (sqrt(PI) * dD); }

/* B_λ, src/radiation.jl:17-19, returned in kW m^-2 nm^-1 (SI x 1e-12); lambda in nm. */
double orc_B_lambda(double lambda_nm, double T) {
    double lam = lambda_nm * 1e-9;
    double lam5 = lam * lam * lam * lam * lam;
    return 2 * H_PLANCK * C_0 * C_0 / lam5 * 1 / (exp(H_PLANCK * C_0 / (lam * K_B * T)) - 1) * 1e-12;
}

/* γ_constant, src/broadening.jl:63-82 with Transparency.jl's closed forms. */
void orc_gamma_constant(const vrt_line* line, int64_t n, const double* T, const double* n_HI, const double* n_e, double* gamma) {
    for (int64_t i = 0; i < n; i++) {
        double g = line->c_unsold * pow(T[i], 0.3) * n_HI[i];
        g += line->gamma_natural;
        g += line->c_linear_stark * pow(n_e[i], 2.0 / 3.0);
        g += line->c_quadratic_stark * pow(T[i], 1.0 / 6.0) * n_e[i];
        gamma[i] = g;
    }
}

/* damping, src/broadening.jl:87-89: γ λ² / (4π c ΔD), λ and ΔD in nm -> dimensionless needs 1e-9 */
static inline double orc_damping(double gamma, double lambda_nm, double dD_nm) {
    return gamma * (lambda_nm * lambda_nm) / (4 * PI * C_0 * dD_nm) * 1e-9;
}

/* ------------------------------------------------------------------ J_λ_voronoi (line)
 * src/lambda_iteration.jl:60-113.  S, J, damping: nlam x n over the FULL wavelength set; [l0,l1) restricts
 * the wavelengths actually solved (used by the sharding tests; the reference always does all).
 * populations n x 3.  hoist as in orc_delaunay. */
void orc_J_lambda_voronoi(const orc_sites* s, const vrt_line* line, const double* lambda, const vrt_site_data* sd,
                          const vrt_quadrature* q, int n_sweeps, double p, const double* S, const double* pops,
                          double* J, double* damping, int64_t l0, int64_t l1, int hoist) {
    int64_t n = s->n, nlam = line->nlam;
    if (l1 <= l0) { l0 = 0; l1 = nlam; }
    /* a wavelength sub-range [l0, l1) (bounded CPU-baseline sample) only touches its own columns of J and damping: the
     * sample then carries its own share of this per-call work, not that of all nlam wavelengths */
    for (int64_t i = 0; i < n; i++)
        for (int64_t l = l0; l < l1; l++) J[l + nlam * i] = 0.0;
    double* gamma = (double*)malloc(sizeof(double) * (size_t)n);
    double* nHI = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; i++) nHI[i] = pops[i] + pops[i + n];
    orc_gamma_constant(line, n, sd->temperature, nHI, sd->electron_density, gamma);
    for (int64_t i = 0; i < n; i++)
        for (int64_t l = l0; l < l1; l++) damping[l + nlam * i] = orc_damping(gamma[i], lambda[l], sd->doppler_width[i]);
    double* vlos = (double*)malloc(sizeof(double) * (size_t)n);
    int64_t* st_up = NULL; double *st_w = NULL, *st_r = NULL;
    if (hoist) {
        st_up = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)n);
        st_w = (double*)malloc(sizeof(double) * 2 * (size_t)n);
        st_r = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    }
    double c_line = H_PLANCK * C_0 / (4 * PI * (line->lambda0 * 1e-9)); /* αline_λ, line.jl:219-225 */
    for (int64_t d = 0; d < q->n_dirs; d++) {
        double th = q->theta[d], ph = q->phi[d];
        double k[3] = {cos(th * PI / 180), cos(ph * PI / 180) * sin(th * PI / 180), sin(ph * PI / 180) * sin(th * PI / 180)};
        if (!(th > 90) && !(th < 90)) continue;
        int down = !(th > 90);
        /* line_of_sight_velocity(sites, -k), line.jl:198-208 */
        for (int64_t i = 0; i < n; i++)
            vlos[i] = sd->velocity_z[i] * (-k[0]) + sd->velocity_x[i] * (-k[1]) + sd->velocity_y[i] * (-k[2]);
        if (hoist) orc_stencil(s, k, p, st_up, NULL, st_w, st_r);
        const int64_t* perm = down ? s->perm_down : s->perm_up;
        int64_t n1 = (down ? s->layers_down : s->layers_up)[1] - 1;
#pragma omp parallel
        {
            double* Sv = (double*)malloc(sizeof(double) * (size_t)n);
            double* av = (double*)malloc(sizeof(double) * (size_t)n);
            double* Iv = (double*)malloc(sizeof(double) * (size_t)n);
            double* I0v = (double*)malloc(sizeof(double) * (size_t)(n1 > 0 ? n1 : 1));
#pragma omp for schedule(dynamic, 1)
            for (int64_t l = l0; l < l1; l++) {
                for (int64_t i = 0; i < n; i++) {
                    double dD = sd->doppler_width[i];
                    double v = (lambda[l] - line->lambda0 + line->lambda0 * vlos[i] / C_0) / dD; /* line.jl:132 */
                    double prof = orc_voigt_profile(damping[l + nlam * i], v, dD * 1e-9);         /* m^-1 */
                    av[i] = c_line * prof * (pops[i] * line->Bij - pops[i + n] * line->Bji) + sd->alpha_cont[i];
                    Sv[i] = S[l + nlam * i];
                }
                if (!down)
                    for (int64_t i = 0; i < n1; i++) I0v[i] = orc_B_lambda(lambda[l], sd->temperature[perm[i] - 1]);
                else
                    for (int64_t i = 0; i < n1; i++) I0v[i] = 0.0;
                orc_delaunay(s, down, k, Sv, I0v, av, n_sweeps, p, Iv, hoist, st_up, st_w, st_r);
                for (int64_t i = 0; i < n; i++) J[l + nlam * i] += q->weights[d] * Iv[i];
            }
            free(Sv); free(av); free(Iv); free(I0v);
        }
    }
    free(gamma); free(nHI); free(vlos); free(st_up); free(st_w); free(st_r);
}

/* J_λ_voronoi (continuum), src/lambda_continuum.jl:27-56: single wavelength, serial over directions;
 * I_0 = blackbody_λ(500 nm, T_bottom) = B0 at the bottom sites. */
void orc_J_continuum(const orc_sites* s, const vrt_quadrature* q, int n_sweeps, double p, const double* S,
                     const double* alpha, const double* B0, double* J, int hoist) {
    int64_t n = s->n;
    for (int64_t i = 0; i < n; i++) J[i] = 0.0;
    double* Iv = (double*)malloc(sizeof(double) * (size_t)n);
    double* I0v = (double*)malloc(sizeof(double) * (size_t)n);
    int64_t* st_up = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)n);
    double* st_w = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* st_r = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    for (int64_t d = 0; d < q->n_dirs; d++) {
        double th = q->theta[d], ph = q->phi[d];
        double k[3] = {cos(th * PI / 180), cos(ph * PI / 180) * sin(th * PI / 180), sin(ph * PI / 180) * sin(th * PI / 180)};
        if (!(th > 90) && !(th < 90)) continue;
        int down = !(th > 90);
        const int64_t* perm = down ? s->perm_down : s->perm_up;
        int64_t n1 = (down ? s->layers_down : s->layers_up)[1] - 1;
        for (int64_t i = 0; i < n1; i++) I0v[i] = down ? 0.0 : B0[perm[i] - 1];
        if (hoist) orc_stencil(s, k, p, st_up, NULL, st_w, st_r);
        orc_delaunay(s, down, k, S, I0v, alpha, n_sweeps, p, Iv, hoist, st_up, st_w, st_r);
        for (int64_t i = 0; i < n; i++) J[i] += q->weights[d] * Iv[i];
    }
    free(Iv); free(I0v); free(st_up); free(st_w); free(st_r);
}

/* ------------------------------------------------------------------ rates
 * gaunt_bf, src/rates.jl:562-572 (λ nm) */
static double orc_gaunt_bf(double lambda_nm, double charge, double n_eff) {
    double x = 1 / (lambda_nm * 1e-9 * R_INF * charge * charge);
    double x3 = pow(x, 1.0 / 3);
    double nsqx = 1 / (n_eff * n_eff * x);
    return 1 + 0.1728 * x3 * (1 - 2 * nsqx) - 0.0496 * (x3 * x3) * (1 - (1 - nsqx) * 0.66666667 * nsqx);
}

/* σic, src/rates.jl:422-438 (m²); λ_edge = last λ of the range, n_eff from (χj - χi) for both levels (Q9) */
void orc_sigma_ic(const vrt_line* line, const double* lam, int64_t nl, double* sigma) {
    double lam_edge = lam[nl - 1];
    double E_inf = R_INF * C_0 * H_PLANCK;
    double n_eff = sqrt(E_inf / (line->chi_j - line->chi_i));
    double charge = (double)line->Z;
    double sc = 4 * E_CHARGE * E_CHARGE / (3 * PI * sqrt(3.0) * EPS_0 * M_ELECTRON * C_0 * C_0 * R_INF);
    for (int64_t l = 0; l < nl; l++) {
        double r = lam[l] / lam_edge;
        sigma[l] = sc * (charge * charge * charge * charge) * n_eff * (r * r * r) * orc_gaunt_bf(lam[l], charge, n_eff);
    }
}

/* calculate_R(sites, ...), src/rates.jl:154-201 with Rij (:264-278 / :226-240), Rji (:348-364 / :305-321),
 * σij (:394-413), Gij (:467-484).  J, damping nlam x n (kW m^-2 nm^-1 / dimensionless); lte n x 3;
 * R 3 x 3 x n (s^-1), R[a + 3b + 9i] = R[a+1, b+1, i+1]. */
void orc_calculate_R(const vrt_line* line, const double* lambda, int64_t n, const double* T, const double* dD,
                     const double* J, const double* damping, const double* lte, double* R) {
    int64_t nlam = line->nlam;
    double hc = H_PLANCK * C_0;
    double pref = 2 * PI / hc;
    for (int64_t i = 0; i < 9 * n; i++) R[i] = 0.0;
    /* ionisation: level = 1, 2 -> 3 */
    for (int level = 1; level <= 2; level++) {
        int64_t start = line->lidx[level], stop = line->lidx[level + 1]; /* 0-based [start, stop) */
        int64_t nl = stop - start;
        double* sig = (double*)malloc(sizeof(double) * (size_t)nl);
        orc_sigma_ic(line, lambda + start, nl, sig);
#pragma omp parallel for
        for (int64_t i = 0; i < n; i++) {
            double n_ratio = lte[i + n * (level - 1)] / lte[i + n * 2];
            double up = 0.0, dn = 0.0;
            for (int64_t l = 0; l + 1 < nl; l++) {
                double la = lambda[start + l] * 1e-9, lb = lambda[start + l + 1] * 1e-9;
                double Ja = J[start + l + nlam * i] * 1e12, Jb = J[start + l + 1 + nlam * i] * 1e12;
                double Ga = n_ratio * exp(-hc / (K_B * la * T[i]));
                double Gb = n_ratio * exp(-hc / (K_B * lb * T[i]));
                up += pref * (la * sig[l] * Ja + lb * sig[l + 1] * Jb) * (lb - la) / 1000;
                dn += pref * (sig[l] * Ga * la * (2 * hc * C_0 / (la * la * la * la * la) + Ja) +
                              sig[l + 1] * Gb * lb * (2 * hc * C_0 / (lb * lb * lb * lb * lb) + Jb)) * (lb - la);
            }
            R[(level - 1) + 3 * 2 + 9 * i] = up; /* R[level, 3] */
            R[2 + 3 * (level - 1) + 9 * i] = dn; /* R[3, level] */
        }
        free(sig);
    }
    /* bound-bound 1 <-> 2 */
    {
        int64_t start = line->lidx[0], stop = line->lidx[1];
        int64_t nl = stop - start;
        double sc = hc / (4 * PI * (line->lambda0 * 1e-9)) * line->Bij;
#pragma omp parallel for
        for (int64_t i = 0; i < n; i++) {
            double n_ratio = lte[i] / lte[i + n];
            double up = 0.0, dn = 0.0;
            double sa = 0, Ga = 0;
            for (int64_t l = 0; l < nl; l++) {
                double la = lambda[start + l];
                double v = (la - line->lambda0) / dD[i];
                double sb = sc * orc_voigt_profile(damping[start + l + nlam * i], v, dD[i] * 1e-9);
                double Gb = n_ratio * exp(-hc / (K_B * (la * 1e-9) * T[i]));
                if (l > 0) {
                    double l_a = lambda[start + l - 1] * 1e-9, l_b = la * 1e-9;
                    double Ja = J[start + l - 1 + nlam * i] * 1e12, Jb = J[start + l + nlam * i] * 1e12;
                    up += pref * ((l_a * sa * Ja + l_b * sb * Jb) * (l_b - l_a)) / 1000;
                    dn += pref * (sa * Ga * l_a * (2 * H_PLANCK * C_0 * C_0 / (l_a * l_a * l_a * l_a * l_a) + Ja) +
                                  sb * Gb * l_b * (2 * H_PLANCK * C_0 * C_0 / (l_b * l_b * l_b * l_b * l_b) + Jb)) * (l_b - l_a);
                }
                sa = sb;
                Ga = Gb;
            }
            R[0 + 3 * 1 + 9 * i] = up; /* R[1,2] */
            R[1 + 3 * 0 + 9 * i] = dn; /* R[2,1] */
        }
    }
}

/* get_revised_populations, src/populations.jl:191-221: per-site 2x2 inv(A)*b (LU with partial pivoting,
 * as LAPACK getrf/getri would), n1 = N_H - n2 - n3.  pops n x 3. */
void orc_get_revised_populations(int64_t n, const double* R, const double* C, const double* NH, double* pops) {
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        double P[9];
        for (int a = 0; a < 9; a++) P[a] = R[a + 9 * i] + C[a + 9 * i];
#define PP(r, c) P[((r)-1) + 3 * ((c)-1)]
        double A11 = PP(1, 2) + PP(2, 1);
        double A12 = PP(1, 2) - PP(3, 2);
        A11 += PP(2, 3);
        double A22 = PP(1, 3) + PP(3, 1);
        double A21 = PP(1, 3) - PP(2, 3);
        A22 += PP(3, 2);
        double b1 = NH[i] * PP(1, 2), b2 = NH[i] * PP(1, 3);
#undef PP
        /* x = inv(A) b through LU with partial pivoting */
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
        pops[i + n] = x1;
        pops[i + 2 * n] = x2;
        pops[i] = NH[i] - (x1 + x2);
    }
}

/* criterion, src/lambda_iteration.jl:325-349: max_l max_i |1 - S_old/S_new| (NaN propagates like Julia's maximum) */
double orc_criterion(const double* S_new, const double* S_old, int64_t count, const uint8_t* mask, int64_t stride) {
    double diff = 0;
    int nan = 0;
    for (int64_t i = 0; i < count; i++) {
        if (mask && !mask[i / stride]) continue;
        double d = fabs(1 - S_old[i] / S_new[i]);
        if (d != d) nan = 1;
        else if (d > diff) diff = d;
    }
    return nan ? NAN : diff;
}

/* Λ_voronoi (line), src/lambda_iteration.jl:207-297.  S and pops are in/out: on entry S = B_0 (or a
 * restart state), pops = LTE populations.  conv (maxiter+1) receives the criterion values.  Returns
 * the number of iterations done. */
int orc_lambda_voronoi(const orc_sites* s, const vrt_line* line, const double* lambda, const vrt_site_data* sd,
                       const vrt_quadrature* q, int n_sweeps, double p, double eps, int maxiter,
                       double* S, double* J, double* pops, double* conv, int hoist) {
    int64_t n = s->n, nlam = line->nlam;
    size_t tot = (size_t)n * (size_t)nlam;
    double* S_old = (double*)calloc(tot, sizeof(double));
    double* damping = (double*)malloc(sizeof(double) * tot);
    double* R = (double*)malloc(sizeof(double) * 9 * (size_t)n);
    int i = 0;
    for (;;) {
        double diff = orc_criterion(S, S_old, (int64_t)tot, NULL, 1);
        if (conv) conv[i] = diff;
        if (!(diff > eps && i < maxiter)) break;
        memcpy(S_old, S, sizeof(double) * tot);
        orc_J_lambda_voronoi(s, line, lambda, sd, q, n_sweeps, p, S_old, pops, J, damping, 0, 0, hoist);
        for (int64_t c = 0; c < n; c++) {
            double e = sd->destruction[c];
            for (int64_t l = 0; l < nlam; l++)
                S[l + nlam * c] = (1 - e) * J[l + nlam * c] + e * orc_B_lambda(lambda[l], sd->temperature[c]);
        }
        orc_calculate_R(line, lambda, n, sd->temperature, sd->doppler_width, J, damping, sd->lte_pops, R);
        orc_get_revised_populations(n, R, sd->C, sd->hydrogen_density, pops);
        i++;
    }
    free(S_old); free(damping); free(R);
    return i;
}

/* ---- regular-grid twins of the line path.  The per-cell physics is the same code; only the formal solver differs
 * (orc_short_characteristics, vrt_oracle_regular.c).  Cells are the (nz, nx, ny) arrays flattened column-major. */
void orc_short_characteristics(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                               const double* k, int down, const double* S, const double* I_0, const double* al, int n_sweeps,
                               double* I, int32_t* plane_out);

/* J_λ_regular (line), src/lambda_iteration.jl:1-58: γ, damping (:13-21); per direction the Voigt profile with
 * v_los = velocity . (-k) (line.jl:80-96,175-190), α_tot (:32-35), I_0 = B_λ(λ_l, T[1,:,:]) for θ > 90 (:38) or zero (:46). */
void orc_J_lambda_regular(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                          const vrt_line* line, const double* lambda, const vrt_site_data* sd, const vrt_quadrature* q,
                          int n_sweeps, const double* S, const double* pops, double* J, double* damping) {
    int64_t n = nz * nx * ny, nlam = line->nlam, plane = nx * ny;
    for (int64_t i = 0; i < n * nlam; i++) J[i] = 0.0;
    double* gamma = (double*)malloc(sizeof(double) * (size_t)n);
    double* nHI = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; i++) nHI[i] = pops[i] + pops[i + n];
    orc_gamma_constant(line, n, sd->temperature, nHI, sd->electron_density, gamma);
    for (int64_t i = 0; i < n; i++)
        for (int64_t l = 0; l < nlam; l++) damping[l + nlam * i] = orc_damping(gamma[i], lambda[l], sd->doppler_width[i]);
    double* vlos = (double*)malloc(sizeof(double) * (size_t)n);
    double c_line = H_PLANCK * C_0 / (4 * PI * (line->lambda0 * 1e-9)); /* αline_λ, line.jl:219-225 */
    for (int64_t d = 0; d < q->n_dirs; d++) {
        double th = q->theta[d], ph = q->phi[d];
        double k[3] = {cos(th * PI / 180), cos(ph * PI / 180) * sin(th * PI / 180), sin(ph * PI / 180) * sin(th * PI / 180)};
        if (!(th > 90) && !(th < 90)) continue;
        int down = !(th > 90);
        for (int64_t i = 0; i < n; i++)
            vlos[i] = sd->velocity_z[i] * (-k[0]) + sd->velocity_x[i] * (-k[1]) + sd->velocity_y[i] * (-k[2]);
#pragma omp parallel
        {
            double* Sv = (double*)malloc(sizeof(double) * (size_t)n);
            double* av = (double*)malloc(sizeof(double) * (size_t)n);
            double* Iv = (double*)malloc(sizeof(double) * (size_t)n);
            double* I0v = (double*)malloc(sizeof(double) * (size_t)plane);
#pragma omp for schedule(dynamic, 1)
            for (int64_t l = 0; l < nlam; l++) {
                for (int64_t i = 0; i < n; i++) {
                    double dD = sd->doppler_width[i];
                    double v = (lambda[l] - line->lambda0 + line->lambda0 * vlos[i] / C_0) / dD; /* line.jl:91 */
                    double prof = orc_voigt_profile(damping[l + nlam * i], v, dD * 1e-9);         /* m^-1 */
                    av[i] = c_line * prof * (pops[i] * line->Bij - pops[i + n] * line->Bji) + sd->alpha_cont[i];
                    Sv[i] = S[l + nlam * i];
                }
                for (int64_t i = 0; i < plane; i++) I0v[i] = down ? 0.0 : orc_B_lambda(lambda[l], sd->temperature[nz * i]);
                orc_short_characteristics(nz, nx, ny, z, x, y, k, down, Sv, I0v, av, n_sweeps, Iv, NULL);
                for (int64_t i = 0; i < n; i++) J[l + nlam * i] += q->weights[d] * Iv[i];
            }
            free(Sv); free(av); free(Iv); free(I0v);
        }
    }
    free(gamma); free(nHI); free(vlos);
}

/* Λ_regular (line), src/lambda_iteration.jl:116-205 with criterion :299-323 */
int orc_lambda_regular_line(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                            const vrt_line* line, const double* lambda, const vrt_site_data* sd, const vrt_quadrature* q,
                            int n_sweeps, double eps, int maxiter, double* S, double* J, double* pops, double* conv) {
    int64_t n = nz * nx * ny, nlam = line->nlam;
    size_t tot = (size_t)n * (size_t)nlam;
    double* S_old = (double*)calloc(tot, sizeof(double));
    double* damping = (double*)malloc(sizeof(double) * tot);
    double* R = (double*)malloc(sizeof(double) * 9 * (size_t)n);
    int i = 0;
    for (;;) {
        double diff = orc_criterion(S, S_old, (int64_t)tot, NULL, 1);
        if (conv) conv[i] = diff;
        if (!(diff > eps && i < maxiter)) break;
        memcpy(S_old, S, sizeof(double) * tot);
        orc_J_lambda_regular(nz, nx, ny, z, x, y, line, lambda, sd, q, n_sweeps, S_old, pops, J, damping);
        for (int64_t c = 0; c < n; c++) {
            double e = sd->destruction[c];
            for (int64_t l = 0; l < nlam; l++)
                S[l + nlam * c] = (1 - e) * J[l + nlam * c] + e * orc_B_lambda(lambda[l], sd->temperature[c]);
        }
        orc_calculate_R(line, lambda, n, sd->temperature, sd->doppler_width, J, damping, sd->lte_pops, R);
        orc_get_revised_populations(n, R, sd->C, sd->hydrogen_density, pops);
        i++;
    }
    free(S_old); free(damping); free(R);
    return i;
}

/* Λ_voronoi (continuum), src/lambda_continuum.jl:109-160 with `criterion` over thick = ε > 1e-4 (:133,:181-198). */
int orc_lambda_continuum(const orc_sites* s, const vrt_quadrature* q, int n_sweeps, double p, double eps, int maxiter,
                         const double* alpha, const double* eps_l, const double* B0, double* S, double* J, double* conv, int hoist) {
    int64_t n = s->n;
    double* S_old = (double*)calloc((size_t)n, sizeof(double));
    uint8_t* thick = (uint8_t*)malloc((size_t)n);
    for (int64_t i = 0; i < n; i++) thick[i] = eps_l[i] > 1e-4;
    int i = 0;
    for (;;) {
        double diff = orc_criterion(S, S_old, n, thick, 1);
        if (conv) conv[i] = diff;
        if (!(diff > eps && i < maxiter)) break;
        memcpy(S_old, S, sizeof(double) * (size_t)n);
        orc_J_continuum(s, q, n_sweeps, p, S_old, alpha, B0, J, hoist);
        for (int64_t c = 0; c < n; c++) S[c] = (1 - eps_l[c]) * J[c] + eps_l[c] * B0[c];
        i++;
    }
    free(S_old); free(thick);
    return i;
}

/* LTE_populations(line, sites), src/populations.jl:112-138 (one-time host input; kept here so that the
 * oracle can build its own inputs in tests).  pops n x 3. */
void orc_LTE_populations(const vrt_line* line, int64_t n, const double* T, const double* ne, const double* NH, double* pops) {
    double chi[3] = {line->chi_i, line->chi_j, line->chi_inf};
    double g[3] = {(double)line->gi, (double)line->gj, 1.0};
    double saha_const = (K_B / H_PLANCK) * (2 * PI * M_ELECTRON) / H_PLANCK;
    for (int64_t i = 0; i < n; i++) {
        double saha_factor = 2 * (pow(saha_const * T[i], 1.5) / ne[i]);
        double r[3] = {1.0, 0, 0};
        for (int l = 1; l < 3; l++) r[l] = g[l] / g[0] * exp(-(chi[l] - chi[0]) / (K_B * T[i]));
        r[2] *= saha_factor;
        double r0 = 1 / (r[0] + r[1] + r[2]);
        pops[i] = r0 * NH[i];
        pops[i + n] = r[1] * r0 * NH[i];
        pops[i + 2 * n] = r[2] * r0 * NH[i];
    }
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline sets its thread count explicitly (JULIA_NUM_THREADS of the reference) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
