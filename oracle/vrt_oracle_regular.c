/*
 * vrt_oracle_regular.c — CPU restatement of the reference's regular-grid short-characteristics solver
 * (src/characteristics.jl:19-835 with bilinear src/functions.jl:328-355, xy_intersect :430-457,
 * range_bounds :466-475, linear_weights :484-500, trapezoidal :392-395).
 *
 * TEST INFRASTRUCTURE ONLY (see vrt_oracle.c).  Unlike the irregular path this part IS pinned by the reference's
 * own data: data/searchlight_data/I_160_45_regular.npy and I_20_15_regular.npy are reproduced to <= 1e-15
 * (tests/test_regular_golden.py).  Every quirk of SURVEY App. A Q13 is kept: the yz/xz branches sample S, α and the
 * upwind-plane intensity at column idx+sign (the NEXT column of the loop) while the carried row is the previous one;
 * xz_down_ray takes α_centre/S_centre from the UPPER plane; yz_up_ray updates the x ghost rows inside the sweep loop.
 *
 * Arrays use the Julia layouts: S_0, α, I are (nz, nx, ny) column-major (z fastest) INCLUDING the periodic ghost
 * columns in x and y; I_0 is (nx, ny).  Indices below are 1-based like the source, through the macros.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t nz, nx, ny;
    const double *z, *x, *y;
} reg_atmos;

#define ORC_PI 3.14159265358979323846
#define A3(a, iz, ix, iy) ((a)[((iz)-1) + at->nz * (((ix)-1) + at->nx * ((iy)-1))])
#define P2(a, ix, iy) ((a)[((ix)-1) + at->nx * ((iy)-1)])

static void reg_linear_weights(double dtau, double* a, double* b, double* e) {
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

/* bilinear (functions.jl:328-355): vals = [Q11 Q12; Q21 Q22] with the FIRST coordinate selecting the row */
static double reg_bilinear(double xm, double ym, double x1, double x2, double y1, double y2, double Q11, double Q12, double Q21, double Q22) {
    double dx = x2 - x1, dy = y2 - y1;
    double f1 = ((x2 - xm) * Q11 + (xm - x1) * Q21) / dx;
    double f2 = ((x2 - xm) * Q12 + (xm - x1) * Q22) / dx;
    return ((y2 - ym) * f1 + (ym - y1) * f2) / dy;
}

/* xy_intersect(k) (functions.jl:430-457) */
static void reg_xy_intersect(const double* k, int* sx, int* sy) {
    if (k[1] > 0 && k[2] > 0) { *sx = -1; *sy = -1; }
    else if (k[1] < 0 && k[2] > 0) { *sx = 1; *sy = -1; }
    else if (k[1] < 0 && k[2] < 0) { *sx = 1; *sy = 1; }
    else if (k[1] > 0 && k[2] < 0) { *sx = -1; *sy = 1; }
    else { *sx = 1; *sy = 1; }
}
/* range_bounds (functions.jl:466-475) */
static void reg_range_bounds(int sign, int64_t bound, int64_t* start, int64_t* stop) {
    if (sign == 1) { *start = 2; *stop = bound - 1; }
    else { *start = bound - 1; *stop = 2; }
}

/* xy_up_ray / xy_down_ray (characteristics.jl:191-280, :290-373): idzu = upwind plane (idz-1 / idz+1) */
static void reg_xy_ray(const reg_atmos* at, const double* k, int64_t idz, int64_t idzu, int sx, int sy, const double* I_0,
                       const double* S, const double* al, double* I) {
    int64_t nx = at->nx, ny = at->ny;
    memset(I, 0, sizeof(double) * (size_t)(nx * ny));
    double dz = at->z[idzu - 1] - at->z[idz - 1];
    double r = fabs(dz / k[0]);
    double xinc = r * k[1], yinc = r * k[2];
    for (int64_t idx = 2; idx <= nx - 1; idx++) {
        int64_t ixl = idx - (sx + 1) / 2, ixu = ixl + 1;
        double xup = at->x[idx - 1] + xinc;
        double xb1 = at->x[ixl - 1], xb2 = at->x[ixu - 1];
        for (int64_t idy = 2; idy <= ny - 1; idy++) {
            double yup = at->y[idy - 1] + yinc;
            int64_t iyl = idy - (sy + 1) / 2, iyu = iyl + 1;
            double yb1 = at->y[iyl - 1], yb2 = at->y[iyu - 1];
            double a_c = A3(al, idz, idx, idy);
            double a_u = reg_bilinear(xup, yup, xb1, xb2, yb1, yb2, A3(al, idzu, ixl, iyl), A3(al, idzu, ixl, iyu), A3(al, idzu, ixu, iyl), A3(al, idzu, ixu, iyu));
            double dtau = r * (a_c + a_u) / 2;
            double S_c = A3(S, idz, idx, idy);
            double S_u = reg_bilinear(xup, yup, xb1, xb2, yb1, yb2, A3(S, idzu, ixl, iyl), A3(S, idzu, ixl, iyu), A3(S, idzu, ixu, iyl), A3(S, idzu, ixu, iyu));
            double a, b, e;
            reg_linear_weights(dtau, &a, &b, &e);
            double I_u = reg_bilinear(xup, yup, xb1, xb2, yb1, yb2, P2(I_0, ixl, iyl), P2(I_0, ixl, iyu), P2(I_0, ixu, iyl), P2(I_0, ixu, iyu));
            P2(I, idx, idy) = e * I_u + a * S_u + b * S_c;
        }
        P2(I, idx, 1) = P2(I, idx, ny - 1);
        P2(I, idx, ny) = P2(I, idx, 2);
    }
    for (int64_t iy = 1; iy <= ny; iy++) {
        P2(I, 1, iy) = P2(I, nx - 1, iy);
        P2(I, nx, iy) = P2(I, 2, iy);
    }
}

/* yz_up_ray (:383-486) and yz_down_ray (:496-604).  up != 0: z_bounds = (z[idz-1], z[idz]), centre values from the
 * upper plane (= idz), I_vals = [I_0 row; carried row]; down: z_bounds = (z[idz], z[idz+1]), centre from the lower
 * plane (= idz), I_vals = [carried row; I_0 row].  The x ghost rows are refreshed inside the sweep loop for up,
 * after it for down (Q13). */
static void reg_yz_ray(const reg_atmos* at, const double* k, int64_t idz, int up, int sx, int sy, const double* I_0,
                       const double* S, const double* al, int n_sweeps, double* I) {
    int64_t nx = at->nx, ny = at->ny;
    double dx = at->x[1] - at->x[0];
    int64_t x0, x1, y0, y1;
    reg_range_bounds(sx, nx, &x0, &x1);
    reg_range_bounds(sy, ny, &y0, &y1);
    memset(I, 0, sizeof(double) * (size_t)(nx * ny));
    double* carried = (double*)calloc((size_t)ny, sizeof(double));
    int64_t izl = up ? idz - 1 : idz, izu = up ? idz : idz + 1;
    double r = fabs(dx / k[1]);
    double zinc = r * k[0], yinc = r * k[2];
    double zup = at->z[idz - 1] + zinc;
    double zb1 = at->z[izl - 1], zb2 = at->z[izu - 1];
    for (int sweep = 1; sweep <= n_sweeps; sweep++) {
        for (int64_t idx = x0; sx > 0 ? idx <= x1 : idx >= x1; idx += sx) {
            int64_t ixu = idx + sx;
            for (int64_t idy = y0; sy > 0 ? idy <= y1 : idy >= y1; idy += sy) {
                int64_t iyl = idy - (sy + 1) / 2, iyu = iyl + 1;
                double yup = at->y[idy - 1] + yinc;
                double yb1 = at->y[iyl - 1], yb2 = at->y[iyu - 1];
                double a_c = A3(al, idz, idx, idy);
                double a_u = reg_bilinear(zup, yup, zb1, zb2, yb1, yb2, A3(al, izl, ixu, iyl), A3(al, izl, ixu, iyu), A3(al, izu, ixu, iyl), A3(al, izu, ixu, iyu));
                double dtau = r * (a_c + a_u) / 2;
                double S_c = A3(S, idz, idx, idy);
                double S_u = reg_bilinear(zup, yup, zb1, zb2, yb1, yb2, A3(S, izl, ixu, iyl), A3(S, izl, ixu, iyu), A3(S, izu, ixu, iyl), A3(S, izu, ixu, iyu));
                double a, b, e;
                reg_linear_weights(dtau, &a, &b, &e);
                double I_u = up ? reg_bilinear(zup, yup, zb1, zb2, yb1, yb2, P2(I_0, ixu, iyl), P2(I_0, ixu, iyu), carried[iyl - 1], carried[iyu - 1])
                                : reg_bilinear(zup, yup, zb1, zb2, yb1, yb2, carried[iyl - 1], carried[iyu - 1], P2(I_0, ixu, iyl), P2(I_0, ixu, iyu));
                P2(I, idx, idy) = e * I_u + a * S_u + b * S_c;
            }
            P2(I, idx, 1) = P2(I, idx, ny - 1);
            P2(I, idx, ny) = P2(I, idx, 2);
            for (int64_t iy = 1; iy <= ny; iy++) carried[iy - 1] = P2(I, idx, iy);
        }
        if (up)
            for (int64_t iy = 1; iy <= ny; iy++) { P2(I, 1, iy) = P2(I, nx - 1, iy); P2(I, nx, iy) = P2(I, 2, iy); }
    }
    if (!up)
        for (int64_t iy = 1; iy <= ny; iy++) { P2(I, 1, iy) = P2(I, nx - 1, iy); P2(I, nx, iy) = P2(I, 2, iy); }
    free(carried);
}

/* xz_up_ray (:614-718) and xz_down_ray (:728-835).  Note: BOTH take α_centre / S_centre from `α_upper` / `S_upper`,
 * which for the down ray is the plane idz+1, not idz (Q13). */
static void reg_xz_ray(const reg_atmos* at, const double* k, int64_t idz, int up, int sx, int sy, const double* I_0,
                       const double* S, const double* al, int n_sweeps, double* I) {
    int64_t nx = at->nx, ny = at->ny;
    double dy = at->y[1] - at->y[0];
    int64_t x0, x1, y0, y1;
    reg_range_bounds(sx, nx, &x0, &x1);
    reg_range_bounds(sy, ny, &y0, &y1);
    memset(I, 0, sizeof(double) * (size_t)(nx * ny));
    double* carried = (double*)calloc((size_t)nx, sizeof(double));
    int64_t izl = up ? idz - 1 : idz, izu = up ? idz : idz + 1;
    double r = fabs(dy / k[2]);
    double zinc = r * k[0], xinc = r * k[1];
    double zup = at->z[idz - 1] + zinc;
    double zb1 = at->z[izl - 1], zb2 = at->z[izu - 1];
    for (int sweep = 1; sweep <= n_sweeps; sweep++) {
        for (int64_t idy = y0; sy > 0 ? idy <= y1 : idy >= y1; idy += sy) {
            int64_t iyu = idy + sy;
            for (int64_t idx = x0; sx > 0 ? idx <= x1 : idx >= x1; idx += sx) {
                int64_t ixl = idx - (sx + 1) / 2, ixh = ixl + 1;
                double xup = at->x[idx - 1] + xinc;
                double xb1 = at->x[ixl - 1], xb2 = at->x[ixh - 1];
                double a_c = A3(al, izu, idx, idy);
                double a_u = reg_bilinear(zup, xup, zb1, zb2, xb1, xb2, A3(al, izl, ixl, iyu), A3(al, izl, ixh, iyu), A3(al, izu, ixl, iyu), A3(al, izu, ixh, iyu));
                double dtau = r * (a_c + a_u) / 2;
                double S_c = A3(S, izu, idx, idy);
                double S_u = reg_bilinear(zup, xup, zb1, zb2, xb1, xb2, A3(S, izl, ixl, iyu), A3(S, izl, ixh, iyu), A3(S, izu, ixl, iyu), A3(S, izu, ixh, iyu));
                double a, b, e;
                reg_linear_weights(dtau, &a, &b, &e);
                double I_u = up ? reg_bilinear(zup, xup, zb1, zb2, xb1, xb2, P2(I_0, ixl, iyu), P2(I_0, ixh, iyu), carried[ixl - 1], carried[ixh - 1])
                                : reg_bilinear(zup, xup, zb1, zb2, xb1, xb2, carried[ixl - 1], carried[ixh - 1], P2(I_0, ixl, iyu), P2(I_0, ixh, iyu));
                P2(I, idx, idy) = e * I_u + a * S_u + b * S_c;
            }
            P2(I, 1, idy) = P2(I, nx - 1, idy);
            P2(I, nx, idy) = P2(I, 2, idy);
            for (int64_t ix = 1; ix <= nx; ix++) carried[ix - 1] = P2(I, ix, idy);
        }
    }
    for (int64_t ix = 1; ix <= nx; ix++) { P2(I, ix, 1) = P2(I, ix, ny - 1); P2(I, ix, ny) = P2(I, ix, 2); }
    free(carried);
}

/* short_characteristics_up (down = 0, :19-95) / short_characteristics_down (down = 1, :110-180).
 * plane_out (optional, nz) receives the branch taken per plane (1 xy, 2 yz, 3 xz; 0 for the boundary plane). */
void orc_short_characteristics(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                               const double* k, int down, const double* S, const double* I_0, const double* al, int n_sweeps,
                               double* I, int32_t* plane_out) {
    reg_atmos A = {nz, nx, ny, z, x, y};
    const reg_atmos* at = &A;
    memset(I, 0, sizeof(double) * (size_t)(nz * nx * ny));
    double dx = x[1] - x[0], dy = y[1] - y[0];
    double r_x = fabs(dx / k[1]), r_y = fabs(dy / k[2]);
    int sx, sy;
    reg_xy_intersect(k, &sx, &sy);
    double* prev = (double*)malloc(sizeof(double) * (size_t)(nx * ny));
    double* cur = (double*)malloc(sizeof(double) * (size_t)(nx * ny));
    int64_t zb = down ? nz : 1;
    for (int64_t iy = 1; iy <= ny; iy++)
        for (int64_t ix = 1; ix <= nx; ix++) {
            A3(I, zb, ix, iy) = P2(I_0, ix, iy);
            P2(prev, ix, iy) = P2(I_0, ix, iy);
        }
    if (plane_out) plane_out[zb - 1] = 0;
    for (int64_t step = 1; step < nz; step++) {
        int64_t idz = down ? nz - step : 1 + step;
        double dz = down ? z[idz] - z[idz - 1] : z[idz - 1] - z[idz - 2];
        double r_z = fabs(dz / k[0]);
        /* argmin([r_z, r_x, r_y]): first minimum wins; NaN never arises for finite non-zero k */
        int cut = 1;
        double best = r_z;
        if (r_x < best) { best = r_x; cut = 2; }
        if (r_y < best) { best = r_y; cut = 3; }
        if (cut == 1) reg_xy_ray(at, k, idz, down ? idz + 1 : idz - 1, sx, sy, prev, S, al, cur);
        else if (cut == 2) reg_yz_ray(at, k, idz, !down, sx, sy, prev, S, al, n_sweeps, cur);
        else reg_xz_ray(at, k, idz, !down, sx, sy, prev, S, al, n_sweeps, cur);
        if (plane_out) plane_out[idz - 1] = cut;
        for (int64_t iy = 1; iy <= ny; iy++)
            for (int64_t ix = 1; ix <= nx; ix++) A3(I, idz, ix, iy) = P2(cur, ix, iy);
        double* t = prev; prev = cur; cur = t;
    }
    free(prev);
    free(cur);
}

/* J_λ_regular, continuum form (lambda_continuum.jl:1-24): J = Σ_i w_i I_i; θ > 90 up from I0_up at z[1], θ < 90 down
 * from zero at z[end]; θ = 90 contributes nothing.  The source passes I_0 as a keyword to a positional parameter (Q12)
 * and cannot run as shipped; this is the evident intent (the same as the Voronoi twin, lambda_continuum.jl:27-56). */
void orc_J_regular(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, int64_t n_dirs,
                   const double* weights, const double* theta, const double* phi, int n_sweeps, const double* S, const double* al,
                   const double* I0_up, double* J) {
    size_t vol = (size_t)(nz * nx * ny);
    double* I = (double*)malloc(sizeof(double) * vol);
    double* zero = (double*)calloc((size_t)(nx * ny), sizeof(double));
    memset(J, 0, sizeof(double) * vol);
    for (int64_t i = 0; i < n_dirs; i++) {
        double th = theta[i], ph = phi[i];
        double k[3] = {cos(th * ORC_PI / 180), cos(ph * ORC_PI / 180) * sin(th * ORC_PI / 180), sin(ph * ORC_PI / 180) * sin(th * ORC_PI / 180)};
        if (th > 90) orc_short_characteristics(nz, nx, ny, z, x, y, k, 0, S, I0_up, al, n_sweeps, I, NULL);
        else if (th < 90) orc_short_characteristics(nz, nx, ny, z, x, y, k, 1, S, zero, al, n_sweeps, I, NULL);
        else continue;
        for (size_t c = 0; c < vol; c++) J[c] += weights[i] * I[c];
    }
    free(I);
    free(zero);
}

/* Λ_regular (lambda_continuum.jl:58-107) with criterion (:162-179): S = B_0; while max over thick (ε > 1e-4) of
 * |1 - S_old/S_new| > ϵ and i < maxiter: J = J_λ_regular(S_old); S_new = (1-ε)J + εB_0.  conv[i] = criterion value
 * evaluated before iteration i+1.  Returns the number of iterations. */
int orc_lambda_regular(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, int64_t n_dirs,
                       const double* weights, const double* theta, const double* phi, int n_sweeps, double eps, int maxiter,
                       const double* al, const double* eps_l, const double* B0, double* S, double* J, double* conv) {
    size_t vol = (size_t)(nz * nx * ny);
    double* S_old = (double*)calloc(vol, sizeof(double));
    double* I0 = (double*)malloc(sizeof(double) * (size_t)(nx * ny));
    for (int64_t iy = 0; iy < ny; iy++)
        for (int64_t ix = 0; ix < nx; ix++) I0[ix + nx * iy] = B0[0 + nz * (ix + nx * iy)];
    memcpy(S, B0, sizeof(double) * vol);
    memset(J, 0, sizeof(double) * vol);
    int i = 0;
    for (;;) {
        double diff = 0;
        int isnan_ = 0;
        for (size_t c = 0; c < vol; c++)
            if (eps_l[c] > 1e-4) {
                double d = fabs(1 - S_old[c] / S[c]);
                if (d != d) isnan_ = 1;
                else if (d > diff) diff = d;
            }
        if (isnan_) diff = NAN;
        conv[i] = diff;
        if (!(diff > eps && i < maxiter)) break;
        memcpy(S_old, S, sizeof(double) * vol);
        orc_J_regular(nz, nx, ny, z, x, y, n_dirs, weights, theta, phi, n_sweeps, S_old, al, I0, J);
        for (size_t c = 0; c < vol; c++) S[c] = (1 - eps_l[c]) * J[c] + eps_l[c] * B0[c];
        i++;
    }
    free(S_old);
    free(I0);
    return i;
}

/* trilinear (functions.jl:207-248) broadcast over n sites; vals (nz, nx, ny) column-major, pos 3 x n rows (z, x, y).
 * searchsortedfirst(a, x) - 1 = (number of elements < x) - 1 as a 0-based lower corner.  Out-of-range sites (BoundsError
 * in Julia) give NaN; returns their number.  Compiled with -ffp-contract=off: every operation rounds on its own. */
static int64_t orc_lower_corner(const double* a, int64_t n, double x) {
    int64_t c = 0;
    while (c < n && a[c] < x) c++;
    return c - 1;
}
int64_t orc_trilinear(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, const double* vals,
                      int64_t n, const double* pos, double* out) {
    int64_t bad = 0;
    for (int64_t k = 0; k < n; k++) {
        double zm = pos[3 * k], xm = pos[3 * k + 1], ym = pos[3 * k + 2];
        int64_t iz = orc_lower_corner(z, nz, zm), ix = orc_lower_corner(x, nx, xm), iy = orc_lower_corner(y, ny, ym);
        if (iz < 0 || iz > nz - 2 || ix < 0 || ix > nx - 2 || iy < 0 || iy > ny - 2) { out[k] = NAN; bad++; continue; }
        double x_d = (xm - x[ix]) / (x[ix + 1] - x[ix]);
        double y_d = (ym - y[iy]) / (y[iy + 1] - y[iy]);
        double z_d = (zm - z[iz]) / (z[iz + 1] - z[iz]);
#define VV(a, b, c) vals[(a) + nz * ((b) + nx * (c))]
        double c000 = VV(iz, ix, iy), c010 = VV(iz, ix, iy + 1), c100 = VV(iz, ix + 1, iy), c110 = VV(iz, ix + 1, iy + 1);
        double c001 = VV(iz + 1, ix, iy), c011 = VV(iz + 1, ix, iy + 1), c101 = VV(iz + 1, ix + 1, iy), c111 = VV(iz + 1, ix + 1, iy + 1);
#undef VV
        double c00 = c000 * (1 - x_d) + c100 * x_d;
        double c01 = c001 * (1 - x_d) + c101 * x_d;
        double c10 = c010 * (1 - x_d) + c110 * x_d;
        double c11 = c011 * (1 - x_d) + c111 * x_d;
        double c0 = c00 * (1 - y_d) + c10 * y_d;
        double c1 = c01 * (1 - y_d) + c11 * y_d;
        out[k] = c0 * (1 - z_d) + c1 * z_d;
    }
    return bad;
}
