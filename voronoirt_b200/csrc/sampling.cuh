// sampling.cuh — host/device pieces of the site set-up: the trilinear interpolation of an atmosphere field
// (functions.jl:207-248) and the acceptance-rejection sampling of sites from it (rejection_sampling, functions.jl:79-121).
// Shared by initialise.cu's kernels and by the CPU harness of the tests (tests/voronoi_harness.cpp), which is how the
// arithmetic is checked without a GPU; every floating-point operation that decides an outcome is rounded on its own
// (no fma contraction), so host and device agree bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>
#include "voronoi_cell.cuh"   // VC_HD, vc_mul, vc_add

namespace vrt {

VC_HD double vc_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
VC_HD double vc_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

struct TriGrid {
    int64_t nz, nx, ny;
    const double *z, *x, *y;   // axes
    const double* vals;        // (nz, nx, ny) column-major
};

// Julia's searchsortedfirst(a, x) - 1 as a 0-based lower-corner index: number of elements < x, minus one
VC_HD int64_t tri_lower_corner(const double* a, int64_t n, double x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

VC_HD double tri_lerp(double c0, double c1, double t) { return vc_add(vc_mul(c0, vc_sub(1.0, t)), vc_mul(c1, t)); }   // c0*(1 - t) + c1*t

// trilinear(z_mrk, x_mrk, y_mrk, atmos, vals); false where the reference throws a BoundsError (point outside the axes)
VC_HD bool trilinear_at(const TriGrid& T, double zm, double xm, double ym, double* out) {
    const int64_t iz = tri_lower_corner(T.z, T.nz, zm), ix = tri_lower_corner(T.x, T.nx, xm), iy = tri_lower_corner(T.y, T.ny, ym);
    if (iz < 0 || iz > T.nz - 2 || ix < 0 || ix > T.nx - 2 || iy < 0 || iy > T.ny - 2 || !(zm == zm) || !(xm == xm) || !(ym == ym)) return false;
    const double z0 = T.z[iz], z1 = T.z[iz + 1], x0 = T.x[ix], x1 = T.x[ix + 1], y0 = T.y[iy], y1 = T.y[iy + 1];
    const double x_d = vc_div(vc_sub(xm, x0), vc_sub(x1, x0));
    const double y_d = vc_div(vc_sub(ym, y0), vc_sub(y1, y0));
    const double z_d = vc_div(vc_sub(zm, z0), vc_sub(z1, z0));
    const int64_t nz = T.nz, nx = T.nx;
#define VRT_TRI_V(a, b, c) T.vals[(a) + nz * ((b) + nx * (c))]
    const double c000 = VRT_TRI_V(iz, ix, iy), c010 = VRT_TRI_V(iz, ix, iy + 1), c100 = VRT_TRI_V(iz, ix + 1, iy), c110 = VRT_TRI_V(iz, ix + 1, iy + 1);
    const double c001 = VRT_TRI_V(iz + 1, ix, iy), c011 = VRT_TRI_V(iz + 1, ix, iy + 1), c101 = VRT_TRI_V(iz + 1, ix + 1, iy),
                 c111 = VRT_TRI_V(iz + 1, ix + 1, iy + 1);
#undef VRT_TRI_V
    const double c00 = tri_lerp(c000, c100, x_d), c01 = tri_lerp(c001, c101, x_d);
    const double c10 = tri_lerp(c010, c110, x_d), c11 = tri_lerp(c011, c111, x_d);
    const double c0 = tri_lerp(c00, c10, y_d), c1 = tri_lerp(c01, c11, y_d);
    *out = tri_lerp(c0, c1, z_d);
    return true;
}

// initialiseII (voronoi_utils.jl:716-770): the value at the nearest of the eight corners of the enclosing cell, corners in
// the reference's order (y fastest, then x, then z), distances as `euclidean` computes them, first minimum wins
VC_HD bool nearest_corner_at(const TriGrid& T, double zm, double xm, double ym, double* out) {
    const int64_t iz = tri_lower_corner(T.z, T.nz, zm), ix = tri_lower_corner(T.x, T.nx, xm), iy = tri_lower_corner(T.y, T.ny, ym);
    if (iz < 0 || iz > T.nz - 2 || ix < 0 || ix > T.nx - 2 || iy < 0 || iy > T.ny - 2 || !(zm == zm) || !(xm == xm) || !(ym == ym)) return false;
    double best = 0.0;
    int64_t arg = -1;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++)
            for (int c = 0; c < 2; c++) {
                const double ez = vc_sub(T.z[iz + a], zm), ex = vc_sub(T.x[ix + b], xm), ey = vc_sub(T.y[iy + c], ym);
                const double d = sqrt(vc_add(vc_add(vc_mul(ez, ez), vc_mul(ex, ex)), vc_mul(ey, ey)));
                if (arg < 0 || d < best) { best = d; arg = (iz + a) + T.nz * ((ix + b) + T.nx * (iy + c)); }
            }
    *out = T.vals[arg];
    return true;
}

// Philox4x32-10 (Salmon, Moraes, Dror & Shaw 2011): counter-based, so site i's candidate stream does not depend on how
// the sites are spread over threads.  (The reference draws from Julia's task-local Xoshiro stream, which cannot be
// reproduced; what is kept is the acceptance rule and therefore the distribution.)
VC_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// 53-bit uniform in [0, 1) from two 32-bit words
VC_HD double u01_53(uint32_t hi, uint32_t lo) {
    return (double)(((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// One site by acceptance-rejection (functions.jl:100-118): candidate uniform in the box of the axes, accepted when
// trilinear(quantity) > U(q_min, q_max).  Trial t of site i uses the Philox counters (i, t, 0) and (i, t, 1).
// A candidate on the lower faces of the box (u = 0, where the reference would throw) counts as rejected.
VC_HD int64_t rejection_site(const TriGrid& T, uint64_t seed, int64_t i, double q_min, double dq, int64_t max_trials, double out[3]) {
    const double z_min = T.z[0], x_min = T.x[0], y_min = T.y[0];
    const double dz = vc_sub(T.z[T.nz - 1], z_min), dx = vc_sub(T.x[T.nx - 1], x_min), dy = vc_sub(T.y[T.ny - 1], y_min);
    for (int64_t t = 0; t < max_trials; t++) {
        uint32_t a[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)t, 0u};
        uint32_t b[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)t, 1u};
        philox4x32_10(a, (uint32_t)seed, (uint32_t)(seed >> 32));
        philox4x32_10(b, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double z_ref = vc_add(vc_mul(u01_53(a[0], a[1]), dz), z_min);
        const double x_ref = vc_add(vc_mul(u01_53(a[2], a[3]), dx), x_min);
        const double y_ref = vc_add(vc_mul(u01_53(b[0], b[1]), dy), y_min);
        double density_ref;
        if (!trilinear_at(T, z_ref, x_ref, y_ref, &density_ref)) continue;
        const double density_ran = vc_add(vc_mul(u01_53(b[2], b[3]), dq), q_min);
        if (density_ref > density_ran) {
            out[0] = z_ref; out[1] = x_ref; out[2] = y_ref;
            return t + 1;
        }
    }
    return -1;
}

}  // namespace vrt
