// physics.cuh — device-side physics shared by the opacity (K4), source (K6) and rates (K7) kernels.
// Constants are CODATA 2018 as used by the reference (src/atmosphere.jl:1-8 via PhysicalConstants).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace vrt {

constexpr double H_PLANCK = 6.62607015e-34;
constexpr double K_B = 1.380649e-23;
constexpr double C_0 = 299792458.0;
constexpr double E_CHARGE = 1.602176634e-19;
constexpr double M_ELECTRON = 9.1093837015e-31;
constexpr double EPS_0 = 8.8541878128e-12;
constexpr double R_INF = 10973731.568160;
constexpr double PI = 3.14159265358979323846;
constexpr double INV_SQRT_PI = 0.5641895835477563;

struct cplx {
    double re, im;
};
__host__ __device__ __forceinline__ cplx c_mul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__host__ __device__ __forceinline__ cplx c_add_r(double x, cplx a) { return {x + a.re, a.im}; }
__host__ __device__ __forceinline__ cplx c_rsub(double x, cplx a) { return {x - a.re, -a.im}; }
__host__ __device__ __forceinline__ cplx c_scale(double x, cplx a) { return {x * a.re, x * a.im}; }
__host__ __device__ __forceinline__ cplx c_div(cplx a, cplx b) {
    double den = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / den, (a.im * b.re - a.re * b.im) / den};
}
// Re(a / b): one division (the Voigt function only needs the real part of the Faddeeva function)
__host__ __device__ __forceinline__ double c_div_re(cplx a, cplx b) {
    return (a.re * b.re + a.im * b.im) / (b.re * b.re + b.im * b.im);
}

// Coefficients of Humlíček's w4 regions III and IV, in the constant bank: as literals every one of them costs two UMOV issue
// slots in front of its DFMA (14 % of the instructions k_opacity executed in the ncu source view).
static __constant__ double HUM3N[5] = {0.5642236, 3.778987, 11.96482, 20.20933, 16.4955};
static __constant__ double HUM3D[5] = {6.699398, 21.69274, 39.27121, 38.82363, 16.4955};
static __constant__ double HUM4N[7] = {0.56419, 1.320522, 35.7668, 219.031, 1540.787, 3321.99, 36183.31};
static __constant__ double HUM4D[7] = {1.84144, 61.5704, 364.219, 2186.18, 9022.23, 24322.8, 32066.6};

// Re w(v + i a), Humlíček (1982) w4 — what Transparency.jl's voigt_profile evaluates
// (call sites: reference src/line.jl:133, src/rates.jl:408).
__device__ __forceinline__ double humlicek_re(double a, double v) {
    cplx z = {v, a};
    double s = fabs(v) + a;
    if (s > 15.0) {
        cplx zz = c_mul(z, z);
        cplx den = {zz.re - 0.5, zz.im};
        cplx num = {-INV_SQRT_PI * z.im, INV_SQRT_PI * z.re};
        return c_div_re(num, den);
    } else if (s > 5.5) {
        cplx zz = c_mul(z, z);
        cplx t1 = {zz.re * INV_SQRT_PI - 1.4104739589, zz.im * INV_SQRT_PI};
        cplx zt = c_mul(z, t1);
        cplx num = {-zt.im, zt.re};
        cplx zz3 = {zz.re - 3.0, zz.im};
        cplx den = c_add_r(0.75, c_mul(zz, zz3));
        return c_div_re(num, den);
    } else {
        double x = v, y = a;
        cplx t = {y, -x};
        if (y >= 0.195 * fabs(x) - 0.176) {
            cplx num = c_add_r(HUM3N[1], c_scale(HUM3N[0], t));
            num = c_add_r(HUM3N[2], c_mul(t, num));
            num = c_add_r(HUM3N[3], c_mul(t, num));
            num = c_add_r(HUM3N[4], c_mul(t, num));
            cplx den = c_add_r(HUM3D[0], t);
            den = c_add_r(HUM3D[1], c_mul(t, den));
            den = c_add_r(HUM3D[2], c_mul(t, den));
            den = c_add_r(HUM3D[3], c_mul(t, den));
            den = c_add_r(HUM3D[4], c_mul(t, den));
            return c_div_re(num, den);
        } else {
            cplx u = c_mul(t, t);
            cplx num = c_rsub(HUM4N[1], c_scale(HUM4N[0], u));
            num = c_rsub(HUM4N[2], c_mul(u, num));
            num = c_rsub(HUM4N[3], c_mul(u, num));
            num = c_rsub(HUM4N[4], c_mul(u, num));
            num = c_rsub(HUM4N[5], c_mul(u, num));
            num = c_rsub(HUM4N[6], c_mul(u, num));
            num = c_mul(t, num);
            cplx den = c_rsub(HUM4D[0], u);
            den = c_rsub(HUM4D[1], c_mul(u, den));
            den = c_rsub(HUM4D[2], c_mul(u, den));
            den = c_rsub(HUM4D[3], c_mul(u, den));
            den = c_rsub(HUM4D[4], c_mul(u, den));
            den = c_rsub(HUM4D[5], c_mul(u, den));
            den = c_rsub(HUM4D[6], c_mul(u, den));
            return exp(u.re) * cos(u.im) - c_div_re(num, den);
        }
    }
}

// voigt_profile(a, v, ΔD) with ΔD in metres -> m^-1
__device__ __forceinline__ double voigt_profile(double a, double v, double dD_m) { return humlicek_re(a, v) / (sqrt(PI) * dD_m); }

// B_λ (reference src/radiation.jl:17-19) in kW m^-2 nm^-1, λ in nm
__device__ __forceinline__ double B_lambda(double lambda_nm, double T) {
    double lam = lambda_nm * 1e-9;
    double lam5 = lam * lam * lam * lam * lam;
    return 2 * H_PLANCK * C_0 * C_0 / lam5 * 1 / (exp(H_PLANCK * C_0 / (lam * K_B * T)) - 1) * 1e-12;
}

// damping (reference src/broadening.jl:87-89), λ and ΔD in nm
__device__ __forceinline__ double damping_param(double gamma, double lambda_nm, double dD_nm) {
    return gamma * (lambda_nm * lambda_nm) / (4 * PI * C_0 * dD_nm) * 1e-9;
}

}  // namespace vrt
