// initialise.cu — site initialisation: trilinear interpolation of an atmosphere field to the Voronoi sites
// (first half; reference: trilinear src/functions.jl:207-248, initialise src/voronoi_utils.jl:687-708, which
// calls it for temperature, electron density, hydrogen density and the three velocity components).
//
// One thread per site: three binary searches with searchsortedfirst semantics (first index whose value is >= the
// coordinate, minus one = lower corner), eight loads, and the seven lerps in the order the reference writes them, each
// product and sum rounded separately (no fma contraction), so the result equals a plain IEEE evaluation of the Julia
// expressions bit for bit.  A site outside the axes (where Julia throws a BoundsError) yields NaN and an error return.
#include <math.h>
#include "vrt_internal.h"

namespace vrt {
namespace {

// Julia's searchsortedfirst(a, x) - 1 as a 0-based lower-corner index: number of elements < x, minus one
__device__ __forceinline__ int64_t lower_corner(const double* __restrict__ a, int64_t n, double x) {
    int64_t lo = 0, hi = n;   // first index in [0, n] with a[i] >= x
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

__device__ __forceinline__ double lerp_rn(double c0, double c1, double t) {
    // c0*(1 - t) + c1*t
    return __dadd_rn(__dmul_rn(c0, __dsub_rn(1.0, t)), __dmul_rn(c1, t));
}

__global__ void k_trilinear(int64_t nz, int64_t nx, int64_t ny, const double* __restrict__ z, const double* __restrict__ x,
                            const double* __restrict__ y, const double* __restrict__ vals, int64_t n, const double* __restrict__ pos,
                            double* __restrict__ out, int* __restrict__ n_outside) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double zm = pos[3 * k], xm = pos[3 * k + 1], ym = pos[3 * k + 2];
    const int64_t iz = lower_corner(z, nz, zm), ix = lower_corner(x, nx, xm), iy = lower_corner(y, ny, ym);
    if (iz < 0 || iz > nz - 2 || ix < 0 || ix > nx - 2 || iy < 0 || iy > ny - 2 || !(zm == zm) || !(xm == xm) || !(ym == ym)) {
        out[k] = nan("");
        atomicAdd(n_outside, 1);
        return;
    }
    const double z0 = z[iz], z1 = z[iz + 1], x0 = x[ix], x1 = x[ix + 1], y0 = y[iy], y1 = y[iy + 1];
    const double x_d = __ddiv_rn(__dsub_rn(xm, x0), __dsub_rn(x1, x0));
    const double y_d = __ddiv_rn(__dsub_rn(ym, y0), __dsub_rn(y1, y0));
    const double z_d = __ddiv_rn(__dsub_rn(zm, z0), __dsub_rn(z1, z0));
    auto V = [&](int64_t a, int64_t b, int64_t c) { return vals[a + nz * (b + nx * c)]; };   // (nz, nx, ny) column-major
    const double c000 = V(iz, ix, iy), c010 = V(iz, ix, iy + 1), c100 = V(iz, ix + 1, iy), c110 = V(iz, ix + 1, iy + 1);
    const double c001 = V(iz + 1, ix, iy), c011 = V(iz + 1, ix, iy + 1), c101 = V(iz + 1, ix + 1, iy), c111 = V(iz + 1, ix + 1, iy + 1);
    const double c00 = lerp_rn(c000, c100, x_d), c01 = lerp_rn(c001, c101, x_d);
    const double c10 = lerp_rn(c010, c110, x_d), c11 = lerp_rn(c011, c111, x_d);
    const double c0 = lerp_rn(c00, c10, y_d), c1 = lerp_rn(c01, c11, y_d);
    out[k] = lerp_rn(c0, c1, z_d);
}

}  // namespace
}  // namespace vrt

using namespace vrt;

extern "C" int vrt_trilinear(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                             const double* vals, int64_t n, const double* positions, double* out) {
    if (nz < 2 || nx < 2 || ny < 2 || !z || !x || !y || !vals || n <= 0 || !positions || !out) {
        set_error("vrt_trilinear: bad arguments");
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("vrt_trilinear: no CUDA device (this library has no CPU path)");
        return VRT_E_CUDA;
    }
    const size_t vol = (size_t)nz * nx * ny;
    DevBuf<double> dz, dx, dy, dv, dp, dout;
    DevBuf<int> flag;
    auto dev = [&](const double* p, size_t cnt, DevBuf<double>& buf, const double** q) -> int {
        *q = p;
        if (!is_device_ptr(p)) {
            VRT_TRY(buf.alloc(cnt));
            VRT_CUDA(cudaMemcpy(buf.p, p, sizeof(double) * cnt, cudaMemcpyHostToDevice));
            *q = buf.p;
        }
        return VRT_OK;
    };
    const double *pz, *px, *py, *pv, *pp;
    VRT_TRY(dev(z, nz, dz, &pz)); VRT_TRY(dev(x, nx, dx, &px)); VRT_TRY(dev(y, ny, dy, &py));
    VRT_TRY(dev(vals, vol, dv, &pv)); VRT_TRY(dev(positions, (size_t)3 * n, dp, &pp));
    double* po = out;
    const bool dev_out = is_device_ptr(out);
    if (!dev_out) {
        VRT_TRY(dout.alloc(n));
        po = dout.p;
    }
    VRT_TRY(flag.alloc(1));
    VRT_CUDA(cudaMemset(flag.p, 0, sizeof(int)));
    k_trilinear<<<(unsigned)((n + 255) / 256), 256>>>(nz, nx, ny, pz, px, py, pv, n, pp, po, flag.p);
    VRT_CUDA(cudaGetLastError());
    int outside = 0;
    VRT_CUDA(cudaMemcpy(&outside, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (!dev_out) VRT_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaDeviceSynchronize());
    if (outside > 0) {
        set_error("vrt_trilinear: %d sites lie outside the atmosphere axes (the reference throws a BoundsError there); their values are NaN", outside);
        return VRT_E_INVALID;
    }
    return VRT_OK;
}
