// initialise.cu — site initialisation: trilinear interpolation of an atmosphere field to the Voronoi sites
// (first half; reference: trilinear src/functions.jl:207-248, initialise src/voronoi_utils.jl:687-708, which
// calls it for temperature, electron density, hydrogen density and the three velocity components).
//
// One thread per site: three binary searches with searchsortedfirst semantics (first index whose value is >= the
// coordinate, minus one = lower corner), eight loads, and the seven lerps in the order the reference writes them, each
// product and sum rounded separately (no fma contraction), so the result equals a plain IEEE evaluation of the Julia
// expressions bit for bit.  A site outside the axes (where Julia throws a BoundsError) yields NaN and an error return.
#include <algorithm>
#include <math.h>
#include <cub/cub.cuh>
#include "sampling.cuh"
#include "vrt_internal.h"

namespace vrt {
namespace {

__global__ void k_trilinear(TriGrid T, int64_t n, const double* __restrict__ pos, double* __restrict__ out, int* __restrict__ n_outside) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    double v;
    if (!trilinear_at(T, pos[3 * k], pos[3 * k + 1], pos[3 * k + 2], &v)) {
        v = nan("");
        atomicAdd(n_outside, 1);
    }
    out[k] = v;
}

__global__ void k_nearest_corner(TriGrid T, int64_t n, const double* __restrict__ pos, double* __restrict__ out, int* __restrict__ n_outside) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    double v;
    if (!nearest_corner_at(T, pos[3 * k], pos[3 * k + 1], pos[3 * k + 2], &v)) {
        v = nan("");
        atomicAdd(n_outside, 1);
    }
    out[k] = v;
}

__global__ void k_rejection_sampling(TriGrid T, uint64_t seed, int64_t n, double q_min, double dq, int64_t max_trials,
                                     double* __restrict__ pos, unsigned long long* __restrict__ trials, int* __restrict__ n_failed) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = {nan(""), nan(""), nan("")};
    const int64_t t = rejection_site(T, seed, i, q_min, dq, max_trials, p);
    pos[3 * i] = p[0]; pos[3 * i + 1] = p[1]; pos[3 * i + 2] = p[2];
    if (t < 0) atomicAdd(n_failed, 1);
    else atomicAdd(trials, (unsigned long long)t);
}

// device copies of the axes and the field
struct TriDev {
    DevBuf<double> dz, dx, dy, dv;
    TriGrid T;
    int init(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y, const double* vals) {
        auto dev = [&](const double* p, size_t cnt, DevBuf<double>& buf, const double** q) -> int {
            *q = p;
            if (!is_device_ptr(p)) {
                VRT_TRY(buf.alloc(cnt));
                VRT_CUDA(cudaMemcpy(buf.p, p, sizeof(double) * cnt, cudaMemcpyHostToDevice));
                *q = buf.p;
            }
            return VRT_OK;
        };
        T.nz = nz; T.nx = nx; T.ny = ny;
        VRT_TRY(dev(z, nz, dz, &T.z)); VRT_TRY(dev(x, nx, dx, &T.x)); VRT_TRY(dev(y, ny, dy, &T.y));
        VRT_TRY(dev(vals, (size_t)nz * nx * ny, dv, &T.vals));
        return VRT_OK;
    }
};

}  // namespace
}  // namespace vrt

using namespace vrt;

static int interpolate_sites(const char* who, int mode, int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                             const double* vals, int64_t n, const double* positions, double* out) {
    if (nz < 2 || nx < 2 || ny < 2 || !z || !x || !y || !vals || n <= 0 || !positions || !out) {
        set_error("%s: bad arguments", who);
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("%s: no CUDA device (this library has no CPU path)", who);
        return VRT_E_CUDA;
    }
    TriDev D;
    VRT_TRY(D.init(nz, nx, ny, z, x, y, vals));
    DevBuf<double> dp, dout;
    DevBuf<int> flag;
    const double* pp = positions;
    if (!is_device_ptr(positions)) {
        VRT_TRY(dp.alloc((size_t)3 * n));
        VRT_CUDA(cudaMemcpy(dp.p, positions, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
        pp = dp.p;
    }
    double* po = out;
    const bool dev_out = is_device_ptr(out);
    if (!dev_out) {
        VRT_TRY(dout.alloc(n));
        po = dout.p;
    }
    VRT_TRY(flag.alloc(1));
    VRT_CUDA(cudaMemset(flag.p, 0, sizeof(int)));
    if (mode == 0) k_trilinear<<<(unsigned)((n + 255) / 256), 256>>>(D.T, n, pp, po, flag.p);
    else k_nearest_corner<<<(unsigned)((n + 255) / 256), 256>>>(D.T, n, pp, po, flag.p);
    VRT_CUDA(cudaGetLastError());
    int outside = 0;
    VRT_CUDA(cudaMemcpy(&outside, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (!dev_out) VRT_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaDeviceSynchronize());
    if (outside > 0) {
        set_error("%s: %d sites lie outside the atmosphere axes (the reference throws a BoundsError there); their values are NaN", who, outside);
        return VRT_E_INVALID;
    }
    return VRT_OK;
}

extern "C" int vrt_trilinear(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                             const double* vals, int64_t n, const double* positions, double* out) {
    return interpolate_sites("vrt_trilinear", 0, nz, nx, ny, z, x, y, vals, n, positions, out);
}

extern "C" int vrt_nearest_corner(int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x, const double* y,
                                  const double* vals, int64_t n, const double* positions, double* out) {
    return interpolate_sites("vrt_nearest_corner", 1, nz, nx, ny, z, x, y, vals, n, positions, out);
}

extern "C" int vrt_rejection_sampling(int64_t n_sites, int64_t nz, int64_t nx, int64_t ny, const double* z, const double* x,
                                      const double* y, const double* quantity, uint64_t seed, double* positions, double* mean_trials) {
    if (n_sites <= 0 || nz < 2 || nx < 2 || ny < 2 || !z || !x || !y || !quantity || !positions) {
        set_error("vrt_rejection_sampling: bad arguments");
        return VRT_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("vrt_rejection_sampling: no CUDA device (this library has no CPU path)");
        return VRT_E_CUDA;
    }
    TriDev D;
    VRT_TRY(D.init(nz, nx, ny, z, x, y, quantity));
    // q_min, q_max over the grid (functions.jl:90-93)
    const size_t vol = (size_t)nz * nx * ny;
    DevBuf<double> mm;
    VRT_TRY(mm.alloc(2));
    {
        size_t tb = 0, tb2 = 0;
        VRT_CUDA(cub::DeviceReduce::Min(nullptr, tb, D.T.vals, mm.p, (int)vol));
        VRT_CUDA(cub::DeviceReduce::Max(nullptr, tb2, D.T.vals, mm.p + 1, (int)vol));
        DevBuf<char> tmp;
        VRT_TRY(tmp.alloc(std::max(tb, tb2)));
        VRT_CUDA(cub::DeviceReduce::Min(tmp.p, tb, D.T.vals, mm.p, (int)vol));
        VRT_CUDA(cub::DeviceReduce::Max(tmp.p, tb2, D.T.vals, mm.p + 1, (int)vol));
        VRT_CUDA(cudaDeviceSynchronize());
    }
    double h[2];
    VRT_CUDA(cudaMemcpy(h, mm.p, sizeof(h), cudaMemcpyDeviceToHost));
    const double q_min = h[0], dq = h[1] - h[0];
    if (!(dq > 0)) {
        set_error("vrt_rejection_sampling: the quantity is constant or not finite (min %g, max %g)", h[0], h[1]);
        return VRT_E_INVALID;
    }
    DevBuf<double> dpos;
    DevBuf<unsigned long long> trials;
    DevBuf<int> failed;
    double* pp = positions;
    const bool dev_out = is_device_ptr(positions);
    if (!dev_out) {
        VRT_TRY(dpos.alloc((size_t)3 * n_sites));
        pp = dpos.p;
    }
    VRT_TRY(trials.alloc(1)); VRT_TRY(failed.alloc(1));
    VRT_CUDA(cudaMemset(trials.p, 0, sizeof(unsigned long long)));
    VRT_CUDA(cudaMemset(failed.p, 0, sizeof(int)));
    const int64_t max_trials = 1 << 20;
    k_rejection_sampling<<<(unsigned)((n_sites + 127) / 128), 128>>>(D.T, seed, n_sites, q_min, dq, max_trials, pp, trials.p, failed.p);
    VRT_CUDA(cudaGetLastError());
    unsigned long long ht = 0;
    int hf = 0;
    VRT_CUDA(cudaMemcpy(&ht, trials.p, sizeof(ht), cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaMemcpy(&hf, failed.p, sizeof(hf), cudaMemcpyDeviceToHost));
    if (!dev_out) VRT_CUDA(cudaMemcpy(positions, dpos.p, sizeof(double) * 3 * (size_t)n_sites, cudaMemcpyDeviceToHost));
    VRT_CUDA(cudaDeviceSynchronize());
    if (mean_trials) *mean_trials = (double)ht / (double)(n_sites - hf > 0 ? n_sites - hf : 1);
    if (hf > 0) {
        set_error("vrt_rejection_sampling: %d sites found no accepted candidate in %lld trials", hf, (long long)max_trials);
        return VRT_E_STATE;
    }
    return VRT_OK;
}
