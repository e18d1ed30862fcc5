// common.cu — error reporting, pointer classification and host<->device copies for libvrt.so.
#include <stdarg.h>
#include <string.h>
#include "vrt_internal.h"

namespace vrt {

static thread_local char g_err[1024] = "";
thread_local SweepStats g_last_stats;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? VRT_E_NOMEM : VRT_E_CUDA;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int copy_in(void* dst_dev, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return VRT_OK;
    VRT_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    return VRT_OK;
}

int copy_out(void* dst, const void* src_dev, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return VRT_OK;
    VRT_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return VRT_OK;
}

}  // namespace vrt

extern "C" {

int vrt_abi_version(void) { return VRT_ABI_VERSION; }

const char* vrt_last_error(void) { return vrt::g_err; }

int vrt_device_count(int32_t* count) {
    if (!count) return VRT_E_INVALID;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *count = 0;
        return vrt::cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    *count = c;
    return VRT_OK;
}

int vrt_set_device(int32_t device) {
    VRT_CUDA(cudaSetDevice(device));
    return VRT_OK;
}

int vrt_last_stats(double out[8]) {
    if (!out) return VRT_E_INVALID;
    memset(out, 0, 8 * sizeof(double));
    out[0] = vrt::g_last_stats.kernels;
    out[1] = vrt::g_last_stats.visits;
    out[2] = vrt::g_last_stats.steps;
    out[3] = vrt::g_last_stats.sweep_ms;
    return VRT_OK;
}

/* voronoi_utils.jl:42-70 — host-side text parsing of the voro++ output (output_sites.cc:49). */
int vrt_read_neighbours(const char* fname, int64_t n, int64_t* nbr, int64_t ld, int64_t* ld_needed) {
    if (!fname || n <= 0) {
        vrt::set_error("vrt_read_neighbours: bad arguments");
        return VRT_E_INVALID;
    }
    FILE* f = fopen(fname, "rb");
    if (!f) {
        vrt::set_error("vrt_read_neighbours: cannot open %s", fname);
        return VRT_E_IO;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) {
        fclose(f);
        vrt::set_error("vrt_read_neighbours: short read on %s", fname);
        return VRT_E_IO;
    }
    fclose(f);
    buf[(size_t)sz] = 0;
    if (nbr) memset(nbr, 0, sizeof(int64_t) * (size_t)n * (size_t)ld);
    int64_t maxn = 0, lines = 0;
    const char* p = buf.data();
    const char* end = p + sz;
    // hand-rolled integer scanner: one pass, no per-token allocation (the reference re-splits the line per token)
    while (p < end) {
        // parse one line
        int64_t vals = 0, id = 0, cnt = 0;
        while (p < end && *p != '\n') {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
            if (p >= end || *p == '\n') break;
            bool neg = false;
            if (*p == '-') { neg = true; p++; }
            else if (*p == '+') p++;
            if (p >= end || *p < '0' || *p > '9') {
                vrt::set_error("vrt_read_neighbours: unexpected character in %s (line %lld)", fname, (long long)lines + 1);
                return VRT_E_IO;
            }
            int64_t v = 0;
            while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
            if (neg) v = -v;
            if (vals == 0) id = v;
            else {
                cnt++;
                if (nbr) {
                    if (id < 1 || id > n) {
                        vrt::set_error("vrt_read_neighbours: site id %lld out of range", (long long)id);
                        return VRT_E_IO;
                    }
                    if (cnt < ld) nbr[(id - 1) + n * cnt] = v;
                }
            }
            vals++;
        }
        if (p < end) p++;  // newline
        if (vals > 0) {
            lines++;
            if (nbr && id >= 1 && id <= n) nbr[id - 1] = cnt;
            if (cnt > maxn) maxn = cnt;
        }
    }
    if (ld_needed) *ld_needed = maxn + 1;
    if (nbr && ld < maxn + 1) {
        vrt::set_error("vrt_read_neighbours: ld=%lld too small, need %lld", (long long)ld, (long long)maxn + 1);
        return VRT_E_INVALID;
    }
    return VRT_OK;
}

}  // extern "C"
