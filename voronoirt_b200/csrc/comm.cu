// comm.cu — the collectives of the multi-GPU Λ-iteration, issued by the library itself.
//
// For fixed S and populations every (direction, wavelength) formal solution is independent (reference
// src/lambda_iteration.jl:84-110); the only exchange steps are the sum of the directions' contributions to J (:102,:107), the
// wavelength integrals of the rates (src/rates.jl:226-364) and the criterion's maximum (:325-349).  One process per GPU; the
// host (Julia, Python, C) only ferries a 128-byte NCCL unique id between the processes with whatever it has (MPI, sockets,
// torch.distributed) and hands it to vrt_solver_comm_init — it needs no NCCL binding of its own.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, types from <nccl.h>): libvrt.so itself has no link-time dependency on it,
// so a single-GPU host does not need NCCL installed, and inside a process that already loaded NCCL (PyTorch) the same copy is used.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include "vrt_internal.h"

namespace vrt {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("VRT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define VRT_SYM(field, name) *(void**)(&api.field) = dlsym(api.handle, name)
        VRT_SYM(GetUniqueId, "ncclGetUniqueId");
        VRT_SYM(CommInitRank, "ncclCommInitRank");
        VRT_SYM(CommDestroy, "ncclCommDestroy");
        VRT_SYM(GetErrorString, "ncclGetErrorString");
        VRT_SYM(AllReduce, "ncclAllReduce");
        VRT_SYM(ReduceScatter, "ncclReduceScatter");
        VRT_SYM(AllGather, "ncclAllGather");
        VRT_SYM(GetVersion, "ncclGetVersion");
#undef VRT_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GetErrorString && api.AllReduce && api.ReduceScatter &&
                 api.AllGather;
    });
    return &api;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    NcclApi* a = nccl_api();
    set_error("NCCL error %d (%s) in %s", (int)r, a->GetErrorString ? a->GetErrorString(r) : "?", what);
    return VRT_E_CUDA;
}
#define VRT_NCCL(call)                                   \
    do {                                                 \
        ncclResult_t _r = (call);                        \
        if (_r != ncclSuccess) return nccl_fail(_r, #call); \
    } while (0)

struct Comm {
    ncclComm_t dir = nullptr, lam = nullptr;   // over the direction shards (= cell shards) / over the wavelength shards
    int dir_rank = 0, dir_size = 1, lam_rank = 0, lam_size = 1;
    cudaStream_t stream = nullptr;             // the collectives' own stream (non-blocking): they overlap the compute stream
    ~Comm() {
        NcclApi* a = nccl_api();
        if (dir && a->ok) a->CommDestroy(dir);
        if (lam && a->ok) a->CommDestroy(lam);
        if (stream) cudaStreamDestroy(stream);
    }
};

void comm_free(void* c) { delete static_cast<Comm*>(c); }

int comm_create(const char* dir_id, int dir_rank, int dir_size, const char* lam_id, int lam_rank, int lam_size, void** out) {
    NcclApi* a = nccl_api();
    if (!a->ok) {
        set_error("NCCL is not available (libnccl.so.2 could not be loaded; set VRT_NCCL_LIB)");
        return VRT_E_STATE;
    }
    if ((dir_size > 1 && (!dir_id || dir_rank < 0 || dir_rank >= dir_size)) || (lam_size > 1 && (!lam_id || lam_rank < 0 || lam_rank >= lam_size))) {
        set_error("vrt_solver_comm_init: bad communicator description");
        return VRT_E_INVALID;
    }
    Comm* c = new Comm();
    struct Guard { Comm* c; ~Guard() { delete c; } } guard{c};
    c->dir_rank = dir_rank; c->dir_size = dir_size > 1 ? dir_size : 1;
    c->lam_rank = lam_rank; c->lam_size = lam_size > 1 ? lam_size : 1;
    VRT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (dir_size > 1) {
        ncclUniqueId id;
        memcpy(id.internal, dir_id, NCCL_UNIQUE_ID_BYTES);
        VRT_NCCL(a->CommInitRank(&c->dir, dir_size, id, dir_rank));
    }
    if (lam_size > 1) {
        ncclUniqueId id;
        memcpy(id.internal, lam_id, NCCL_UNIQUE_ID_BYTES);
        VRT_NCCL(a->CommInitRank(&c->lam, lam_size, id, lam_rank));
    }
    guard.c = nullptr;
    *out = c;
    return VRT_OK;
}

cudaStream_t comm_stream(void* comm) { return static_cast<Comm*>(comm)->stream; }

// The five exchange steps of vrt_allreduce_fn (include/vrt.h), enqueued on `st`:
//   op 0 sum over the wavelength group, 1 max over all processes, 2 sum over the direction group,
//   3 reduce-scatter / 4 all-gather of `count` doubles in dir_size equal slices over the direction group (in place).
int comm_op(void* comm, double* buf, int64_t count, int op, cudaStream_t st) {
    Comm* c = static_cast<Comm*>(comm);
    NcclApi* a = nccl_api();
    switch (op) {
        case 0:
            if (c->lam) VRT_NCCL(a->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, c->lam, st));
            break;
        case 1:
            if (c->dir) VRT_NCCL(a->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclMax, c->dir, st));
            if (c->lam) VRT_NCCL(a->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclMax, c->lam, st));
            break;
        case 2:
            if (c->dir) VRT_NCCL(a->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, c->dir, st));
            break;
        case 3:
        case 4: {
            if (!c->dir) break;
            if (count % c->dir_size) {
                set_error("collective: %lld doubles do not split into %d slices", (long long)count, c->dir_size);
                return VRT_E_INVALID;
            }
            const size_t per = (size_t)(count / c->dir_size);
            double* mine = buf + per * (size_t)c->dir_rank;
            if (op == 3) VRT_NCCL(a->ReduceScatter(buf, mine, per, ncclDouble, ncclSum, c->dir, st));
            else VRT_NCCL(a->AllGather(mine, buf, per, ncclDouble, c->dir, st));
            break;
        }
        default:
            set_error("collective: unknown op %d", op);
            return VRT_E_INVALID;
    }
    return VRT_OK;
}

}  // namespace vrt

extern "C" {

int vrt_nccl_available(void) { return vrt::nccl_api()->ok ? 1 : 0; }

int vrt_nccl_version(int32_t* version) {
    vrt::NcclApi* a = vrt::nccl_api();
    if (!version) return VRT_E_INVALID;
    *version = 0;
    if (!a->ok || !a->GetVersion) {
        vrt::set_error("NCCL is not available");
        return VRT_E_STATE;
    }
    int v = 0;
    a->GetVersion(&v);
    *version = v;
    return VRT_OK;
}

int vrt_nccl_unique_id(char id[128]) {
    vrt::NcclApi* a = vrt::nccl_api();
    if (!id) return VRT_E_INVALID;
    if (!a->ok) {
        vrt::set_error("NCCL is not available (libnccl.so.2 could not be loaded; set VRT_NCCL_LIB)");
        return VRT_E_STATE;
    }
    ncclUniqueId u;
    ncclResult_t r = a->GetUniqueId(&u);
    if (r != ncclSuccess) return vrt::nccl_fail(r, "ncclGetUniqueId");
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return VRT_OK;
}

}  // extern "C"
