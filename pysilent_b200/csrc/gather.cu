// Multi-GPU feature-point gather over NCCL behind the C ABI. The reference is single-device
// (recognition_testing.py:64): frames shard across ranks with no data-path exchange, and the ONE exchange of the path is
// this gather -- every rank contributes a fixed-size packed block (silent_pack_points: rows + count row) to one
// ncclAllGather on the caller's (side) stream.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded into the process -- PyTorch's -- or the
// system one), with the handful of entry points declared here from NCCL's public header; the library itself has no link
// dependency on NCCL, so single-GPU users never need it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace silent {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;   // NCCL_UNIQUE_ID_BYTES
typedef int ncclResult_t;                              // ncclSuccess = 0, ncclInProgress = 7
constexpr int kNcclInt64 = 4;                          // ncclDataType_t: ncclInt64

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    bool ok = false;
};

static const NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already uses (PyTorch's)
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return;
#define SILENT_NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, name))
        SILENT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        SILENT_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        SILENT_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        SILENT_NCCL_SYM(CommAbort, "ncclCommAbort");
        SILENT_NCCL_SYM(CommGetAsyncError, "ncclCommGetAsyncError");
        SILENT_NCCL_SYM(AllGather, "ncclAllGather");
        SILENT_NCCL_SYM(GetErrorString, "ncclGetErrorString");
        SILENT_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef SILENT_NCCL_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.CommGetAsyncError && api.AllGather &&
                 api.GetErrorString;
    });
    return api.ok ? &api : nullptr;
}

static int nccl_fail(const NcclApi *api, ncclResult_t r, const char *what)
{
    return fail(SILENT_E_NCCL, "%s failed: %s", what, api->GetErrorString(r));
}

}  // namespace silent

using namespace silent;

struct silent_comm {
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0, device = 0;
};

extern "C" {

int silent_comm_unique_id(void *id_out)
{
    if (!id_out) return fail(SILENT_E_INVAL, "silent_comm_unique_id: null argument");
    const NcclApi *api = nccl_api();
    if (!api) return fail(SILENT_E_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "not found");
    ncclUniqueId id;
    const ncclResult_t r = api->GetUniqueId(&id);
    if (r != 0) return nccl_fail(api, r, "ncclGetUniqueId");
    std::memcpy(id_out, &id, sizeof(id));
    return SILENT_OK;
}

int silent_comm_create(const void *id, int nranks, int rank, silent_comm **out_comm)
{
    if (!id || !out_comm) return fail(SILENT_E_INVAL, "silent_comm_create: null argument");
    if (nranks <= 0 || rank < 0 || rank >= nranks) return fail(SILENT_E_INVAL, "silent_comm_create: rank %d of %d", rank, nranks);
    const NcclApi *api = nccl_api();
    if (!api) return fail(SILENT_E_NCCL, "libnccl.so.2 could not be loaded");
    silent_comm *c = new silent_comm;
    c->nranks = nranks, c->rank = rank;
    SILENT_CUDA(cudaGetDevice(&c->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    const ncclResult_t r = api->CommInitRank(&c->comm, nranks, uid, rank);   // collective: every rank calls it
    if (r != 0) {
        delete c;
        return nccl_fail(api, r, "ncclCommInitRank");
    }
    *out_comm = c;
    return SILENT_OK;
}

void silent_comm_destroy(silent_comm *comm)
{
    if (!comm) return;
    const NcclApi *api = nccl_api();
    if (api && comm->comm) api->CommDestroy(comm->comm);
    delete comm;
}

int silent_comm_check(silent_comm *comm)
{
    if (!comm) return fail(SILENT_E_INVAL, "silent_comm_check: null communicator");
    const NcclApi *api = nccl_api();
    if (!api) return fail(SILENT_E_NCCL, "libnccl.so.2 could not be loaded");
    ncclResult_t async = 0;
    const ncclResult_t r = api->CommGetAsyncError(comm->comm, &async);
    if (r != 0) return nccl_fail(api, r, "ncclCommGetAsyncError");
    if (async != 0 && async != 7) return nccl_fail(api, async, "an earlier point gather (asynchronous NCCL error)");
    return SILENT_OK;
}

int silent_gather_points(silent_comm *comm, const int64_t *packed_send_dev, int64_t rows, int64_t *packed_recv_dev,
                         silent_stream stream)
{
    if (!comm || !packed_send_dev || !packed_recv_dev) return fail(SILENT_E_INVAL, "silent_gather_points: null argument");
    if (rows <= 0) return fail(SILENT_E_INVAL, "silent_gather_points: rows must be positive");
    int rc = silent_comm_check(comm);   // a failure of an earlier collective surfaces here, before a new one is queued
    if (rc != SILENT_OK) return rc;
    const NcclApi *api = nccl_api();
    const ncclResult_t r = api->AllGather(packed_send_dev, packed_recv_dev, (size_t)rows * 4, kNcclInt64, comm->comm,
                                          (cudaStream_t)stream);
    if (r != 0) return nccl_fail(api, r, "ncclAllGather");
    return SILENT_OK;
}

}  // extern "C"
