// Native NCCL communicator for callers that do not bring their own all-reduce (SURVEY.md §8b:
// "hmmb_comm_init(rank, world, nccl_id) for multi-GPU").  hmmb_comm_allreduce has the signature of
// hmmb_allreduce_fn, so it is passed to hmmb_bw_set_dist / hmmb_lbg_fit like any other hook; the Python
// shim keeps using torch.distributed (dist.make_allreduce) unless asked for allreduce="native".
//
// libnccl is opened with dlopen at the first call instead of being linked: a process that never shards
// does not need it, and inside a PyTorch process dlopen("libnccl.so.2") resolves to the copy torch has
// already mapped (same SONAME), so one process never holds two NCCL versions.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace hmmb {

namespace {

struct NcclUniqueId { char internal[128]; };  // NCCL_UNIQUE_ID_BYTES
using ncclComm_t = void *;
constexpr int NCCL_FLOAT64 = 8;  // ncclDataType_t::ncclFloat64
constexpr int NCCL_SUM = 0;      // ncclRedOp_t::ncclSum

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*CommGetAsyncError)(ncclComm_t, int *) = nullptr;  // optional
};

NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;

int nccl_load() {
    if (g_nccl.AllReduce) return HMMB_OK;
    const char *names[] = {getenv("HMMB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    }
    if (!h) {
        set_error("hmmb_comm: cannot open libnccl.so.2 (%s); set HMMB_NCCL_LIB", dlerror());
        return HMMB_ERR_UNSUPPORTED;
    }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.CommGetAsyncError = reinterpret_cast<decltype(a.CommGetAsyncError)>(dlsym(h, "ncclCommGetAsyncError"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GetErrorString) {
        set_error("hmmb_comm: libnccl lacks an expected symbol");
        dlclose(h);
        return HMMB_ERR_UNSUPPORTED;
    }
    g_nccl = a;
    return HMMB_OK;
}

int nccl_fail(int rc, const char *what) {
    set_error("NCCL error %d (%s) in %s", rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?", what);
    return HMMB_ERR_CUDA;
}

}  // namespace

}  // namespace hmmb

using namespace hmmb;

extern "C" int hmmb_comm_unique_id(void *id_out, int id_bytes) {
    if (!id_out || id_bytes < (int)sizeof(NcclUniqueId)) {
        set_error("hmmb_comm_unique_id: need a buffer of %d bytes", (int)sizeof(NcclUniqueId));
        return HMMB_ERR_ARG;
    }
    HMMB_TRY(nccl_load());
    NcclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return nccl_fail(rc, "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof(id));
    return HMMB_OK;
}

extern "C" int hmmb_comm_init(int rank, int world, const void *nccl_id) {
    HMMB_TRY(require_init());  // the communicator belongs to the context's device
    if (world < 1 || rank < 0 || rank >= world || !nccl_id) {
        set_error("hmmb_comm_init: bad arguments (rank=%d world=%d)", rank, world);
        return HMMB_ERR_ARG;
    }
    if (g_comm) { set_error("hmmb_comm_init: a communicator already exists (hmmb_comm_destroy first)"); return HMMB_ERR_ARG; }
    HMMB_TRY(nccl_load());
    NcclUniqueId id;
    memcpy(&id, nccl_id, sizeof(id));
    HMMB_CUDA(cudaSetDevice(ctx().device));
    const int rc = g_nccl.CommInitRank(&g_comm, world, id, rank);
    if (rc != 0) { g_comm = nullptr; return nccl_fail(rc, "ncclCommInitRank"); }
    g_rank = rank;
    g_world = world;
    return HMMB_OK;
}

// hmmb_allreduce_fn: in-place fp64 sum over the ranks, ordered on hmmb_get_stream().  (NULL, 0) — the "join"
// call of hmmb_bw_set_overlap — is a no-op because nothing runs on a side stream here.
extern "C" int hmmb_comm_allreduce(void *dev_buf, int64_t n_doubles, void *user) {
    (void)user;
    if (!g_comm) { set_error("hmmb_comm_allreduce: no communicator (hmmb_comm_init)"); return HMMB_ERR_ARG; }
    if (!dev_buf || n_doubles <= 0) return HMMB_OK;
    const int rc = g_nccl.AllReduce(dev_buf, dev_buf, (size_t)n_doubles, NCCL_FLOAT64, NCCL_SUM, g_comm, ctx().stream);
    if (rc != 0) return nccl_fail(rc, "ncclAllReduce");
    // failure detection: an asynchronous error of the communicator — a peer that died, a
    // transport fault — surfaces here, at the next collective, instead of as a hang at the next synchronisation
    if (g_nccl.CommGetAsyncError) {
        int async_rc = 0;
        const int q = g_nccl.CommGetAsyncError(g_comm, &async_rc);
        if (q != 0) return nccl_fail(q, "ncclCommGetAsyncError");
        if (async_rc != 0) return nccl_fail(async_rc, "ncclAllReduce (asynchronous error of the communicator)");
    }
    return HMMB_OK;
}

extern "C" int hmmb_comm_rank(int *rank, int *world) {
    if (rank) *rank = g_rank;
    if (world) *world = g_comm ? g_world : 1;
    return HMMB_OK;
}

extern "C" int hmmb_comm_destroy(void) {
    if (!g_comm) return HMMB_OK;
    cudaStreamSynchronize(ctx().stream);
    const int rc = g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    g_rank = 0;
    g_world = 1;
    if (rc != 0) return nccl_fail(rc, "ncclCommDestroy");
    return HMMB_OK;
}
