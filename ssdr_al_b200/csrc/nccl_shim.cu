// nccl_shim.cu -- NCCL reached through dlopen, so single-GPU use of libssdr_b200.so has no NCCL dependency and a
// process that already loaded torch's bundled NCCL shares that copy (same SONAME).  Only the row-sharded FPS uses
// a collective on this hot path: one 8-byte max all-reduce per pick.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace ssdr {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};

static NcclApi g_api;
static bool g_loaded = false;

static int load_nccl() {
    if (g_loaded) return SSDR_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(SSDR_ERR_UNSUPPORTED, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SSDR_SYM(field, name)                                                            \
    *(void**)(&g_api.field) = dlsym(h, name);                                            \
    if (!g_api.field) return set_error(SSDR_ERR_UNSUPPORTED, "libnccl lacks %s", name)
    SSDR_SYM(GetUniqueId, "ncclGetUniqueId");
    SSDR_SYM(CommInitRank, "ncclCommInitRank");
    SSDR_SYM(CommDestroy, "ncclCommDestroy");
    SSDR_SYM(AllReduce, "ncclAllReduce");
    SSDR_SYM(GetErrorString, "ncclGetErrorString");
#undef SSDR_SYM
    g_loaded = true;
    return SSDR_OK;
}

#define SSDR_CHECK_NCCL(expr)                                                                               \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) return set_error(SSDR_ERR_CUDA, "%s failed: %s", #expr, g_api.GetErrorString(_r)); \
    } while (0)

// used by selection.cu: in-place max all-reduce of `count` u64 words on `stream`
int nccl_allreduce_max_u64(void* comm, unsigned long long* buf, size_t count, cudaStream_t stream) {
    SSDR_TRY(load_nccl());
    SSDR_CHECK_NCCL(g_api.AllReduce(buf, buf, count, ncclUint64, ncclMax, (ncclComm_t)comm, stream));
    return SSDR_OK;
}

}  // namespace ssdr

using namespace ssdr;

extern "C" {
int ssdr_nccl_unique_id(void* id128) {
    SSDR_REQUIRE(id128, SSDR_ERR_INVALID, "id128 is NULL");
    SSDR_TRY(load_nccl());
    ncclUniqueId id;
    SSDR_CHECK_NCCL(g_api.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return SSDR_OK;
}
int ssdr_nccl_comm_init(void** comm, int nranks, const void* id128, int rank) {
    SSDR_REQUIRE(comm && id128, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_TRY(load_nccl());
    Ctx* c;
    SSDR_TRY(get_ctx(&c));  // binds the calling thread's current device
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t cm = nullptr;
    SSDR_CHECK_NCCL(g_api.CommInitRank(&cm, nranks, id, rank));
    *comm = cm;
    return SSDR_OK;
}
int ssdr_nccl_comm_destroy(void* comm) {
    if (!comm) return SSDR_OK;
    SSDR_TRY(load_nccl());
    SSDR_CHECK_NCCL(g_api.CommDestroy((ncclComm_t)comm));
    return SSDR_OK;
}
}
