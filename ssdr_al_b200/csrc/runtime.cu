// runtime.cu -- status reporting, per-thread contexts, device management for libssdr_b200.so.
#include <stdarg.h>
#include <unistd.h>

#include "common.cuh"

namespace ssdr {

static thread_local char g_err[512] = "";
static pid_t g_init_pid = 0;

int set_error(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

static thread_local unsigned long long g_ws_generation = 0;
unsigned long long ws_generation() { return g_ws_generation; }

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return SSDR_OK;
    ++g_ws_generation;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = align_up(bytes + bytes / 8, 1 << 20);  // a little head-room so repeated calls stop reallocating
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        return set_error(SSDR_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = want;
    return SSDR_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
Ctx::~Ctx() {
    if (device < 0 || getpid() != g_init_pid) return;  // never initialised, or a fork()ed child that owns no CUDA state
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();  // the runtime is already shutting down (process exit): nothing left to release
        return;
    }
    if (stream) cudaStreamSynchronize(stream);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    for (int k = 1; k < MAX_BRANCH; ++k) {
        if (branch_stream[k]) cudaStreamSynchronize(branch_stream[k]);
        if (bank[k]) {
            for (int i = 0; i < WS_SLOTS; ++i) bank[k][i].release();
            delete[] bank[k];
        }
        if (ev_branch[k]) cudaEventDestroy(ev_branch[k]);
        if (branch_stream[k]) cudaStreamDestroy(branch_stream[k]);
        bank[k] = nullptr;
        ev_branch[k] = nullptr;
        branch_stream[k] = nullptr;
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (int i = 0; i < WS_SLOTS; ++i) ws0[i].release();
    ws = ws0;
    if (ev) cudaEventDestroy(ev);
    if (ev_main) cudaEventDestroy(ev_main);
    if (ev_async) cudaEventDestroy(ev_async);
    for (int i = 0; i < 8; ++i)
        if (ev_chunk[i]) cudaEventDestroy(ev_chunk[i]);
    for (int i = 0; i < 5; ++i)
        if (tev[i]) cudaEventDestroy(tev[i]);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (cur >= 0) cudaSetDevice(cur);
    cudaGetLastError();
    device = -1;
}

enum { MAX_DEV = 16 };
static thread_local Ctx g_ctx[MAX_DEV];

int get_ctx(Ctx** out) {
    pid_t me = getpid();
    if (g_init_pid == 0) g_init_pid = me;
    else if (g_init_pid != me)
        return set_error(SSDR_ERR_FORK,
                         "libssdr_b200 was initialised in process %d and cannot be used in its fork()ed child %d; "
                         "use the 'spawn' or 'forkserver' multiprocessing start method (or num_workers=0)",
                         (int)g_init_pid, (int)me);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(SSDR_ERR_CUDA, "no usable CUDA device: %s (this library has no CPU fallback)",
                         cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= MAX_DEV) return set_error(SSDR_ERR_CUDA, "device ordinal %d out of range", dev);
    Ctx* c = &g_ctx[dev];
    if (c->device != dev) {
        cudaDeviceProp prop;
        SSDR_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major < 10)
            return set_error(SSDR_ERR_CUDA, "device %d (%s, sm_%d%d) is not a Blackwell sm_100 GPU", dev, prop.name,
                             prop.major, prop.minor);
        SSDR_CHECK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        SSDR_CHECK_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        SSDR_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming));
        SSDR_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        for (int i = 0; i < 8; ++i) SSDR_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
        for (int i = 0; i < 5; ++i) SSDR_CHECK_CUDA(cudaEventCreate(&c->tev[i]));
        SSDR_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_async, cudaEventDisableTiming));
        cudaMemPool_t pool;  // keep freed stream-ordered blocks cached instead of returning them to the driver
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        c->sm_count = prop.multiProcessorCount;
        c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
        c->device = dev;
    }
    if (c->async_pending) {
        if (cudaEventQuery(c->ev_async) == cudaSuccess) c->async_pending = false;
        else if (c->async_stream != c->stream) SSDR_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_async, 0));
        cudaGetLastError();
    }
    *out = c;
    return SSDR_OK;
}

int ctx_order(Ctx* c, cudaStream_t s) {
    if (c->async_pending && s != c->async_stream) SSDR_CHECK_CUDA(cudaStreamWaitEvent(s, c->ev_async, 0));
    return SSDR_OK;
}

int ctx_mark_async(Ctx* c, cudaStream_t s) {
    SSDR_CHECK_CUDA(cudaEventRecord(c->ev_async, s));
    c->async_stream = s;
    c->async_pending = true;
    return SSDR_OK;
}

int ctx_branch(Ctx* c, int k, cudaStream_t* stream) {
    if (k < 0 || k >= Ctx::MAX_BRANCH) return set_error(SSDR_ERR_INVALID, "branch %d out of range", k);
    if (!c->ev_fork) SSDR_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    if (k > 0) {
        if (!c->bank[k]) c->bank[k] = new DevBuf[WS_SLOTS];
        if (!c->branch_stream[k]) SSDR_CHECK_CUDA(cudaStreamCreateWithFlags(&c->branch_stream[k], cudaStreamNonBlocking));
        if (!c->ev_branch[k]) SSDR_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_branch[k], cudaEventDisableTiming));
    }
    ctx_use_bank(c, k);
    if (stream) *stream = k > 0 ? c->branch_stream[k] : nullptr;
    return SSDR_OK;
}

void ctx_use_bank(Ctx* c, int k) { c->ws = (k > 0 && c->bank[k]) ? c->bank[k] : c->ws0; }

int h2d(Ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    (void)c;
    if (bytes == 0) return SSDR_OK;
    SSDR_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return SSDR_OK;
}

int d2h_sync(Ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    (void)c;
    if (bytes) SSDR_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    SSDR_CHECK_CUDA(cudaStreamSynchronize(s));
    return SSDR_OK;
}

}  // namespace ssdr

using namespace ssdr;

extern "C" {

const char* ssdr_last_error(void) { return g_err; }
int ssdr_version(void) { return 100; }

int ssdr_device_count(int* count) {
    SSDR_REQUIRE(count, SSDR_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return set_error(SSDR_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return SSDR_OK;
}
int ssdr_set_device(int device) {
    SSDR_CHECK_CUDA(cudaSetDevice(device));
    return SSDR_OK;
}
int ssdr_get_device(int* device) {
    SSDR_REQUIRE(device, SSDR_ERR_INVALID, "device is NULL");
    SSDR_CHECK_CUDA(cudaGetDevice(device));
    return SSDR_OK;
}
int ssdr_device_sm_count(int* sms) {
    SSDR_REQUIRE(sms, SSDR_ERR_INVALID, "sms is NULL");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    *sms = c->sm_count;
    return SSDR_OK;
}
int ssdr_synchronize(void) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return SSDR_OK;
}
int ssdr_host_alloc(void** ptr, size_t bytes) {
    SSDR_REQUIRE(ptr, SSDR_ERR_INVALID, "ptr is NULL");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_CHECK_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SSDR_OK;
}
int ssdr_host_free(void* ptr) {
    if (ptr) SSDR_CHECK_CUDA(cudaFreeHost(ptr));
    return SSDR_OK;
}
}
