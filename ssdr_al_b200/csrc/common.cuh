// common.cuh -- runtime plumbing shared by every translation unit of libssdr_b200.so:
// status/error reporting, per-thread context (stream + grow-only workspaces), fork detection, host<->device staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ssdr_b200.h"

namespace ssdr {

// ---- error plumbing ---------------------------------------------------------------------------------
int set_error(int status, const char* fmt, ...);
#define SSDR_CHECK_CUDA(expr)                                                                          \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return ::ssdr::set_error(_e == cudaErrorMemoryAllocation ? SSDR_ERR_NOMEM : SSDR_ERR_CUDA, \
                                     "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                                     __LINE__);                                                        \
    } while (0)
#define SSDR_TRY(expr)           \
    do {                         \
        int _s = (expr);         \
        if (_s != SSDR_OK) return _s; \
    } while (0)
#define SSDR_REQUIRE(cond, status, ...)                         \
    do {                                                        \
        if (!(cond)) return ::ssdr::set_error(status, __VA_ARGS__); \
    } while (0)

// ---- grow-only device / pinned buffers ----------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);  // keeps contents only if no growth is needed
    void release();
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// bumped whenever a workspace of the calling thread is re-allocated: anything that cached raw workspace pointers (a
// captured CUDA graph) compares generations before it trusts them
unsigned long long ws_generation();

enum { WS_SLOTS = 24 };

// One per calling thread and device.
struct Ctx {
    int device = -1;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;  // the library's own stream for host entry points
    cudaStream_t copy_stream = nullptr;  // result read-back that overlaps later kernels of the same call
    cudaEvent_t ev = nullptr;
    cudaEvent_t ev_main = nullptr;
    cudaEvent_t ev_chunk[8] = {};   // one per result chunk in flight on the copy stream
    cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // timing events (stats paths)
    // an entry point that returns before its work is done (ssdr_knn_pyramid_dev) leaves this behind: the stream it ran
    // on and an event behind its last kernel.  The workspaces are shared by every call of the thread, so a later call
    // on ANOTHER stream is ordered behind that event (get_ctx / ctx_order); the same stream is ordered anyway.
    cudaEvent_t ev_async = nullptr;
    cudaStream_t async_stream = nullptr;
    bool async_pending = false;
    DevBuf ws0[WS_SLOTS];           // workspaces, addressed by the owning module
    DevBuf* ws = ws0;               // the ACTIVE bank: every module says c->ws[slot]
    // Branches of one call that run side by side on their own streams (the levels of the KNN pyramid) each work in a
    // bank of their own; bank 0 is ws0.  Banks, streams and events are created on first use and live as long as the Ctx.
    enum { MAX_BRANCH = 52 };  // (16 levels + 1) support clouds x up to 3 item parts, + 1
    DevBuf* bank[MAX_BRANCH] = {};
    cudaStream_t branch_stream[MAX_BRANCH] = {};
    cudaEvent_t ev_branch[MAX_BRANCH] = {};
    cudaEvent_t ev_fork = nullptr;
    // A calling thread that exits gives its streams, events and workspaces back (loader thread pools come and go).
    ~Ctx();
};

// Returns the calling thread's context for its current device (creating it on first use).  The library's own stream
// (host entry points) is ordered behind a pending asynchronous call.
int get_ctx(Ctx** out);
// Device entry points that enqueue on the caller's stream `s`: order it behind a pending asynchronous call that ran on
// a different stream (the per-thread workspaces are about to be reused).
int ctx_order(Ctx* c, cudaStream_t s);
// Marks `s` as carrying an asynchronous call that ends here.
int ctx_mark_async(Ctx* c, cudaStream_t s);
// Branch k of a forked call: makes bank k the active workspace bank and returns its stream (k = 0: bank 0, no stream of
// its own -- the caller's).  ctx_use_bank(c, 0) restores the default before the entry point returns.
int ctx_branch(Ctx* c, int k, cudaStream_t* stream);
void ctx_use_bank(Ctx* c, int k);

// Host -> device copy of `bytes` from an arbitrary host pointer on stream s (pinned sources are truly asynchronous;
// pageable ones are staged by the driver and return once the source may be reused).
int h2d(Ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t s);
// Device -> host, synchronous on return (the caller's buffer is valid afterwards).
int d2h_sync(Ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t s);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace ssdr
