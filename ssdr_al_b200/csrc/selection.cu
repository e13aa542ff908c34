// selection.cu -- farthest-feature sampling and k-center greedy on sm_100a.
//
// Replaces the per-pick numpy passes of farthest_features_sample (fps_gcn_cpu.py:137-146) and
// kCenterGreedy.update_distances/select_batch_ (kcenterGreedy.py:60-128).
//
// One PERSISTENT cooperative kernel runs every pick.  Every WARP owns a contiguous run of rows and streams it from
// HBM/L2 through its own multi-stage shared-memory ring (no block-wide barrier inside a pick; the ring keeps running
// across picks, so the next pick's first rows are already in flight while the grid agrees on the next centre).
// Ring feeders:
//   * D == 32 float32 (the benchmark dimension): ONE TMA tensor copy per 32-row unit (cp.async.bulk.tensor.2d through a
//     128-byte-swizzled tensor map, completion on an mbarrier); every lane then owns one whole row and reads it with
//     conflict-free LDS.128 (the swizzle spreads the eight rows of a quarter-warp over the banks), all eight pairwise
//     accumulators live in the lane -- no shuffles, half the instructions of the 8-lanes-per-row mapping.
//   * D in {64,128,256}, 16-byte aligned rows: 1-D TMA bulk copies (cp.async.bulk.shared::cluster.global), one per row,
//     into padded rows, completion on an mbarrier (expect-tx); 8 lanes share a row as before.
//   * anything else (odd D, float64 rows that are not 16-byte multiples): the 16/8/4-byte cp.async ring.
// Multi-GPU (row shards, one process per GPU): the per-pick argmax exchange is FUSED into the same kernel -- CTA 0 of
// every rank stores the rank winner straight into every peer's mailbox over NVLink (peer memory mapped with CUDA IPC),
// all CTAs of all ranks poll their local copy; no per-pick launch, no NCCL call, the ring never drains.  The warp evaluates the distance of its rows to the current centre in
// the reference's exact floating-point order, folds it into the running min-distance with coalesced 128-byte
// accesses and keeps a (distance, index) candidate per lane.  A grid barrier publishes one candidate per CTA; every
// CTA then reduces them redundantly, so the next centre is known everywhere with a single barrier per pick and no
// host round trip.  The step is HBM/L2-bandwidth bound: N*(sizeof(T)*D + 2*sizeof(T)).  D in {32,64,128,256} is
// compiled in (fully unrolled, centre row in registers); any other D takes the generic leaf-plan path.
//
// Arithmetic contracts (bit-exact picks):
//   FPS     : d = pairwise_sum_j((F[i,j]-F[c,j])^2) in T with numpy's summation tree: eight strided
//             accumulators per <=128-wide leaf, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), sequential tail, leaves
//             combined by recursive halving.  Eight lanes own the eight accumulators of one row; the xor-butterfly
//             reproduces the combine tree exactly because IEEE addition is commutative.
//   k-center: d = sqrt(max(0, T(-2*dot64 + |x|^2 + |c|^2))) like sklearn's euclidean_distances (float64
//             accumulation, rounded to T before clamp and sqrt).  BLAS's summation order is unspecified; ours is
//             eight strided fp64 FMA chains + butterfly.
#include <cooperative_groups.h>
#include <cuda.h>
#include <math.h>
#include <mutex>
#include <stdlib.h>

#include "common.cuh"

namespace ssdr {
int nccl_allreduce_max_u64(void* comm, unsigned long long* buf, size_t count, cudaStream_t stream);  // nccl_shim.cu
namespace sel {

constexpr int THREADS = 512;          // upper bound; the host launches p.nwarps*32 threads
constexpr int WARPS = THREADS / 32;
constexpr int MAX_LEAVES = 128;
constexpr int MAX_STACK = 10;
constexpr int MAX_STAGES = 8;

enum Mode { MODE_FPS = 0, MODE_KCENTER = 1 };

__host__ __device__ constexpr size_t align_up_dev(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct LeafPlan {
    int n_leaves;
    unsigned short start[MAX_LEAVES];
    unsigned short len[MAX_LEAVES];
    unsigned char merges[MAX_LEAVES];  // number of (left + right) combines after finishing leaf i
};

constexpr int MAX_PEERS = 16;
struct Mailbox;

// Cross-GPU part of the per-pick exchange (world == 1: unused).  xbox[r] is rank r's mailbox block, mapped into this
// process (xbox[rank] is the local one): [2 parities][world source ranks].  Tags are tag_base + step + 1 and the
// parity is the tag's low bit, so consecutive picks -- also across calls -- alternate slots and no block ever needs
// to be cleared while peers may already be writing the next pick into it.
struct PeerParams {
    int world = 1, rank = 0;
    unsigned tag_base = 0;
    Mailbox* xbox[MAX_PEERS] = {};
    unsigned long long timeout_ns = 0;
    unsigned* error = nullptr;  // device flag (local): 1 = a peer did not answer in time
};

template <typename T>
struct Params {
    const T* F;
    unsigned long long N;
    int D;
    unsigned long long row_begin, row_end;  // rows scanned by this launch (global indices)
    int stride;                             // smem row stride in elements
    int groups_per_unit;                    // a warp iteration ("unit") covers 4*groups_per_unit rows (<= 32)
    int nwarps;                             // warps per CTA (8 unless a very wide row forces fewer)
    int nstages;
    int vec;                                // elements per cp.async (16B when possible)
    const long long* forced;                // device: centres of the first n_forced steps
    int n_forced;
    int step_begin, step_end;
    T* mind;                                // running min distance, N entries
    const double* xx;                       // k-center: squared row norms (fp64)
    unsigned long long* winners;            // per step: packed winner (see pack())
    unsigned long long* cand;               // 2 * gridDim * 2 u64 candidate mailboxes
    unsigned int* barrier;                  // zero at launch
    long long* picks;                       // nullable: picks[s] = winner row of step s
    PeerParams peer;
    LeafPlan plan;
};

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        default: cp_async_wait<5>(); break;
    }
}

template <typename T>
__device__ __forceinline__ void cp_elems(T* dst, const T* src, int vec) {
    if (sizeof(T) * vec == 16) cp_async_16(dst, src);
    else if (sizeof(T) * vec == 8) cp_async_8(dst, src);
    else cp_async_4(dst, src);
}

// ---- mbarrier + TMA (bulk async copies complete on a shared-memory barrier by transaction bytes) -----------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 2-D tiled tensor copy global -> shared through a tensor map (c0 = innermost coordinate)
__device__ __forceinline__ void tma_g2s_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(smem_u32(dst)),
        "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// exact, non-contracted arithmetic
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
template <typename T>
__device__ __forceinline__ T sqdiff(T a, T c) {
    T d = xsub(a, c);
    return xmul(d, d);
}
__device__ __forceinline__ float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) over the 8 lanes of a row group
template <typename T>
__device__ __forceinline__ T butterfly8(T v) {
    v = xadd(v, shfl_xor(v, 1));
    v = xadd(v, shfl_xor(v, 2));
    v = xadd(v, shfl_xor(v, 4));
    return v;
}

// numpy pairwise sum of (a[i]-c[i])^2 for one row; the 8 lanes with the same (lane>>3) cooperate, j = lane&7.
template <typename T>
__device__ __forceinline__ T row_sqdist(const T* __restrict__ a, const T* __restrict__ c, int D, int j,
                                        const LeafPlan& plan) {
    if (D < 8) {
        T res = (T)0;
        for (int i = 0; i < D; ++i) res = xadd(res, sqdiff(a[i], c[i]));
        return res;
    }
    T st[MAX_STACK];
#pragma unroll 1
    for (int l = 0; l < plan.n_leaves; ++l) {
        const int s = plan.start[l], n = plan.len[l];
        const int nm = n - (n & 7);
        T acc = sqdiff(a[s + j], c[s + j]);
        for (int i = 8; i < nm; i += 8) acc = xadd(acc, sqdiff(a[s + i + j], c[s + i + j]));
        T res = butterfly8(acc);
        for (int i = nm; i < n; ++i) res = xadd(res, sqdiff(a[s + i], c[s + i]));
        if (plan.n_leaves == 1) return res;
        // shift-register stack (static indexing keeps it in registers)
#pragma unroll
        for (int q = MAX_STACK - 1; q > 0; --q) st[q] = st[q - 1];
        st[0] = res;
        for (int m = plan.merges[l]; m > 0; --m) {
            st[0] = xadd(st[1], st[0]);  // left + right
#pragma unroll
            for (int q = 1; q < MAX_STACK - 1; ++q) st[q] = st[q + 1];
        }
    }
    return st[0];
}

// fp64 dot product of one row with the centre (k-center); order: 8 strided FMA chains + butterfly + tail.
template <typename T>
__device__ __forceinline__ double row_dot64(const T* __restrict__ a, const T* __restrict__ c, int D, int j) {
    double acc = 0.0;
    const int nm = D - (D & 7);
    for (int i = 0; i < nm; i += 8) acc = fma((double)a[i + j], (double)c[i + j], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    for (int i = nm; i < D; ++i) acc = fma((double)a[i], (double)c[i], acc);
    return acc;
}

// ---- candidate = (distance, row) with "max distance, then lowest row" order ----------------------------
struct Cand {
    unsigned long long hi, lo;  // float: hi = dist_bits<<32 | ~row (lo unused); double: hi = dist bits, lo = ~row
};
template <typename T>
__device__ __forceinline__ Cand make_cand(T d, unsigned long long row);
template <>
__device__ __forceinline__ Cand make_cand<float>(float d, unsigned long long row) {
    Cand c;
    c.hi = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)row);
    c.lo = 0;
    return c;
}
template <>
__device__ __forceinline__ Cand make_cand<double>(double d, unsigned long long row) {
    Cand c;
    c.hi = (unsigned long long)__double_as_longlong(d);
    c.lo = ~row;
    return c;
}
template <typename T>
__device__ __forceinline__ unsigned long long cand_row(const Cand& c);
template <>
__device__ __forceinline__ unsigned long long cand_row<float>(const Cand& c) {
    return (unsigned long long)(0xFFFFFFFFu - (unsigned)(c.hi & 0xFFFFFFFFull));
}
template <>
__device__ __forceinline__ unsigned long long cand_row<double>(const Cand& c) {
    return ~c.lo;
}
template <typename T>
__device__ __forceinline__ Cand cand_max(const Cand& a, const Cand& b) {
    if (sizeof(T) == 4) {  // packed in hi alone
        Cand r;
        r.hi = a.hi < b.hi ? b.hi : a.hi;
        r.lo = 0;
        return r;
    }
    const bool less = a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
    return less ? b : a;
}
template <typename T>
__device__ __forceinline__ Cand warp_max(Cand c) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        Cand o;
        o.hi = __shfl_xor_sync(0xffffffffu, c.hi, m);
        o.lo = sizeof(T) == 4 ? 0ull : __shfl_xor_sync(0xffffffffu, c.lo, m);
        c = cand_max<T>(c, o);
    }
    return c;
}

// compile-time pairwise tree for D % 8 == 0 (numpy: leaves <= 128 wide, recursive halving above); creg[k] = c[8k+j]
template <typename T, int N, int OFF>
__device__ __forceinline__ T pw_fixed(const T* __restrict__ a, const T* creg, int j) {
    if constexpr (N <= 128) {
        T acc = sqdiff(a[OFF + j], creg[OFF / 8]);
#pragma unroll
        for (int i = 8; i < N; i += 8) acc = xadd(acc, sqdiff(a[OFF + i + j], creg[(OFF + i) / 8]));
        return butterfly8(acc);
    } else {
        constexpr int N2 = (N / 2) - ((N / 2) % 8);
        const T l = pw_fixed<T, N2, OFF>(a, creg, j);
        const T r = pw_fixed<T, N - N2, OFF + N2>(a, creg, j);
        return xadd(l, r);
    }
}
template <typename T, int N>
__device__ __forceinline__ double dot64_fixed(const T* __restrict__ a, const T* creg, int j) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; i += 8) acc = fma((double)a[i + j], (double)creg[i / 8], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    return acc;
}

// Mailbox of one CTA for one step parity, low-latency style: the candidate travels as 32-bit chunks, each chunk
// paired with the 32-bit step tag inside ONE naturally atomic 8-byte word.  A reader that sees the tag in a word has
// the chunk too, so a pick costs one plain store burst and one (parallel) poll -- no fence, no counter, no second read.
// The same layout crosses NVLink: an aligned 8-byte store to peer memory is a single transaction.
struct __align__(32) Mailbox {
    unsigned long long w[4];  // float: w[0] = {dist bits, tag}, w[1] = {~row, tag}; double: four chunks
};
template <bool SYS>
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    if (SYS) asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
    else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
template <bool SYS>
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    if (SYS) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <typename T, bool SYS>
__device__ __forceinline__ void mailbox_post(Mailbox* mb, const Cand& c, unsigned tag) {
    const unsigned long long t = (unsigned long long)tag;
    st_relaxed_u64<SYS>(&mb->w[0], ((c.hi >> 32) << 32) | t);
    st_relaxed_u64<SYS>(&mb->w[1], ((c.hi & 0xFFFFFFFFull) << 32) | t);
    if (sizeof(T) == 8) {
        st_relaxed_u64<SYS>(&mb->w[2], ((c.lo >> 32) << 32) | t);
        st_relaxed_u64<SYS>(&mb->w[3], ((c.lo & 0xFFFFFFFFull) << 32) | t);
    }
}
// raw chunk words of one mailbox (all loads independent, so several mailboxes can be in flight per lane)
template <typename T, bool SYS>
__device__ __forceinline__ void mailbox_load(const Mailbox* mb, unsigned long long (&v)[4]) {
    v[0] = ld_relaxed_u64<SYS>(&mb->w[0]);
    v[1] = ld_relaxed_u64<SYS>(&mb->w[1]);
    if (sizeof(T) == 8) {
        v[2] = ld_relaxed_u64<SYS>(&mb->w[2]);
        v[3] = ld_relaxed_u64<SYS>(&mb->w[3]);
    } else {
        v[2] = v[3] = 0;
    }
}
// true when every chunk carries `tag`; then `out` holds the candidate
template <typename T>
__device__ __forceinline__ bool mailbox_decode(const unsigned long long (&v)[4], unsigned tag, Cand* out) {
    bool ok = (unsigned)v[0] == tag && (unsigned)v[1] == tag;
    out->hi = ((v[0] >> 32) << 32) | (v[1] >> 32);
    out->lo = 0;
    if (sizeof(T) == 8) {
        ok = ok && (unsigned)v[2] == tag && (unsigned)v[3] == tag;
        out->lo = ((v[2] >> 32) << 32) | (v[3] >> 32);
    }
    return ok;
}

// ---- the per-pick exchange, executed by warp 0 of every CTA -----------------------------------------------------------
// c = this CTA's candidate (warp uniform).  Level 1 (inside the GPU): every CTA posts into its own mailbox and polls all
// G of them, so the rank winner is known in every CTA after one store burst + one poll.  Level 2 (between the GPUs of a
// row-sharded job): CTA 0 stores the rank winner into every peer's block over NVLink, every CTA of every rank polls
// its LOCAL block (world mailboxes).  Returns the global winner; *abort is set when a peer did not answer in time (or
// another CTA of this rank already gave up), in which case the caller leaves the pick loop.
template <typename T>
__device__ __forceinline__ Cand grid_exchange(const Params<T>& p, Cand c, int step, int lane, bool* abort) {
    const int G = gridDim.x;
    Mailbox* boxes = reinterpret_cast<Mailbox*>(p.cand);
    const unsigned tag = (unsigned)(step + 1);
    Mailbox* row_boxes = boxes + (size_t)(step & 1) * G;
    if (lane == 0) mailbox_post<T, false>(row_boxes + blockIdx.x, c, tag);
    // poll all mailboxes: every lane keeps its loads in flight together and retries only the missing ones
    Cand w{0, 0};
    constexpr int KM = 5;  // 5 x 32 = 160 >= 148 CTAs
    const bool multi = p.peer.world > 1;
    unsigned pending = 0;
#pragma unroll
    for (int k = 0; k < KM; ++k)
        if (lane + 32 * k < G) pending |= 1u << k;
    unsigned spins = 0;
    bool dead = false;
    while (__any_sync(0xffffffffu, pending != 0)) {
        unsigned long long v[KM][4];
#pragma unroll
        for (int k = 0; k < KM; ++k)
            if (pending & (1u << k)) mailbox_load<T, false>(row_boxes + lane + 32 * k, v[k]);
#pragma unroll
        for (int k = 0; k < KM; ++k) {
            if (pending & (1u << k)) {
                Cand o;
                if (mailbox_decode<T>(v[k], tag, &o)) {
                    w = cand_max<T>(w, o);
                    pending &= ~(1u << k);
                }
            }
        }
        // a CTA of this rank that gave up on a peer never posts again: notice it instead of spinning forever
        if (multi && (++spins & 1023u) == 0 && ld_relaxed_u32(p.peer.error) != 0) {
            dead = true;
            break;
        }
    }
    for (int bb = lane + 32 * KM; bb < G && !dead; bb += 32) {  // larger grids (not on B200): simple spin
        Cand o;
        unsigned long long v[4];
        do {
            mailbox_load<T, false>(row_boxes + bb, v);
        } while (!mailbox_decode<T>(v, tag, &o));
        w = cand_max<T>(w, o);
    }
    __syncwarp();
    w = warp_max<T>(w);
    if (multi) {
        const int W = p.peer.world;
        const unsigned gtag = p.peer.tag_base + (unsigned)step + 1u;
        const unsigned par = gtag & 1u;
        if (blockIdx.x == 0 && lane < W && !dead)  // NVLink stores, one peer per lane (the local block included)
            mailbox_post<T, true>(p.peer.xbox[lane] + (size_t)par * W + p.peer.rank, w, gtag);
        const Mailbox* mine = p.peer.xbox[p.peer.rank] + (size_t)par * W;
        Cand o{0, 0};
        bool got = lane >= W;
        const unsigned long long t0 = global_timer_ns();
        unsigned it = 0;
        while (!dead && __any_sync(0xffffffffu, !got)) {
            if (!got) {
                unsigned long long v[4];
                mailbox_load<T, true>(mine + lane, v);
                got = mailbox_decode<T>(v, gtag, &o);
            }
            if ((++it & 255u) == 0) {
                if (ld_relaxed_u32(p.peer.error) != 0 || global_timer_ns() - t0 > p.peer.timeout_ns) dead = true;
                dead = __any_sync(0xffffffffu, dead);
            }
        }
        dead = __any_sync(0xffffffffu, dead);
        if (dead) {
            if (lane == 0 && atomicExch(p.peer.error, 1u) == 0u) {
                // first CTA to give up leaves a trace for the host's error message: step, expected tag, and the tag /
                // payload words found in this rank's block for the current parity
                unsigned* dbg = p.peer.error;
                dbg[1] = (unsigned)step;
                dbg[2] = gtag;
                dbg[3] = blockIdx.x;
                for (int r = 0; r < W && r < 8; ++r) {
                    const unsigned long long w0 = ld_relaxed_u64<true>(&mine[r].w[0]);
                    const unsigned long long w1 = ld_relaxed_u64<true>(&mine[r].w[1]);
                    dbg[4 + 4 * r] = (unsigned)w0;
                    dbg[5 + 4 * r] = (unsigned)(w0 >> 32);
                    dbg[6 + 4 * r] = (unsigned)w1;
                    dbg[7 + 4 * r] = (unsigned)(w1 >> 32);
                }
            }
            *abort = true;
        }
        w = warp_max<T>(o);
    }
    return w;
}

// One fetch of a centre row per CTA (16-byte chunks by the lanes of warp 0) instead of one per warp: every CTA of the
// grid wants the same row at the same moment, and 148 x 16 warps asking one L2 slice for the same line was measured
// as 0.3 us per warp of a CTA on the critical path of every pick.
template <typename T>
__device__ __forceinline__ void fetch_center(const Params<T>& p, unsigned long long c_row, T* s_center, double* s_xx,
                                             int lane) {
    const int chunks = (int)((unsigned)p.D * sizeof(T) / 16u);
    const float4* src = reinterpret_cast<const float4*>(p.F + c_row * (unsigned long long)p.D);
    for (int i = lane; i < chunks; i += 32) reinterpret_cast<float4*>(s_center)[i] = __ldcg(src + i);
    if (p.xx && lane == 0) *s_xx = __ldcg(p.xx + c_row);  // k-center: the centre's squared norm travels with its row
}

// Everything after the row scan of a pick: CTA candidate -> exchange -> next centre (and the pick record).  Called by all
// threads; returns false when the pick loop must be left (peer time-out).  s_center (nullable): staging of the next
// centre row for kernels whose rows are 16-byte multiples.
template <typename T>
__device__ __forceinline__ bool finish_pick(const Params<T>& p, Cand best, int step, int lane, int warp, int nwarp,
                                            Cand* s_red, unsigned long long* s_next_center, int* s_abort,
                                            T* s_center = nullptr, double* s_xx = nullptr) {
    best = warp_max<T>(best);
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (warp == 0) {
        Cand c = lane < nwarp ? s_red[lane] : Cand{0, 0};
        c = warp_max<T>(c);
        bool abort = false;
        const Cand w = grid_exchange<T>(p, c, step, lane, &abort);
        if (s_center && !abort && step + 1 < p.step_end) {
            const int next = step + 1;
            fetch_center<T>(p, next < p.n_forced ? (unsigned long long)p.forced[next] : cand_row<T>(w), s_center, s_xx,
                            lane);
        }
        if (lane == 0) {
            *s_next_center = cand_row<T>(w);
            if (abort) *s_abort = 1;
            if (blockIdx.x == 0 && !abort) {
                p.winners[2 * step] = w.hi;
                p.winners[2 * step + 1] = w.lo;
                if (p.picks) p.picks[step] = (long long)cand_row<T>(w);
            }
        }
    }
    __syncthreads();  // publishes s_next_center / s_abort / the staged centre row
    return *s_abort == 0;
}

// first centre of a launch range: forced (FPS start / already selected rows) or the previous launch's winner
template <typename T>
__device__ __forceinline__ unsigned long long step_center(const Params<T>& p, int step, unsigned long long s_next) {
    if (step < p.n_forced) return (unsigned long long)p.forced[step];
    if (step == p.step_begin) {
        Cand w;
        w.hi = p.winners[2 * (step - 1)];
        w.lo = p.winners[2 * (step - 1) + 1];
        return cand_row<T>(w);
    }
    return s_next;
}

// min-distance update of one row and its candidate: d is the squared distance (FPS) or the fp64 dot product (k-center)
template <typename T, int MODE, typename RowVal>
__device__ __forceinline__ void update_row(const Params<T>& p, RowVal mine, T m_old, double xx_r, double xx_c,
                                           unsigned long long gr, Cand* best) {
    T d;
    if (MODE == MODE_FPS) {
        d = (T)mine;
    } else {
        double v = -2.0 * (double)mine;
        v = __dadd_rn(v, xx_r);
        v = __dadd_rn(v, xx_c);
        T tv = (T)v;
        tv = tv > (T)0 ? tv : (T)0;  // np.maximum(d, 0)
        d = sizeof(T) == 4 ? (T)__fsqrt_rn((float)tv) : (T)__dsqrt_rn((double)tv);
    }
    const T m = d < m_old ? d : m_old;
    if (d < m_old) __stcg(p.mind + gr, m);
    *best = cand_max<T>(*best, make_cand<T>(m, gr));
}

// contiguous run of `unit`-row blocks owned by one warp
struct WarpRun {
    unsigned long long row0, rows;  // first row (global index) and number of rows of this warp
    int n_it;                       // units per pick
};
__device__ __forceinline__ WarpRun warp_run(unsigned long long row_begin, unsigned long long row_end, int RW, int G,
                                            int nwarp, int warp) {
    const unsigned long long nrows = row_end - row_begin;
    const unsigned long long units_total = (nrows + RW - 1) / RW;
    const unsigned long long TW = (unsigned long long)G * nwarp;
    // spread the remainder: the first `extra` warps (interleaved over the CTAs) take one unit more
    const unsigned long long base = units_total / TW, extra = units_total % TW;
    const unsigned long long gw = (unsigned long long)warp * G + blockIdx.x;  // CTA-interleaved global warp id
    const unsigned long long u_begin = gw * base + (gw < extra ? gw : extra);
    const unsigned long long n = base + (gw < extra ? 1 : 0);
    WarpRun r;
    r.n_it = (int)n;
    r.row0 = row_begin + u_begin * (unsigned long long)RW;
    const unsigned long long lim = row_end > r.row0 ? row_end - r.row0 : 0ull;
    r.rows = n * RW < lim ? n * RW : lim;
    return r;
}

// FEED: 0 = cp.async ring (any D / alignment), 1 = one 1-D TMA bulk copy per row (rows are 16-byte multiples)
template <typename T, int MODE, int DT, int GT, int FEED>  // DT: compile-time D (0 = generic), GT: row groups per unit (0 = runtime)
__global__ void __launch_bounds__(THREADS, 1) select_kernel(const Params<T> p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename std::conditional<MODE == MODE_KCENTER, double, T>::type RowVal;
    const int D = DT > 0 ? DT : p.D;
    const int stride = p.stride;
    const int NG = GT > 0 ? GT : p.groups_per_unit;
    const int RW = 4 * NG;  // rows per unit
    const int S = p.nstages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int NWARP = p.nwarps;
    const int g = lane >> 3, j = lane & 7;

    // shared layout: centre row (generic path) | reduction scratch | mbarriers | per-warp stage rings
    T* s_center = reinterpret_cast<T*>(smem_raw);
    unsigned off = (unsigned)align_up_dev((size_t)(p.D + 8) * sizeof(T), 16);
    Cand* s_red = reinterpret_cast<Cand*>(smem_raw + off);
    off += (unsigned)align_up_dev(WARPS * sizeof(Cand), 16);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + off) + warp * MAX_STAGES;
    off += (unsigned)align_up_dev((size_t)WARPS * MAX_STAGES * 8, 128);
    const unsigned unit_elems = (unsigned)RW * (unsigned)stride;
    T* ring = reinterpret_cast<T*>(smem_raw + off) + (size_t)warp * S * unit_elems;
    __shared__ unsigned long long s_next_center;
    __shared__ double s_xx_c;
    __shared__ int s_abort;
    if (tid == 0) s_abort = 0;
    if (FEED == 1) {
        if (lane == 0) {
            for (int q = 0; q < S; ++q) mbar_init(&bars[q], 1);
            mbar_fence_init();
        }
        __syncwarp();
    }

    const WarpRun run = warp_run(p.row_begin, p.row_end, RW, G, NWARP, warp);
    const int n_it = run.n_it;
    const unsigned long long w_row0 = run.row0, w_rows = run.rows;
    long long loads_left = (long long)n_it * (p.step_end - p.step_begin);

    const int cpr = D / p.vec;  // cp.async chunks per row
    // ring feeder state (no divisions in the loop): next unit to load and the stage it goes to
    int ld_it = 0, ld_stage = 0;
    auto issue = [&]() {
        if (loads_left > 0) {
            --loads_left;
            const unsigned r_in = (unsigned)ld_it * (unsigned)RW;
            const int rows = (int)min((unsigned long long)RW, w_rows - r_in);
            T* dst = ring + (unsigned)ld_stage * unit_elems;
            const T* src = p.F + (w_row0 + r_in) * (unsigned long long)D;
            if constexpr (FEED == 1) {
                // the stage was read by generic-proxy loads; order them before the async-proxy writes
                const unsigned row_bytes = (unsigned)D * (unsigned)sizeof(T);
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&bars[ld_stage], (unsigned)rows * row_bytes);
                }
                __syncwarp();
                if (lane < rows)
                    bulk_g2s(dst + (unsigned)lane * (unsigned)stride, src + (unsigned)lane * (unsigned)D, row_bytes,
                             &bars[ld_stage]);
            } else if constexpr (DT > 0 && GT > 0) {
                // compile-time geometry: 16-byte chunks, chunk q of the unit = lane + 32k; the source is contiguous
                constexpr int VEC = 16 / (int)sizeof(T);
                constexpr int CPR = DT / VEC;                // chunks per row (power of two)
                constexpr int CH = 4 * GT * CPR / 32;        // chunks per lane per unit
                const int nchunks = rows * CPR;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const int q = lane + 32 * k;
                    const int row = q / CPR, col = (q % CPR) * VEC;
                    if (q < nchunks) cp_async_16(dst + (unsigned)row * (unsigned)stride + col, src + (unsigned)q * VEC);
                }
            } else if (cpr <= 32 && (32 % cpr) == 0) {  // a warp pass covers 32/cpr whole rows
                const int rpp = 32 / cpr, col = (lane % cpr) * p.vec;
                for (int row = lane / cpr; row < rows; row += rpp)
                    cp_elems(dst + (unsigned)row * stride + col, src + (unsigned)row * D + col, p.vec);
            } else {
                for (int row = 0; row < rows; ++row)
                    for (int c = lane; c < cpr; c += 32)
                        cp_elems(dst + (unsigned)row * stride + c * p.vec, src + (unsigned)row * D + c * p.vec, p.vec);
            }
            if (++ld_it == n_it) ld_it = 0;
            if (++ld_stage == S) ld_stage = 0;
        }
        if (FEED == 0) cp_async_commit();
    };

    for (int q = 0; q < S - 1; ++q) issue();  // the ring never drains between picks
    int cur_stage = 0;
    unsigned cur_phase = 0;
    long long waits_left = (long long)n_it * (p.step_end - p.step_begin);

    for (int step = p.step_begin; step < p.step_end; ++step) {
        const unsigned long long c_row = step_center<T>(p, step, s_next_center);
        T creg[DT > 0 ? DT / 8 : 1];
        if (DT > 0) {
            if (step == p.step_begin) {  // later picks find the row staged by finish_pick
                if (warp == 0) fetch_center<T>(p, c_row, s_center, &s_xx_c, lane);
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < (DT > 0 ? DT / 8 : 1); ++k) creg[k] = s_center[8 * k + j];
        } else {
            for (int i = tid; i < D; i += NWARP * 32) s_center[i] = p.F[c_row * (unsigned long long)D + i];
            if (MODE == MODE_KCENTER && tid == 0) s_xx_c = p.xx[c_row];
            __syncthreads();
        }
        double xx_c = 0.0;
        if (MODE == MODE_KCENTER) xx_c = s_xx_c;

        Cand best;
        best.hi = 0;
        best.lo = 0;
        unsigned r_in = 0;
        for (int it = 0; it < n_it; ++it, r_in += RW) {
            issue();
            if (FEED == 1) mbar_wait(&bars[cur_stage], cur_phase);
            else cp_async_wait_dyn(S - 1);
            --waits_left;
            __syncwarp();
            const int rows = (int)min((unsigned long long)RW, w_rows - r_in);
            const unsigned long long r0 = w_row0 + r_in;
            const T* tile = ring + (unsigned)cur_stage * unit_elems;
            if (++cur_stage == S) {
                cur_stage = 0;
                cur_phase ^= 1u;
            }
            // running min-distance (and row norm) of "my" row: issued now, consumed after the distance loop
            T m_old = (T)0;
            double xx_r = 0.0;
            if (lane < rows) {
                m_old = __ldcg(p.mind + r0 + lane);
                if (MODE == MODE_KCENTER) xx_r = __ldg(p.xx + r0 + lane);
            }
            RowVal mine = (RowVal)0;
#pragma unroll
            for (int i = 0; i < (GT > 0 ? GT : 8); ++i) {
                if (GT == 0 && i >= NG) break;
                const T* a = tile + (unsigned)(4 * i + g) * (unsigned)stride;
                RowVal d;
                if (MODE == MODE_FPS) {
                    if (DT > 0) d = (RowVal)pw_fixed<T, (DT > 0 ? DT : 8), 0>(a, creg, j);
                    else d = (RowVal)row_sqdist<T>(a, s_center, D, j, p.plan);
                } else {
                    if (DT > 0) d = (RowVal)dot64_fixed<T, (DT > 0 ? DT : 8)>(a, creg, j);
                    else d = (RowVal)row_dot64<T>(a, s_center, D, j);
                }
                // lane L owns row L of the unit: it lives in group L/4, row-in-group L%4 (lanes (L%4)*8.. hold it)
                const RowVal v = __shfl_sync(0xffffffffu, d, (lane & 3) * 8);
                if ((lane >> 2) == i) mine = v;
            }
            if (lane < rows) update_row<T, MODE, RowVal>(p, mine, m_old, xx_r, xx_c, r0 + lane, &best);
            __syncwarp();  // every lane is done with this stage before a later issue() overwrites it
        }
        if (!finish_pick<T>(p, best, step, lane, warp, NWARP, s_red, &s_next_center, &s_abort,
                            DT > 0 ? s_center : (T*)nullptr, &s_xx_c))
            break;
    }
    // drain: copies still in flight must land before the CTA's shared memory is released
    if (FEED == 1) {
        long long outstanding = waits_left - loads_left;  // issued - consumed
        while (outstanding-- > 0) {
            mbar_wait(&bars[cur_stage], cur_phase);
            if (++cur_stage == S) {
                cur_stage = 0;
                cur_phase ^= 1u;
            }
        }
    } else {
        cp_async_wait<0>();
    }
}

// ---- D == 32 float32: lane-per-row over a 128-byte-swizzled TMA tile --------------------------------------------------
// A unit is 32 consecutive rows = one 4 KB tensor box.  The swizzle stores the 16-byte chunk c of tile row r at chunk
// position c ^ (r & 7), so the eight lanes of a quarter-warp (rows r..r+7, same logical chunk) hit eight different
// 16-byte bank groups: LDS.128 without conflicts and without padding.  Lane = row: the eight strided accumulators of
// numpy's pairwise sum (r[j] += a[8k+j]) and the final ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) are all in-lane.
constexpr int L32_UNIT_BYTES = 32 * 32 * 4;

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) select32_kernel(const Params<float> p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem32_raw[];
    typedef float T;
    const int S = p.nstages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int NWARP = p.nwarps;
    // shared layout: per-warp stage rings (1024-byte aligned tiles: the swizzle is a function of address bits 7-9) |
    // mbarriers | reduction scratch
    unsigned char* smem_al = smem32_raw + ((1024u - (smem_u32(smem32_raw) & 1023u)) & 1023u);
    unsigned char* ring = smem_al + (size_t)warp * S * L32_UNIT_BYTES;
    unsigned off = (unsigned)NWARP * (unsigned)S * L32_UNIT_BYTES;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_al + off) + warp * MAX_STAGES;
    off += WARPS * MAX_STAGES * 8;
    Cand* s_red = reinterpret_cast<Cand*>(smem_al + off);
    off += WARPS * (unsigned)sizeof(Cand);
    float* s_center = reinterpret_cast<float*>(smem_al + off);  // 32 floats: the current centre row, fetched once per CTA
    __shared__ unsigned long long s_next_center;
    __shared__ double s_xx_c;
    __shared__ int s_abort;
    if (tid == 0) s_abort = 0;
    if (lane == 0) {
        for (int q = 0; q < S; ++q) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const WarpRun run = warp_run(p.row_begin, p.row_end, 32, G, NWARP, warp);
    const int n_it = run.n_it;
    const unsigned long long w_row0 = run.row0, w_rows = run.rows;
    const long long total = (long long)n_it * (p.step_end - p.step_begin);
    long long loads_left = total, waits_left = total;
    int ld_it = 0, ld_stage = 0;
    auto issue = [&]() {
        if (loads_left > 0) {
            --loads_left;
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&bars[ld_stage], L32_UNIT_BYTES);  // rows past the end of the matrix are zero-filled
                tma_g2s_2d(ring + (size_t)ld_stage * L32_UNIT_BYTES, &tmap, 0, (int)(w_row0 + (unsigned)ld_it * 32u),
                           &bars[ld_stage]);
            }
            if (++ld_it == n_it) ld_it = 0;
            if (++ld_stage == S) ld_stage = 0;
        }
    };
    for (int q = 0; q < S - 1; ++q) issue();
    int cur_stage = 0;
    unsigned cur_phase = 0;
    const unsigned sw = (unsigned)(lane & 7);

    for (int step = p.step_begin; step < p.step_end; ++step) {
        const unsigned long long c_row = step_center<T>(p, step, s_next_center);
        // the centre row in registers (every lane holds all 32 values; the loads are warp-uniform broadcasts)
        typedef typename std::conditional<MODE == MODE_KCENTER, double, float>::type CVal;
        CVal creg[32];
        if (step == p.step_begin) {  // later picks find the row staged by finish_pick
            if (warp == 0) fetch_center<float>(p, c_row, s_center, &s_xx_c, lane);
            __syncthreads();
        }
        {
            const float4* c4 = reinterpret_cast<const float4*>(s_center);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = c4[c];
                creg[4 * c + 0] = (CVal)v.x;
                creg[4 * c + 1] = (CVal)v.y;
                creg[4 * c + 2] = (CVal)v.z;
                creg[4 * c + 3] = (CVal)v.w;
            }
        }
        double xx_c = 0.0;
        if (MODE == MODE_KCENTER) xx_c = s_xx_c;
        Cand best;
        best.hi = 0;
        best.lo = 0;
        unsigned r_in = 0;
        for (int it = 0; it < n_it; ++it, r_in += 32) {
            issue();
            mbar_wait(&bars[cur_stage], cur_phase);
            --waits_left;
            const int rows = (int)min(32ull, w_rows - r_in);
            const unsigned long long r0 = w_row0 + r_in;
            const float4* t4 = reinterpret_cast<const float4*>(ring + (size_t)cur_stage * L32_UNIT_BYTES) + lane * 8;
            if (++cur_stage == S) {
                cur_stage = 0;
                cur_phase ^= 1u;
            }
            float m_old = 0.f;
            double xx_r = 0.0;
            if (lane < rows) {
                m_old = __ldcg(p.mind + r0 + lane);
                if (MODE == MODE_KCENTER) xx_r = __ldg(p.xx + r0 + lane);
            }
            float4 v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = t4[(unsigned)c ^ sw];
            if (MODE == MODE_FPS) {
                float acc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int o = (c & 1) * 4;
                    if (c < 2) {
                        acc[o + 0] = sqdiff(v[c].x, (float)creg[4 * c + 0]);
                        acc[o + 1] = sqdiff(v[c].y, (float)creg[4 * c + 1]);
                        acc[o + 2] = sqdiff(v[c].z, (float)creg[4 * c + 2]);
                        acc[o + 3] = sqdiff(v[c].w, (float)creg[4 * c + 3]);
                    } else {
                        acc[o + 0] = xadd(acc[o + 0], sqdiff(v[c].x, (float)creg[4 * c + 0]));
                        acc[o + 1] = xadd(acc[o + 1], sqdiff(v[c].y, (float)creg[4 * c + 1]));
                        acc[o + 2] = xadd(acc[o + 2], sqdiff(v[c].z, (float)creg[4 * c + 2]));
                        acc[o + 3] = xadd(acc[o + 3], sqdiff(v[c].w, (float)creg[4 * c + 3]));
                    }
                }
                const float res = xadd(xadd(xadd(acc[0], acc[1]), xadd(acc[2], acc[3])),
                                       xadd(xadd(acc[4], acc[5]), xadd(acc[6], acc[7])));
                if (lane < rows) update_row<float, MODE_FPS, float>(p, res, m_old, 0.0, 0.0, r0 + lane, &best);
            } else {
                double acc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = 0.0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int o = (c & 1) * 4;
                    acc[o + 0] = fma((double)v[c].x, (double)creg[4 * c + 0], acc[o + 0]);
                    acc[o + 1] = fma((double)v[c].y, (double)creg[4 * c + 1], acc[o + 1]);
                    acc[o + 2] = fma((double)v[c].z, (double)creg[4 * c + 2], acc[o + 2]);
                    acc[o + 3] = fma((double)v[c].w, (double)creg[4 * c + 3], acc[o + 3]);
                }
                const double res = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                                             __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
                if (lane < rows) update_row<float, MODE_KCENTER, double>(p, res, m_old, xx_r, xx_c, r0 + lane, &best);
            }
            __syncwarp();  // every lane has read the stage before lane 0 re-arms it
        }
        if (!finish_pick<T>(p, best, step, lane, warp, NWARP, s_red, &s_next_center, &s_abort, s_center, &s_xx_c)) break;
    }
    long long outstanding = waits_left - loads_left;  // issued - consumed
    while (outstanding-- > 0) {
        mbar_wait(&bars[cur_stage], cur_phase);
        if (++cur_stage == S) {
            cur_stage = 0;
            cur_phase ^= 1u;
        }
    }
}

// squared row norms in fp64 (k-center): one warp per row, coalesced, FMA chain per lane + butterfly
template <typename T>
__global__ void row_norms_kernel(const T* __restrict__ F, unsigned long long N, int D, double* __restrict__ xx) {
    const unsigned long long row = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const int lane = threadIdx.x & 31;
    const T* a = F + row * (unsigned long long)D;
    double acc = 0.0;
    for (int i = lane; i < D; i += 32) {
        double v = (double)a[i];
        acc = fma(v, v, acc);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) xx[row] = acc;
}

template <typename T>
__global__ void fill_kernel(T* p, unsigned long long n, T v) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void fps_emit_kernel(const long long* picks, int first, int* out, unsigned long long n_samples) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_samples) out[i] = i == 0 ? first : (int)picks[i - 1];
}
__global__ void fps_emit_winners_kernel(const unsigned long long* winners, int first, int* out,
                                        unsigned long long n_samples) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_samples) out[i] = i == 0 ? first : (int)(0xFFFFFFFFu - (unsigned)(winners[2 * (i - 1)] & 0xFFFFFFFFull));
}
__global__ void kcenter_emit_kernel(const long long* picks, long long n_sel, long long* out, unsigned long long n_pick) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pick) return;
    if (n_sel == 0) out[i] = i == 0 ? 0 : picks[i - 1];
    else out[i] = picks[n_sel - 1 + i];
}

// ---- host side ------------------------------------------------------------------------------------------
static void plan_emit(LeafPlan& pl, int start, int n) {
    if (n <= 128) {
        pl.start[pl.n_leaves] = (unsigned short)start;
        pl.len[pl.n_leaves] = (unsigned short)n;
        pl.merges[pl.n_leaves] = 0;
        pl.n_leaves++;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    plan_emit(pl, start, n2);
    plan_emit(pl, start + n2, n - n2);
    pl.merges[pl.n_leaves - 1]++;
}

// workspace slots used by this module
enum { WS_MIND = 0, WS_XX = 1, WS_WIN = 2, WS_CAND = 3, WS_BAR = 4, WS_FORCED = 5, WS_PICKS = 6, WS_F = 7, WS_OUT = 8 };

template <typename T>
struct Launch {
    Params<T> p;
    int grid = 0;
    size_t smem = 0;
    void* fn = nullptr;
    bool lane_per_row = false;  // select32_kernel: second kernel argument is the tensor map
    bool partial_grid = false;  // virtual ranks sharing one device (tests): several small grids must run side by side
    CUtensorMap tmap;
};

// tunables (A/B measurements): SSDR_SEL_FEED=0 forces the cp.async ring, SSDR_SEL_STAGES / SSDR_SEL_WARPS override
// the ring depth and the warps per CTA of the TMA variants
struct Tunables {
    int feed = 1, stages = 0, warps = 0;
};
static unsigned long long peer_timeout_ms() {  // read per call: a test shortens it for one call
    const char* e = getenv("SSDR_PEER_TIMEOUT_MS");
    const unsigned long long v = e ? strtoull(e, nullptr, 10) : 0ull;
    return v > 0 ? v : 20000ull;
}
static Tunables tunables() {  // read per call (cheap): one process can sweep the knobs
    Tunables t;
    if (const char* e = getenv("SSDR_SEL_FEED")) t.feed = atoi(e) != 0;
    if (const char* e = getenv("SSDR_SEL_STAGES")) t.stages = atoi(e);
    if (const char* e = getenv("SSDR_SEL_WARPS")) t.warps = atoi(e);
    return t;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    });
    return fn;
}

template <typename K>
static int setup_kernel_fn(K kern, size_t smem, int threads, void** fn_out) {
    SSDR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    SSDR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
    SSDR_REQUIRE(nb >= 1, SSDR_ERR_CUDA, "selection kernel does not fit on an SM (smem %zu, %d threads)", smem, threads);
    *fn_out = (void*)kern;
    return SSDR_OK;
}

// D == 32 float32 rows, 16-byte aligned: the lane-per-row kernel over a swizzled tensor map
template <int MODE>
static int configure32(Ctx* c, Launch<float>& L, const float* dF, size_t N) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return -1;  // no driver entry point: the caller falls back to the bulk-copy variant
    const Tunables tn = tunables();
    Params<float>& p = L.p;
    // measured on B200 (tools/sweep_select.py, N = 500k, centre row staged once per CTA): 16 warps x 2 stages -- FPS
    // 7.9 us/pick (8 warps: 9.1, 4 warps: 12.7), k-center 11.1 us/pick (8 warps: 14.3); deeper rings change nothing
    int nw = tn.warps > 0 ? tn.warps : WARPS;
    int nst = tn.stages > 0 ? tn.stages : 2;
    nw = nw < 1 ? 1 : (nw > WARPS ? WARPS : nw);
    nst = nst < 2 ? 2 : (nst > MAX_STAGES ? MAX_STAGES : nst);
    const size_t budget = (size_t)c->max_smem_optin - 1024;
    auto need = [&](int st, int w) {
        return (size_t)1024 + (size_t)w * st * L32_UNIT_BYTES + WARPS * MAX_STAGES * 8 + WARPS * sizeof(Cand) + 128 + 64;
    };
    while (nst > 2 && need(nst, nw) > budget) --nst;
    while (nw > 1 && need(nst, nw) > budget) --nw;
    p.stride = 32;
    p.vec = 4;
    p.groups_per_unit = 8;
    p.nstages = nst;
    p.nwarps = nw;
    p.plan.n_leaves = 0;
    plan_emit(p.plan, 0, 32);
    L.smem = need(nst, nw);
    const cuuint64_t gdim[2] = {32, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&L.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dF), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SSDR_REQUIRE(r == CUDA_SUCCESS, SSDR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    L.lane_per_row = true;
    SSDR_TRY(setup_kernel_fn(select32_kernel<MODE>, L.smem, nw * 32, &L.fn));
    L.grid = c->sm_count;
    return SSDR_OK;
}

// Fill geometry (stride, unit rows, stages, warps, grid) for a (T, D) problem and pick the kernel variant.
template <typename T, int MODE>
static int configure(Ctx* c, Launch<T>& L, const T* dF, size_t N, size_t D) {
    Params<T>& p = L.p;
    const Tunables tn = tunables();
    p.F = dF;
    p.N = N;
    p.D = (int)D;
    const int vec_full = 16 / (int)sizeof(T);
    int vec = ((D % vec_full) == 0 && ((uintptr_t)dF % 16) == 0) ? vec_full : 1;
    if constexpr (sizeof(T) == 4) {
        if (D == 32 && vec == vec_full && tn.feed == 1 && N < 0x7FFFFFFFull) {
            int rc = configure32<MODE>(c, L, dF, N);
            if (rc >= 0) return rc;
        }
    }
    // conflict-free row stride: 8 lanes x 4 row groups must cover all 32 banks (float: stride%32 in {8,24};
    // double: stride%16 == 8, two half-warp phases)
    int stride = (int)D;
    for (;; stride += vec) {
        if (sizeof(T) == 4 && (stride % 32 == 8 || stride % 32 == 24)) break;
        if (sizeof(T) == 8 && (stride % 16 == 8)) break;
    }
    p.stride = stride;
    p.vec = vec;
    const size_t row_bytes = (size_t)stride * sizeof(T);
    const size_t budget = (size_t)c->max_smem_optin - 2048;
    const size_t fixed = align_up_dev((D + 8) * sizeof(T), 16) + align_up_dev(WARPS * sizeof(Cand), 16) +
                         align_up_dev((size_t)WARPS * MAX_STAGES * 8, 128) + 256;
    auto need = [&](int gr, int st, int w) { return fixed + (size_t)w * st * 4 * gr * row_bytes; };
    // fixed variants: (D, groups per unit) compiled in; units of 2-4 KB, 3 stages, 16 warps
    const bool fixed_ok = vec == vec_full && (D == 32 || D == 64 || D == 128 || D == 256);
    const bool bulk = fixed_ok && tn.feed == 1;
    int groups, nst = 3, nw = WARPS;
    if (fixed_ok) groups = D == 32 ? 4 : (D == 64 ? 2 : 1);
    else {
        groups = (int)(3072 / (4 * row_bytes));
        groups = groups < 1 ? 1 : (groups > 8 ? 8 : groups);
    }
    if (bulk) {  // measured (tools/sweep_select.py, N = 500k): 16 warps x 2 stages is best or equal for every fixed D
                 // (D = 256 FPS 78 us/pick, 8 warps 85; D = 64: 26.9 vs 43; deeper rings add nothing)
        nst = 2;
        if (tn.warps > 0) nw = tn.warps > WARPS ? WARPS : tn.warps;
        if (tn.stages > 0) nst = tn.stages > MAX_STAGES ? MAX_STAGES : (tn.stages < 2 ? 2 : tn.stages);
    }
    while (nst > 3 && need(groups, nst, nw) > budget) --nst;
    while (nw > 1 && need(groups, nst, nw) > budget) nw /= 2;
    while (nst > 2 && need(groups, nst, nw) > budget) --nst;
    while (!fixed_ok && groups > 1 && need(groups, nst, nw) > budget) --groups;
    SSDR_REQUIRE(need(groups, nst, nw) <= budget, SSDR_ERR_UNSUPPORTED,
                 "feature dimension D=%zu needs %zu bytes of shared memory per CTA (limit %zu)", D,
                 need(groups, nst, nw), budget);
    p.groups_per_unit = groups;
    p.nstages = nst;
    p.nwarps = nw;
    L.smem = need(groups, nst, nw);
    p.plan.n_leaves = 0;
    if (D >= 8) plan_emit(p.plan, 0, (int)D);
    const int threads = nw * 32;
#define SSDR_SEL(DT_, GT_, FEED_) SSDR_TRY(setup_kernel_fn(select_kernel<T, MODE, DT_, GT_, FEED_>, L.smem, threads, &L.fn))
    if (bulk) {
        if (D == 32) SSDR_SEL(32, 4, 1);
        else if (D == 64) SSDR_SEL(64, 2, 1);
        else if (D == 128) SSDR_SEL(128, 1, 1);
        else SSDR_SEL(256, 1, 1);
    } else {
        bool done = false;
        if constexpr (sizeof(T) == 4) {  // the cp.async ring with compiled-in geometry, kept for A/B runs (SSDR_SEL_FEED=0)
            if (fixed_ok && D == 32) {
                SSDR_SEL(32, 4, 0);
                done = true;
            } else if (fixed_ok && D == 256) {
                SSDR_SEL(256, 1, 0);
                done = true;
            }
        }
        if (!done) SSDR_SEL(0, 0, 0);
    }
#undef SSDR_SEL
    L.grid = c->sm_count;
    return SSDR_OK;
}

template <typename T, int MODE>
static int launch_steps(Launch<T>& L, int step_begin, int step_end, cudaStream_t s) {
    L.p.step_begin = step_begin;
    L.p.step_end = step_end;
    // mailbox tags are step+1 and unique per launch range, but a previous call may have left equal tags behind
    SSDR_CHECK_CUDA(cudaMemsetAsync(L.p.cand, 0, (size_t)2 * L.grid * sizeof(Mailbox), s));
    void* args[] = {(void*)&L.p, (void*)&L.tmap};
    if (L.partial_grid)  // co-residency comes from the grids being small (the spin waits are bounded by the time-out);
                         // cooperative launches of different streams do not overlap
        SSDR_CHECK_CUDA(cudaLaunchKernel(L.fn, dim3(L.grid), dim3(L.p.nwarps * 32), args, L.smem, s));
    else
        SSDR_CHECK_CUDA(cudaLaunchCooperativeKernel(L.fn, dim3(L.grid), dim3(L.p.nwarps * 32), args, L.smem, s));
    return SSDR_OK;
}

// Common set-up: workspaces, initial min-distance, forced centres.  d_forced may alias ws.
template <typename T, int MODE>
static int prepare(Ctx* c, Launch<T>& L, const T* dF, size_t N, size_t D, size_t row_begin, size_t row_end,
                   const long long* d_forced, int n_forced, int n_steps, cudaStream_t s, int max_ctas = 0) {
    SSDR_REQUIRE(N >= 1 && N < 0xFFFFFFFFull, SSDR_ERR_INVALID, "N=%zu out of range", N);
    SSDR_REQUIRE(D >= 1 && D <= 8192, SSDR_ERR_UNSUPPORTED, "feature dimension D=%zu not in [1, 8192]", D);
    SSDR_TRY((configure<T, MODE>(c, L, dF, N, D)));
    if (max_ctas > 0 && max_ctas < L.grid) {
        L.grid = max_ctas;
        L.partial_grid = true;
    }
    Params<T>& p = L.p;
    p.row_begin = row_begin;
    p.row_end = row_end;
    SSDR_TRY(c->ws[WS_MIND].reserve(N * sizeof(T)));
    SSDR_TRY(c->ws[WS_WIN].reserve((size_t)(n_steps > 0 ? n_steps : 1) * 16));
    SSDR_TRY(c->ws[WS_CAND].reserve((size_t)2 * L.grid * sizeof(Mailbox)));
    SSDR_TRY(c->ws[WS_BAR].reserve(256));
    SSDR_TRY(c->ws[WS_PICKS].reserve((size_t)(n_steps > 0 ? n_steps : 1) * sizeof(long long)));
    p.mind = c->ws[WS_MIND].as<T>();
    p.winners = c->ws[WS_WIN].as<unsigned long long>();
    p.cand = c->ws[WS_CAND].as<unsigned long long>();
    p.barrier = c->ws[WS_BAR].as<unsigned>();
    p.picks = c->ws[WS_PICKS].as<long long>();
    p.forced = d_forced;
    p.n_forced = n_forced;
    p.xx = nullptr;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    if (MODE == MODE_FPS) {
        fill_kernel<T><<<blocks, 256, 0, s>>>(p.mind, N, (T)1e10);  // fps_gcn_cpu.py:135
    } else {
        fill_kernel<T><<<blocks, 256, 0, s>>>(p.mind, N, (T)INFINITY);  // min_distances None => first dist wins
        SSDR_TRY(c->ws[WS_XX].reserve(N * sizeof(double)));
        p.xx = c->ws[WS_XX].as<double>();
        row_norms_kernel<T><<<(unsigned)((N + 7) / 8), 256, 0, s>>>(dF, N, (int)D, c->ws[WS_XX].as<double>());
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// ---- peer groups: the ranks of one node whose selection kernels exchange picks through each other's memory ----------
struct PeerGroup {
    int world = 1, rank = 0, device = 0;
    Mailbox* local = nullptr;          // cudaMalloc'ed: [2][world] mailboxes + error flag behind them
    Mailbox* boxes[MAX_PEERS] = {};    // every rank's block as mapped into this process
    bool opened[MAX_PEERS] = {};       // mapped with cudaIpcOpenMemHandle (to be closed)
    bool connected = false;
    unsigned tag_base = 0;             // advanced by every sharded call; all ranks make the same calls
    size_t block_bytes() const { return ((size_t)2 * world * sizeof(Mailbox) + 255) / 256 * 256 + 256; }
    unsigned* error_flag() const {
        return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(local) + block_bytes() - 256);
    }
};

static int peer_fill(PeerGroup* g, PeerParams* pp, int n_steps) {
    SSDR_REQUIRE(g && g->connected, SSDR_ERR_INVALID, "peer group is not connected");
    pp->world = g->world;
    pp->rank = g->rank;
    pp->tag_base = g->tag_base;
    for (int r = 0; r < g->world; ++r) pp->xbox[r] = g->boxes[r];
    pp->timeout_ns = peer_timeout_ms() * 1000000ull;
    pp->error = g->error_flag();
    g->tag_base += (unsigned)n_steps;
    return SSDR_OK;
}

template <typename T>
static int fps_dev(Ctx* c, const T* dF, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* d_out,
                   cudaStream_t s, size_t row_begin, size_t row_end, PeerGroup* g = nullptr, int max_ctas = 0) {
    SSDR_REQUIRE(dF && d_out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(row_begin <= row_end && row_end <= N, SSDR_ERR_INVALID, "bad row shard [%zu,%zu) of %zu", row_begin,
                 row_end, N);
    SSDR_REQUIRE(n_samples >= 1 && n_samples < (1ull << 31), SSDR_ERR_INVALID, "n_samples=%zu out of range", n_samples);
    SSDR_REQUIRE(first >= 0 && (size_t)first < N, SSDR_ERR_INVALID, "first index %d outside [0, %zu)", first, N);
    const int n_steps = (int)n_samples - 1;
    Launch<T> L;
    SSDR_TRY(c->ws[WS_FORCED].reserve(sizeof(long long)));
    long long f64 = first;
    SSDR_CHECK_CUDA(cudaMemcpyAsync(c->ws[WS_FORCED].p, &f64, sizeof(f64), cudaMemcpyHostToDevice, s));
    SSDR_TRY((prepare<T, MODE_FPS>(c, L, dF, N, D, row_begin, row_end, c->ws[WS_FORCED].as<long long>(), 1, n_steps, s,
                                   max_ctas)));
    if (g && g->world > 1) SSDR_TRY(peer_fill(g, &L.p.peer, n_steps));
    if (n_steps > 0) SSDR_TRY((launch_steps<T, MODE_FPS>(L, 0, n_steps, s)));
    fps_emit_kernel<<<(unsigned)((n_samples + 255) / 256), 256, 0, s>>>(L.p.picks, first, d_out, n_samples);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

template <typename T>
static int kcenter_dev(Ctx* c, const T* dX, size_t N, size_t D, const int64_t* d_sel, size_t n_sel, size_t n_pick,
                       int64_t* d_out, cudaStream_t s, size_t row_begin, size_t row_end, PeerGroup* g = nullptr,
                       int max_ctas = 0) {
    SSDR_REQUIRE(dX && (d_out || n_pick == 0) && (d_sel || n_sel == 0), SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(row_begin <= row_end && row_end <= N, SSDR_ERR_INVALID, "bad row shard [%zu,%zu) of %zu", row_begin,
                 row_end, N);
    SSDR_REQUIRE(n_sel + n_pick < (1ull << 31), SSDR_ERR_INVALID, "too many steps");
    if (n_pick == 0) return SSDR_OK;
    Launch<T> L;
    const long long* forced = reinterpret_cast<const long long*>(d_sel);
    int n_forced = (int)n_sel;
    if (n_sel == 0) {  // np.argmax(None) == 0 in the reference: the first pick is row 0
        SSDR_TRY(c->ws[WS_FORCED].reserve(sizeof(long long)));
        SSDR_CHECK_CUDA(cudaMemsetAsync(c->ws[WS_FORCED].p, 0, sizeof(long long), s));
        forced = c->ws[WS_FORCED].as<long long>();
        n_forced = 1;
    }
    // pick p is the winner of step n_sel-1+p (n_sel==0: pick 0 is row 0 itself, pick p>=1 the winner of step p-1)
    const int n_steps = n_sel == 0 ? (int)n_pick - 1 : (int)(n_sel + n_pick) - 1;
    SSDR_TRY((prepare<T, MODE_KCENTER>(c, L, dX, N, D, row_begin, row_end, forced, n_forced, n_steps, s, max_ctas)));
    if (g && g->world > 1) SSDR_TRY(peer_fill(g, &L.p.peer, n_steps));
    if (n_steps > 0) SSDR_TRY((launch_steps<T, MODE_KCENTER>(L, 0, n_steps, s)));
    kcenter_emit_kernel<<<(unsigned)((n_pick + 255) / 256), 256, 0, s>>>(L.p.picks, (long long)n_sel,
                                                                        reinterpret_cast<long long*>(d_out), n_pick);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// Row-sharded FPS with a host-enqueued collective (kept as the transport of last resort when peer memory cannot be
// mapped): one cooperative launch + one 8-byte NCCL max all-reduce per pick.
static int fps_sharded_dev(Ctx* c, const float* dF, size_t N, size_t D, size_t row_begin, size_t row_end,
                           int32_t first, size_t n_samples, int32_t* d_out, void* comm, cudaStream_t s) {
    SSDR_REQUIRE(dF && d_out && comm, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(row_begin <= row_end && row_end <= N, SSDR_ERR_INVALID, "bad row shard [%zu,%zu) of %zu", row_begin,
                 row_end, N);
    SSDR_REQUIRE(n_samples >= 1 && n_samples < (1ull << 31), SSDR_ERR_INVALID, "n_samples=%zu out of range", n_samples);
    SSDR_REQUIRE(first >= 0 && (size_t)first < N, SSDR_ERR_INVALID, "first index %d outside [0, %zu)", first, N);
    const int n_steps = (int)n_samples - 1;
    Launch<float> L;
    SSDR_TRY(c->ws[WS_FORCED].reserve(sizeof(long long)));
    long long f64 = first;
    SSDR_CHECK_CUDA(cudaMemcpyAsync(c->ws[WS_FORCED].p, &f64, sizeof(f64), cudaMemcpyHostToDevice, s));
    SSDR_TRY((prepare<float, MODE_FPS>(c, L, dF, N, D, row_begin, row_end, c->ws[WS_FORCED].as<long long>(), 1, n_steps, s)));
    for (int step = 0; step < n_steps; ++step) {
        SSDR_TRY((launch_steps<float, MODE_FPS>(L, step, step + 1, s)));
        SSDR_TRY(nccl_allreduce_max_u64(comm, L.p.winners + 2 * (size_t)step, 1, s));
    }
    fps_emit_winners_kernel<<<(unsigned)((n_samples + 255) / 256), 256, 0, s>>>(L.p.winners, first, d_out, n_samples);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// after a sharded call was enqueued and the stream synchronised: did a peer fail to answer?
static int peer_check(PeerGroup* g, cudaStream_t s) {
    unsigned e[40] = {};
    SSDR_CHECK_CUDA(cudaMemcpyAsync(e, g->error_flag(), sizeof(e), cudaMemcpyDeviceToHost, s));
    SSDR_CHECK_CUDA(cudaStreamSynchronize(s));
    if (e[0] == 0) return SSDR_OK;
    char tags[160] = "";
    size_t o = 0;
    for (int r = 0; r < g->world && r < 8 && o + 24 < sizeof(tags); ++r)
        o += (size_t)snprintf(tags + o, sizeof(tags) - o, " r%d:%u/%u", r, e[4 + 4 * r], e[6 + 4 * r]);
    return set_error(SSDR_ERR_CUDA,
                     "a peer rank did not post its pick within %llu ms (SSDR_PEER_TIMEOUT_MS): ranks out of step or a rank "
                     "died [rank %d step %u expects tag %u, CTA %u; tags in its block:%s]",
                     peer_timeout_ms(), g->rank, e[1], e[2], e[3], tags);
}

// SSDR_PEER_DEFER_CHECK=1: the sharded entry points return as soon as their kernel is enqueued and the caller asks
// ssdr_peer_group_check for the outcome (several virtual ranks of one process enqueue first, then wait)
static bool defer_check() {
    const char* e = getenv("SSDR_PEER_DEFER_CHECK");
    return e && e[0] == '1';
}

// host-pointer wrappers: stage in, run, copy picks out
template <typename T>
static int fps_host(const T* F, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* out) {
    SSDR_REQUIRE(F && out, SSDR_ERR_INVALID, "NULL pointer");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(c->ws[WS_F].reserve(N * D * sizeof(T)));
    SSDR_TRY(c->ws[WS_OUT].reserve(n_samples * sizeof(int32_t)));
    SSDR_TRY(h2d(c, c->ws[WS_F].p, F, N * D * sizeof(T), c->stream));
    SSDR_TRY(fps_dev<T>(c, c->ws[WS_F].as<T>(), N, D, first, n_samples, c->ws[WS_OUT].as<int32_t>(), c->stream, 0, N));
    return d2h_sync(c, out, c->ws[WS_OUT].p, n_samples * sizeof(int32_t), c->stream);
}
template <typename T>
static int kcenter_host(const T* X, size_t N, size_t D, const int64_t* sel, size_t n_sel, size_t n_pick, int64_t* out) {
    SSDR_REQUIRE(X && (out || !n_pick) && (sel || !n_sel), SSDR_ERR_INVALID, "NULL pointer");
    for (size_t i = 0; i < n_sel; ++i)
        SSDR_REQUIRE(sel[i] >= 0 && (size_t)sel[i] < N, SSDR_ERR_INVALID, "selected[%zu]=%lld outside [0, %zu)", i,
                     (long long)sel[i], N);
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(c->ws[WS_F].reserve(N * D * sizeof(T)));
    SSDR_TRY(c->ws[WS_OUT].reserve((n_pick + n_sel + 1) * sizeof(int64_t)));
    SSDR_TRY(h2d(c, c->ws[WS_F].p, X, N * D * sizeof(T), c->stream));
    int64_t* d_out = c->ws[WS_OUT].as<int64_t>();
    int64_t* d_sel = d_out + n_pick;
    SSDR_TRY(h2d(c, d_sel, sel, n_sel * sizeof(int64_t), c->stream));
    SSDR_TRY(kcenter_dev<T>(c, c->ws[WS_F].as<T>(), N, D, d_sel, n_sel, n_pick, d_out, c->stream, 0, N));
    return d2h_sync(c, out, d_out, n_pick * sizeof(int64_t), c->stream);
}

}  // namespace sel
}  // namespace ssdr

using namespace ssdr;

extern "C" {
int ssdr_fps_f32(const float* F, size_t N, size_t D, int32_t first, size_t n, int32_t* out) {
    return sel::fps_host<float>(F, N, D, first, n, out);
}
int ssdr_fps_f64(const double* F, size_t N, size_t D, int32_t first, size_t n, int32_t* out) {
    return sel::fps_host<double>(F, N, D, first, n, out);
}
int ssdr_fps_f32_dev(const float* F, size_t N, size_t D, int32_t first, size_t n, int32_t* out, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return sel::fps_dev<float>(c, F, N, D, first, n, out, (cudaStream_t)stream, 0, N);
}
int ssdr_fps_f64_dev(const double* F, size_t N, size_t D, int32_t first, size_t n, int32_t* out, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return sel::fps_dev<double>(c, F, N, D, first, n, out, (cudaStream_t)stream, 0, N);
}
int ssdr_fps_f32_sharded(const float* d_F, size_t N, size_t D, size_t row_begin, size_t row_end, int32_t first,
                         size_t n_samples, int32_t* d_out, void* nccl_comm, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return sel::fps_sharded_dev(c, d_F, N, D, row_begin, row_end, first, n_samples, d_out, nccl_comm,
                                (cudaStream_t)stream);
}
int ssdr_kcenter_f32(const float* X, size_t N, size_t D, const int64_t* sel_, size_t n_sel, size_t n_pick, int64_t* out) {
    return sel::kcenter_host<float>(X, N, D, sel_, n_sel, n_pick, out);
}
int ssdr_kcenter_f64(const double* X, size_t N, size_t D, const int64_t* sel_, size_t n_sel, size_t n_pick, int64_t* out) {
    return sel::kcenter_host<double>(X, N, D, sel_, n_sel, n_pick, out);
}
int ssdr_kcenter_f32_dev(const float* X, size_t N, size_t D, const int64_t* sel_, size_t n_sel, size_t n_pick,
                         int64_t* out, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return sel::kcenter_dev<float>(c, X, N, D, sel_, n_sel, n_pick, out, (cudaStream_t)stream, 0, N);
}
int ssdr_kcenter_f64_dev(const double* X, size_t N, size_t D, const int64_t* sel_, size_t n_sel, size_t n_pick,
                         int64_t* out, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return sel::kcenter_dev<double>(c, X, N, D, sel_, n_sel, n_pick, out, (cudaStream_t)stream, 0, N);
}

// ---- peer groups ---------------------------------------------------------------------------------------------------
int ssdr_peer_group_create(int world, int rank, void** group) {
    SSDR_REQUIRE(group, SSDR_ERR_INVALID, "group is NULL");
    SSDR_REQUIRE(world >= 1 && world <= sel::MAX_PEERS && rank >= 0 && rank < world, SSDR_ERR_INVALID,
                 "bad peer group geometry: world %d (max %d), rank %d", world, sel::MAX_PEERS, rank);
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    sel::PeerGroup* g = new sel::PeerGroup();
    g->world = world;
    g->rank = rank;
    g->device = c->device;
    // plain cudaMalloc (not the stream-ordered pool): the block is exported with cudaIpcGetMemHandle
    cudaError_t e = cudaMalloc((void**)&g->local, g->block_bytes());
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete g;
        return set_error(SSDR_ERR_NOMEM, "cudaMalloc of the peer mailbox block failed: %s", cudaGetErrorString(e));
    }
    e = cudaMemset(g->local, 0, g->block_bytes());
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(g->local);
        delete g;
        return set_error(SSDR_ERR_CUDA, "clearing the peer mailbox block failed: %s", cudaGetErrorString(e));
    }
    g->boxes[rank] = g->local;
    g->connected = world == 1;
    *group = g;
    return SSDR_OK;
}
int ssdr_peer_group_export(void* group, void* handle64) {
    SSDR_REQUIRE(group && handle64, SSDR_ERR_INVALID, "NULL pointer");
    sel::PeerGroup* g = (sel::PeerGroup*)group;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    SSDR_CHECK_CUDA(cudaIpcGetMemHandle(&h, g->local));
    memcpy(handle64, &h, sizeof(h));
    return SSDR_OK;
}
int ssdr_peer_group_connect(void* group, const void* handles) {
    SSDR_REQUIRE(group && handles, SSDR_ERR_INVALID, "NULL pointer");
    sel::PeerGroup* g = (sel::PeerGroup*)group;
    for (int r = 0; r < g->world; ++r) {
        if (r == g->rank || g->boxes[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * 64, sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(SSDR_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle for rank %d failed: %s (no peer access?)", r,
                             cudaGetErrorString(e));
        }
        g->boxes[r] = (sel::Mailbox*)p;
        g->opened[r] = true;
    }
    g->connected = true;
    return SSDR_OK;
}
int ssdr_peer_group_connect_local(void** groups, int world) {
    SSDR_REQUIRE(groups && world >= 1 && world <= sel::MAX_PEERS, SSDR_ERR_INVALID, "bad argument");
    for (int a = 0; a < world; ++a) {
        sel::PeerGroup* g = (sel::PeerGroup*)groups[a];
        SSDR_REQUIRE(g && g->world == world && g->rank == a, SSDR_ERR_INVALID, "groups[%d] is not rank %d of %d", a, a, world);
    }
    for (int a = 0; a < world; ++a) {
        sel::PeerGroup* g = (sel::PeerGroup*)groups[a];
        for (int r = 0; r < world; ++r) g->boxes[r] = ((sel::PeerGroup*)groups[r])->local;
        g->connected = true;
    }
    return SSDR_OK;
}
int ssdr_peer_group_destroy(void* group) {
    if (!group) return SSDR_OK;
    sel::PeerGroup* g = (sel::PeerGroup*)group;
    for (int r = 0; r < g->world; ++r)
        if (g->opened[r]) cudaIpcCloseMemHandle(g->boxes[r]);
    if (g->local) cudaFree(g->local);
    cudaGetLastError();
    delete g;
    return SSDR_OK;
}

int ssdr_fps_sharded_p2p(int dtype, const void* d_F, size_t N, size_t D, size_t row_begin, size_t row_end, int32_t first,
                         size_t n_samples, int32_t* d_out, void* group, void* stream, int max_ctas) {
    SSDR_REQUIRE(group, SSDR_ERR_INVALID, "group is NULL");
    SSDR_REQUIRE(dtype == SSDR_F32 || dtype == SSDR_F64, SSDR_ERR_INVALID, "dtype must be SSDR_F32 or SSDR_F64");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    sel::PeerGroup* g = (sel::PeerGroup*)group;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSDR_F32)
        SSDR_TRY(sel::fps_dev<float>(c, (const float*)d_F, N, D, first, n_samples, d_out, s, row_begin, row_end, g, max_ctas));
    else
        SSDR_TRY(sel::fps_dev<double>(c, (const double*)d_F, N, D, first, n_samples, d_out, s, row_begin, row_end, g, max_ctas));
    return (g->world > 1 && !sel::defer_check()) ? sel::peer_check(g, s) : SSDR_OK;
}
int ssdr_peer_group_check(void* group, void* stream) {
    SSDR_REQUIRE(group, SSDR_ERR_INVALID, "group is NULL");
    sel::PeerGroup* g = (sel::PeerGroup*)group;
    return g->world > 1 ? sel::peer_check(g, (cudaStream_t)stream) : SSDR_OK;
}
int ssdr_kcenter_sharded_p2p(int dtype, const void* d_X, size_t N, size_t D, size_t row_begin, size_t row_end,
                             const int64_t* d_selected, size_t n_sel, size_t n_pick, int64_t* d_out, void* group,
                             void* stream, int max_ctas) {
    SSDR_REQUIRE(group, SSDR_ERR_INVALID, "group is NULL");
    SSDR_REQUIRE(dtype == SSDR_F32 || dtype == SSDR_F64, SSDR_ERR_INVALID, "dtype must be SSDR_F32 or SSDR_F64");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    sel::PeerGroup* g = (sel::PeerGroup*)group;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSDR_F32)
        SSDR_TRY(sel::kcenter_dev<float>(c, (const float*)d_X, N, D, d_selected, n_sel, n_pick, d_out, s, row_begin, row_end, g, max_ctas));
    else
        SSDR_TRY(sel::kcenter_dev<double>(c, (const double*)d_X, N, D, d_selected, n_sel, n_pick, d_out, s, row_begin, row_end, g, max_ctas));
    return (g->world > 1 && !sel::defer_check()) ? sel::peer_check(g, s) : SSDR_OK;
}
}
