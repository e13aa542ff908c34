// grid.cu -- voxel grid (barycentre) subsampling on sm_100a.
//
// Replaces grid_subsampling() (utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106),
// a single-threaded unordered_map loop, by a sort + segmented sequential reduce that reproduces the reference
// bit for bit.  FIVE launches, no host round trip in between (the host reads the voxel count once, at the end):
//
//   pack_rows_kernel  xyz | features | labels of every point packed into ONE record (32 or 16 bytes for the two layouts
//                     of the reference callers), so that a gathered point costs one sector instead of one or two per array
//   sort_kernel   (persistent, cooperative, one CTA per SM; phases separated by a grid barrier)
//     P0  min/max corners (cloud.cpp:27-67) -> origin = floor(min * (1/dl)) * dl, nX, nY (grid_subsampling.cpp:27-31)
//     P1  key = iX + nX*iY + nX*nY*iZ with i* = floor((p-origin)/dl) in IEEE fp32, true division
//         (grid_subsampling.cpp:53-56); points stream through shared memory as float4, the digit histogram of the
//         first radix pass is taken on the fly.  Slab mode (multi-GPU): ONE read of the cloud through per-warp cp.async
//         rings, members compacted in input order and packed into records at their compacted position
//     P2  STABLE LSD radix sort of (key, position), 8 bits per pass, over the significant key bits only -- the number
//         of passes is decided ON THE DEVICE from the largest key.  Per pass: per-CTA digit histogram, barrier, every
//         CTA derives its digit offsets from the CTA-major histogram matrix, stable scatter of its contiguous chunk
//         (same-digit lane masks keep equal keys in input order; every round is put in digit order in shared memory
//         and written out from there), barrier
//     P3  segment heads -> voxel start offsets, M (warp-contiguous runs, ballot ranks)
//   reduce_kernel (M known only on the device: grid-stride over voxels)
//         eight lanes per voxel: the lanes gather up to eight points of the voxel in parallel (index, xyz, features,
//         label) into shared memory, then every lane owns ONE CHANNEL (x, y, z, feature j) and adds the staged values
//         IN INPUT ORDER (the stable sort keeps ascending index inside a voxel) exactly like SampledData::update_*
//         (grid_subsampling.h:42-79); bary = sum * float(1.0/count), feat = sum / float(count)
//         (grid_subsampling.cpp:87-95); the label vote counts in a per-voxel shared-memory table that keeps
//         first-occurrence order, ties resolved by libstdc++'s unordered_map iteration order
//         (grid_subsampling.cpp:97-102)
//   reduce_small_kernel / reduce_rec_kernel: the same arithmetic for the reference layouts read from packed records --
//         one THREAD per small single-label voxel first, the eight-lane groups take what it marks
//   reduce_heavy_kernel: the voxels of more than 1024 points that the groups passed over, one CTA each
//   optional (SSDR_GRID_ORDER_REFERENCE): rows permuted into the reference's libstdc++ hash-iteration order.
// Rows come out in ascending voxel-key order by default (SSDR_GRID_ORDER_KEY).
#include <stdlib.h>

#include "common.cuh"
#include "primitives.cuh"

namespace ssdr {
namespace grid {

struct Meta {
    float mn[3], mx[3];
    float origin[3];
    float dl;
    unsigned long long nX, nY, nZ;
    int key_bits;
    int error;  // 2: more than LABEL_CAP distinct labels in one voxel
    unsigned long long max_key;  // largest key seen -> number of radix bits worth sorting
    unsigned long long n_sel;    // points taking part (all of them, or the ones inside the requested slab)
    unsigned long long M;
    int cur;                     // which ping-pong pair holds the sorted (key, index) arrays
    unsigned n_heavy;            // voxels deferred to reduce_heavy_kernel (zeroed with the rest by the sort kernel)
    unsigned n_deferred;         // voxels reduce_small_kernel left to the eight-lane groups (same)
    // %globaltimer marks (ns) of CTA 0: [0] start, [1] geometry, [2] keys, [3..10] end of radix pass k, [11] heads
    // counted, [12] starts written, [13] reduce start (first CTA), [14] reduce end (last CTA)
    unsigned long long tmark[16];
};

// slots 16.. belong to kdtree.cuh (its tree workspaces must survive between KNN calls): never used here
enum { WS_META = 0, WS_PART = 1, WS_KEYS = 2, WS_KEYS2 = 3, WS_IDX = 4, WS_IDX2 = 5, WS_TEMP = 6, WS_STARTS = 7,
       WS_CTL = 8, WS_IN_P = 9, WS_IN_F = 10, WS_IN_C = 11, WS_HASH = 12, WS_CMIN = 13, WS_HEADS = 14, WS_REC = 15 };

constexpr int LABEL_CAP = 64;
constexpr int MM_BLOCK = 256;

// float -> size_t the way x86-64 gcc does it for |v| < 2^63 (truncate to int64, reinterpret)
__device__ __forceinline__ unsigned long long f2u64(float v) { return (unsigned long long)(long long)v; }

static int bits_for_value(unsigned long long v) {  // radix bits needed to order values <= v
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}
__device__ __forceinline__ int dev_bits_for_value(unsigned long long v) { return v ? 64 - __clzll((long long)v) : 1; }

struct Geo {
    float mn[3], mx[3], origin[3];
    float dl;
    unsigned long long nX, nY, nZ;
};
// origin / grid size from the corners (grid_subsampling.cpp:27-31), the same IEEE operations as the reference
__device__ __forceinline__ void geometry_from_corners(const float* mn, const float* mx, float dl, Geo* g) {
    const float inv = __fdiv_rn(1.0f, dl);
    for (int d = 0; d < 3; ++d) {
        g->mn[d] = mn[d];
        g->mx[d] = mx[d];
        g->origin[d] = __fmul_rn(floorf(__fmul_rn(mn[d], inv)), dl);
    }
    g->dl = dl;
    g->nX = f2u64(floorf(__fdiv_rn(__fsub_rn(mx[0], g->origin[0]), dl))) + 1ull;
    g->nY = f2u64(floorf(__fdiv_rn(__fsub_rn(mx[1], g->origin[1]), dl))) + 1ull;
    g->nZ = f2u64(floorf(__fdiv_rn(__fsub_rn(mx[2], g->origin[2]), dl))) + 1ull;
}
// Keys are computed in wrapping 64-bit arithmetic exactly like the reference's size_t expression, so even
// degenerate inputs (a point rounding to one cell below the origin, overflowing nX*nY*nZ) group identically.
__device__ __forceinline__ unsigned long long voxel_key(const Geo& g, float x, float y, float z) {
    const unsigned long long iX = f2u64(floorf(__fdiv_rn(__fsub_rn(x, g.origin[0]), g.dl)));
    const unsigned long long iY = f2u64(floorf(__fdiv_rn(__fsub_rn(y, g.origin[1]), g.dl)));
    const unsigned long long iZ = f2u64(floorf(__fdiv_rn(__fsub_rn(z, g.origin[2]), g.dl)));
    return iX + g.nX * iY + g.nX * g.nY * iZ;
}
__device__ __forceinline__ unsigned long long voxel_layer(const Geo& g, float v, int axis) {
    return f2u64(floorf(__fdiv_rn(__fsub_rn(v, g.origin[axis]), g.dl)));
}

// ---- stand-alone min / max and geometry (helpers of the multi-GPU slab logic: bbox of a chunk, layer of each point) -
__global__ void minmax_kernel(const float* __restrict__ pts, unsigned long long N, float* __restrict__ partials) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = __ldg(pts + 3 * i + d);
            mn[d] = v < mn[d] ? v : mn[d];
            mx[d] = v > mx[d] ? v : mx[d];
        }
    }
    __shared__ float s[6][MM_BLOCK / 32];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int d = 0; d < 3; ++d) {
            s[d][warp] = mn[d];
            s[3 + d][warp] = mx[d];
        }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[threadIdx.x][0];
        for (int w = 1; w < MM_BLOCK / 32; ++w)
            v = threadIdx.x < 3 ? fminf(v, s[threadIdx.x][w]) : fmaxf(v, s[threadIdx.x][w]);
        partials[blockIdx.x * 6 + threadIdx.x] = v;
    }
}
__global__ void setup_kernel(const float* __restrict__ partials, int nparts, float dl, Meta* meta) {
    // one warp per quantity (3 mins, 3 maxes): strided loads + shuffle reduction
    const int t = threadIdx.x, q = t >> 5, lane = t & 31;
    __shared__ float r[6];
    if (q < 6) {
        float v = q < 3 ? INFINITY : -INFINITY;
        for (int b = lane; b < nparts; b += 32) {
            const float x = partials[b * 6 + q];
            v = q < 3 ? fminf(v, x) : fmaxf(v, x);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, v, m);
            v = q < 3 ? fminf(v, o) : fmaxf(v, o);
        }
        if (lane == 0) r[q] = v;
    }
    __syncthreads();
    if (t == 0) {
        Geo g;
        geometry_from_corners(r, r + 3, dl, &g);
        Meta m;
        for (int d = 0; d < 3; ++d) {
            m.mn[d] = g.mn[d];
            m.mx[d] = g.mx[d];
            m.origin[d] = g.origin[d];
        }
        m.dl = dl;
        m.nX = g.nX;
        m.nY = g.nY;
        m.nZ = g.nZ;
        m.error = 0;
        m.key_bits = 64;
        m.max_key = 0;
        m.n_sel = 0;
        m.M = 0;
        m.cur = 0;
        m.n_heavy = 0;
        m.n_deferred = 0;
        for (int k = 0; k < 16; ++k) m.tmark[k] = 0;
        *meta = m;
    }
}
__global__ void point_layers_kernel(const float* __restrict__ pts, unsigned long long N,
                                    const Meta* __restrict__ meta, int axis, int* __restrict__ layers) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned long long layer =
        f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + axis), meta->origin[axis]), meta->dl)));
    layers[i] = layer < 0x7FFFFFFFull ? (int)layer : 0x7FFFFFFF;
}
// histogram of the voxel layers along one axis (slab balancing of a multi-GPU job).  A scan has a few hundred layers
// and most of its points in a handful of them (the ground), so the counts are taken in shared memory per CTA
// (warp-aggregated) and flushed once; grids with more layers than fit there count straight into global memory.
constexpr unsigned HIST_SMEM_LAYERS = 8192;
__global__ void __launch_bounds__(256) layer_hist_kernel(const float* __restrict__ pts, unsigned long long N,
                                                         const Meta* __restrict__ meta, int axis,
                                                         unsigned long long* __restrict__ hist, unsigned long long n_layers,
                                                         unsigned long long sample_stride) {
    __shared__ unsigned s_h[HIST_SMEM_LAYERS];
    const bool local = n_layers <= HIST_SMEM_LAYERS;
    if (local) {
        for (unsigned i = threadIdx.x; i < (unsigned)n_layers; i += blockDim.x) s_h[i] = 0;
        __syncthreads();
    }
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long ns = (N + sample_stride - 1) / sample_stride;  // every sample_stride-th point is counted
    const unsigned long long rounds = (ns + stride - 1) / stride;
    const float o = meta->origin[axis], dl = meta->dl;
    for (unsigned long long r = 0; r < rounds; ++r) {
        const unsigned long long j = r * stride + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        const unsigned long long i = j < ns ? j * sample_stride : N;
        unsigned long long layer = ~0ull;
        if (i < N) {
            layer = f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + axis), o), dl)));
            if (layer >= n_layers) layer = n_layers - 1;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, layer);
        if (i < N && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) {
            if (local) atomicAdd(&s_h[layer], (unsigned)__popc(peers));
            else atomicAdd(&hist[layer], (unsigned long long)__popc(peers));
        }
    }
    if (local) {
        __syncthreads();
        for (unsigned i = threadIdx.x; i < (unsigned)n_layers; i += blockDim.x)
            if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
    }
}
// destination rank of every point from its layer and the slab bounds (bounds[r] <= layer < bounds[r+1] -> r); written as
// the sort key of a one-pass stable radix sort, together with the identity permutation; per-destination counts
constexpr int MAX_ROUTE_WORLD = 64;
struct RouteBounds {
    unsigned long long b[MAX_ROUTE_WORLD + 1];
    int world;
};
__global__ void route_key_kernel(const float* __restrict__ pts, unsigned long long N, const Meta* __restrict__ meta,
                                 int axis, const RouteBounds rb, unsigned long long* __restrict__ keys,
                                 unsigned* __restrict__ idx, unsigned long long* __restrict__ counts) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned dest = 0xFFFFFFFFu;
    if (i < N) {
        const unsigned long long layer =
            f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + axis), meta->origin[axis]), meta->dl)));
        dest = 0;
        for (int r = 1; r < rb.world; ++r) dest += layer >= rb.b[r] ? 1u : 0u;  // bounds are non-decreasing
        keys[i] = dest;
        idx[i] = (unsigned)i;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, dest);
    if (i < N && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&counts[dest], (unsigned long long)__popc(peers));
}
__global__ void bbox_fill_kernel(float* partials, float a0, float a1, float a2, float b0, float b1, float b2) {
    partials[0] = a0; partials[1] = a1; partials[2] = a2;
    partials[3] = b0; partials[4] = b1; partials[5] = b2;
}

// ---- the persistent sort kernel -------------------------------------------------------------------------------------
constexpr int PA_THREADS = 1024;
constexpr int PA_WARPS = PA_THREADS / 32;
constexpr int BINS = 256;
constexpr int SC_ITEMS = 4;  // elements per thread and scatter round

struct SortParams {
    const float* pts;
    unsigned long long N;
    float dl;
    int has_bbox;
    float bbox[6];
    int slab_axis;  // -1: the whole cloud takes part
    unsigned long long slab_lo, slab_hi;
    Meta* meta;
    unsigned long long* keys[2];
    unsigned* idx[2];
    unsigned* hist;             // [G][BINS] CTA-major digit histograms of the current pass
    float* partials;            // [G][6]
    unsigned long long* pmax;   // [G] largest key per CTA
    unsigned* cta_count;        // [G] segment heads per CTA
    unsigned* wcount;           // [G][PA_WARPS] slab members staged by each warp
    // slab members packed into records at their compacted position (rec_bytes = 32: xyz | 3 x float32 | int32 label,
    // 16: xyz | 3 x uint8 | uint8 label; 0: none -- the reduce then gathers from the caller's arrays by input index)
    const void* feats;
    const void* cls;
    unsigned char* rec;
    int rec_bytes;
    unsigned* starts;           // [N + 1] voxel start offsets
    unsigned* barrier;          // monotonic arrival counter, zero at launch
};


// Lanes of the warp holding the same radix digit (0 .. BINS; BINS = "no element"): nine ballots and a few logic ops,
// a fixed cost, where match.any takes one round per DISTINCT value -- 32 rounds on the random low digits of a key.
__device__ __forceinline__ unsigned match_digit(unsigned digit) {
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
        const bool on = (digit >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        peers &= on ? bal : ~bal;
    }
    return peers;
}
constexpr int SLAB_STAGES = 3;  // quads of 128 points per warp in the slab pass's cp.async ring
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// All CTAs of the (cooperative, hence co-resident) grid meet here.  The counter only grows: barrier k is passed when it
// reaches k * G.  __threadfence after the wait also drops stale L1 lines before the next phase reads other CTAs' data.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned G, unsigned* target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        *target += G;
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_gpu_u32(counter) < *target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// exclusive scan of v over the 1024 threads of the CTA; *total = sum.  s_w: PA_WARPS + 1 words of shared memory.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* s_w, unsigned* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();  // s_w may still be read from a previous call
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = s_w[lane];
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        s_w[lane] = wi - w;
        if (lane == 31) s_w[PA_WARPS] = wi;
    }
    __syncthreads();
    *total = s_w[PA_WARPS];
    return s_w[warp] + inc - v;
}

// One tile = PA_THREADS consecutive points staged through shared memory as float4 (coalesced 16-byte loads; the
// 12-byte records are then read back with stride 3 words, which is conflict free).  The NEXT tile's float4 is already in
// a register while the current one is consumed, so the global-load latency overlaps the work on the tile.
// base % 4 == 0 and pts 16-byte aligned are guaranteed by the caller; otherwise the scalar path loads the three floats.
struct TileLoader {
    const float* pts;
    bool aligned;
    float* s_xyz;
    float4 pre;
    float sx, sy, sz;  // scalar path: this thread's point of the prefetched tile

    __device__ __forceinline__ void prefetch(unsigned long long base, unsigned cnt) {
        const unsigned t = threadIdx.x;
        if (aligned) {
            const unsigned nv = (cnt * 3u) >> 2;
            if (t < nv) pre = __ldg(reinterpret_cast<const float4*>(pts + base * 3ull) + t);
        } else if (t < cnt) {
            const float* q = pts + (base + t) * 3ull;
            sx = __ldg(q);
            sy = __ldg(q + 1);
            sz = __ldg(q + 2);
        }
    }
    // hands out the prefetched tile (base, cnt) and starts the load of the next one (nbase, ncnt; ncnt == 0: none)
    __device__ __forceinline__ void next(unsigned long long base, unsigned cnt, unsigned long long nbase, unsigned ncnt,
                                         float* x, float* y, float* z) {
        const unsigned t = threadIdx.x;
        if (aligned) {
            __syncthreads();  // the previous tile has been consumed
            const unsigned nfl = cnt * 3u, nv = nfl >> 2;
            if (t < nv) reinterpret_cast<float4*>(s_xyz)[t] = pre;
            if (t < (nfl & 3u)) s_xyz[(nv << 2) + t] = __ldg(pts + base * 3ull + (nv << 2) + t);
            if (ncnt) prefetch(nbase, ncnt);
            __syncthreads();
            if (t < cnt) {
                *x = s_xyz[3 * t];
                *y = s_xyz[3 * t + 1];
                *z = s_xyz[3 * t + 2];
            }
        } else {
            *x = sx;
            *y = sy;
            *z = sz;
            if (ncnt) prefetch(nbase, ncnt);
        }
    }
};

__global__ void __launch_bounds__(PA_THREADS, 1) sort_kernel(const SortParams p) {
    __shared__ __align__(16) unsigned s_cnt[PA_WARPS][BINS];     // radix ranking counters (P2) ...
    float* s_xyz = reinterpret_cast<float*>(&s_cnt[0][0]);       // ... and the float4 staging tile of P0 / P1
    __shared__ unsigned s_base[BINS], s_tot[BINS], s_pre[BINS];
    __shared__ unsigned s_w[PA_WARPS + 1];
    __shared__ float s_red[6][PA_WARPS];
    __shared__ unsigned long long s_red64[PA_WARPS];
    __shared__ Geo s_geo;
    __shared__ unsigned long long s_n, s_maxkey;
    __shared__ unsigned s_bar_target;
    __shared__ unsigned long long s_mark[16];
    extern __shared__ __align__(16) unsigned char s_dyn[];  // one scatter round in digit order: keys, then values
    unsigned long long* s_rkey = reinterpret_cast<unsigned long long*>(s_dyn);
    unsigned* s_rval = reinterpret_cast<unsigned*>(s_dyn + (size_t)SC_ITEMS * PA_THREADS * sizeof(unsigned long long));
    const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned G = gridDim.x, c = blockIdx.x;
    if (t == 0) s_bar_target = 0;
    if (t < 16) s_mark[t] = 0;
    __syncthreads();
    if (t == 0) s_mark[0] = gtimer_ns();
    const bool aligned = (reinterpret_cast<unsigned long long>(p.pts) & 15ull) == 0;
    // contiguous chunk of all N points for this CTA (tile aligned)
    const unsigned long long perN = (((p.N + G - 1) / G) + PA_THREADS - 1) / PA_THREADS * PA_THREADS;
    const unsigned long long cb = min((unsigned long long)c * perN, p.N), ce = min(cb + perN, p.N);
    TileLoader tl;
    tl.pts = p.pts;
    tl.aligned = aligned;
    tl.s_xyz = s_xyz;

    // ---- P0: corners -> geometry (identical in every CTA)
    if (!p.has_bbox) {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        if (cb < ce) tl.prefetch(cb, (unsigned)min((unsigned long long)PA_THREADS, ce - cb));
        for (unsigned long long b = cb; b < ce; b += PA_THREADS) {
            const unsigned cnt = (unsigned)min((unsigned long long)PA_THREADS, ce - b);
            const unsigned long long nb = b + PA_THREADS;
            float x = 0.f, y = 0.f, z = 0.f;
            tl.next(b, cnt, nb, nb < ce ? (unsigned)min((unsigned long long)PA_THREADS, ce - nb) : 0u, &x, &y, &z);
            if (t < cnt) {
                mn[0] = x < mn[0] ? x : mn[0];
                mx[0] = x > mx[0] ? x : mx[0];
                mn[1] = y < mn[1] ? y : mn[1];
                mx[1] = y > mx[1] ? y : mx[1];
                mn[2] = z < mn[2] ? z : mn[2];
                mx[2] = z > mx[2] ? z : mx[2];
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
                mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
            }
        }
        if (lane == 0)
            for (int d = 0; d < 3; ++d) {
                s_red[d][warp] = mn[d];
                s_red[3 + d][warp] = mx[d];
            }
        __syncthreads();
        if (t < 6) {
            float v = s_red[t][0];
            for (int w = 1; w < PA_WARPS; ++w) v = t < 3 ? fminf(v, s_red[t][w]) : fmaxf(v, s_red[t][w]);
            p.partials[c * 6 + t] = v;
        }
        grid_barrier(p.barrier, G, &s_bar_target);
        if (warp < 6) {  // one warp per quantity
            float v = warp < 3 ? INFINITY : -INFINITY;
            for (unsigned b = lane; b < G; b += 32) {
                const float o = __ldcg(p.partials + b * 6 + warp);
                v = warp < 3 ? fminf(v, o) : fmaxf(v, o);
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                const float o = __shfl_xor_sync(0xffffffffu, v, m);
                v = warp < 3 ? fminf(v, o) : fmaxf(v, o);
            }
            if (lane == 0) s_red[warp][0] = v;
        }
        __syncthreads();
    } else if (t < 6) {
        s_red[t][0] = p.bbox[t];
    }
    __syncthreads();
    if (t == 0) {
        float mn[3] = {s_red[0][0], s_red[1][0], s_red[2][0]}, mx[3] = {s_red[3][0], s_red[4][0], s_red[5][0]};
        geometry_from_corners(mn, mx, p.dl, &s_geo);
    }
    __syncthreads();
    const Geo geo = s_geo;
    const bool slab = p.slab_axis >= 0;
    if (t == 0) s_mark[1] = gtimer_ns();

    // ---- P1: keys, digit histogram of the first pass, largest key
    unsigned long long kmax = 0;
    if (slab) {
        // Slab members are compacted IN INPUT ORDER with ONE read of the cloud and no block barrier per tile: every warp
        // owns a contiguous run of this CTA's chunk, streams it as coalesced words through its own shared-memory tile
        // (the next two tiles are already in registers), tests the slab layer (one division per point), computes the
        // key of the members only and stages (key, index) at the head of its own run inside the SECOND sort buffers
        // (a run never holds more members than points).  After the barrier every warp knows where its members start
        // in the compacted array and copies them over.
        const unsigned long long wlen = perN / PA_WARPS;  // a multiple of 32
        const unsigned long long wb = min(cb + (unsigned long long)warp * wlen, ce), we = min(wb + wlen, ce);
        // the warp's ring in dynamic shared memory: SLAB_STAGES quads of 128 points, filled by cp.async (16-byte
        // chunks when the cloud is 16-byte aligned), so SLAB_STAGES - 1 quads are in flight while one is worked on
        float* ring = reinterpret_cast<float*>(s_dyn) + (size_t)warp * (SLAB_STAGES * 384);
        const unsigned long long wend = 3ull * we;
        const unsigned nq = (unsigned)((we - wb + 127) / 128);
        auto issue = [&](unsigned q) {
            if (q < nq) {
                const unsigned long long b0 = wb + 128ull * q;
                float* dst = ring + (q % SLAB_STAGES) * 384;
                if (aligned && b0 + 128 <= we) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) cp_async_16(dst + 4 * (lane + 32 * k), p.pts + 3ull * b0 + 4 * (lane + 32 * k));
                } else {
#pragma unroll
                    for (int k = 0; k < 12; ++k) {
                        const unsigned long long wi = 3ull * b0 + lane + 32u * (unsigned)k;
                        if (wi < wend) cp_async_4(dst + lane + 32 * k, p.pts + wi);
                    }
                }
            }
            cp_async_commit();
        };
        unsigned run = 0;
#pragma unroll
        for (int q = 0; q < SLAB_STAGES - 1; ++q) issue((unsigned)q);
        for (unsigned q = 0; q < nq; ++q) {
            issue(q + SLAB_STAGES - 1);
            cp_async_wait<SLAB_STAGES - 1>();
            __syncwarp();
            const float* tile = ring + (q % SLAB_STAGES) * 384;
            const unsigned long long b = wb + 128ull * q;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long i = b + 32u * (unsigned)u + lane;
                bool member = false;
                unsigned long long key = 0;
                if (i < we) {
                    const float x = tile[96 * u + 3 * lane], y = tile[96 * u + 3 * lane + 1], z = tile[96 * u + 3 * lane + 2];
                    const unsigned long long layer =
                        voxel_layer(geo, p.slab_axis == 0 ? x : (p.slab_axis == 1 ? y : z), p.slab_axis);
                    member = layer >= p.slab_lo && layer < p.slab_hi;
                    if (member) key = voxel_key(geo, x, y, z);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, member);
                if (member) {
                    const unsigned long long pos = wb + run + __popc(bal & ((1u << lane) - 1u));
                    p.keys[1][pos] = key;
                    p.idx[1][pos] = (unsigned)i;
                    kmax = key > kmax ? key : kmax;
                }
                run += __popc(bal);
            }
            __syncwarp();  // the stage is refilled by the next iteration's issue
        }
        cp_async_wait<0>();
        if (lane == 0) p.wcount[c * PA_WARPS + warp] = run;
        grid_barrier(p.barrier, G, &s_bar_target);
        unsigned long long pre = 0, all = 0;  // members of the CTAs before this one, of all CTAs
        for (unsigned j = t; j < G * PA_WARPS; j += PA_THREADS) {
            const unsigned v = __ldcg(p.wcount + j);
            all += v;
            if (j < c * PA_WARPS) pre += v;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            pre += __shfl_xor_sync(0xffffffffu, pre, m);
            all += __shfl_xor_sync(0xffffffffu, all, m);
        }
        __syncthreads();
        if (lane == 0) s_red64[warp] = pre;
        __syncthreads();
        if (t == 0) {
            unsigned long long v = 0;
            for (int w = 0; w < PA_WARPS; ++w) v += s_red64[w];
            s_maxkey = v;  // borrowed: prefix
        }
        __syncthreads();
        unsigned long long dst = s_maxkey;
        __syncthreads();
        if (lane == 0) s_red64[warp] = all;
        __syncthreads();
        if (t == 0) {
            unsigned long long v = 0;
            for (int w = 0; w < PA_WARPS; ++w) v += s_red64[w];
            s_n = v;
        }
        {   // members of the warps of this CTA before this one
            const unsigned mine = __ldcg(p.wcount + c * PA_WARPS + lane);
            unsigned inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nn = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += nn;
            }
            dst += __shfl_sync(0xffffffffu, inc - mine, warp);
        }
        for (unsigned i0 = 0; i0 < run; i0 += 128) {  // four rows of 32 in flight
            unsigned long long kk[4];
            unsigned vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned i = i0 + 32u * (unsigned)u + lane;
                if (i < run) {
                    kk[u] = __ldcg(p.keys[1] + wb + i);
                    vv[u] = __ldcg(p.idx[1] + wb + i);
                }
            }
            if (p.rec_bytes) {
                // the member's row, gathered HERE (four independent rows per lane in flight, nothing waits on them) into
                // one record at its compacted position; the sort then carries positions, like a whole cloud's
#pragma unroll
                for (int h = 0; h < 4; h += 2) {  // two rows at a time (registers)
                    uint4 lo[2], hi[2];
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const int u = h + w;
                        const unsigned i = i0 + 32u * (unsigned)u + lane;
                        if (i < run) {
                            const unsigned long long row = vv[u];
                            lo[w].x = __float_as_uint(__ldg(p.pts + 3ull * row));
                            lo[w].y = __float_as_uint(__ldg(p.pts + 3ull * row + 1));
                            lo[w].z = __float_as_uint(__ldg(p.pts + 3ull * row + 2));
                            if (p.rec_bytes == 32) {
                                const unsigned* f = reinterpret_cast<const unsigned*>(p.feats) + 3ull * row;
                                lo[w].w = __ldg(f);
                                hi[w] = make_uint4(__ldg(f + 1), __ldg(f + 2), __ldg(reinterpret_cast<const unsigned*>(p.cls) + row), 0u);
                            } else {
                                const unsigned char* f = reinterpret_cast<const unsigned char*>(p.feats) + 3ull * row;
                                lo[w].w = (unsigned)__ldg(f) | ((unsigned)__ldg(f + 1) << 8) | ((unsigned)__ldg(f + 2) << 16) |
                                          ((unsigned)__ldg(reinterpret_cast<const unsigned char*>(p.cls) + row) << 24);
                            }
                        }
                    }
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const int u = h + w;
                        const unsigned i = i0 + 32u * (unsigned)u + lane;
                        if (i < run) {
                            uint4* r = reinterpret_cast<uint4*>(p.rec + (dst + i) * (unsigned long long)p.rec_bytes);
                            r[0] = lo[w];
                            if (p.rec_bytes == 32) r[1] = hi[w];
                            p.keys[0][dst + i] = kk[u];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned i = i0 + 32u * (unsigned)u + lane;
                    if (i < run) {
                        p.keys[0][dst + i] = kk[u];
                        p.idx[0][dst + i] = vv[u];
                    }
                }
            }
        }
        __syncthreads();
    } else {
        if (t == 0) s_n = p.N;
        for (unsigned i = t; i < BINS; i += PA_THREADS) s_tot[i] = 0;  // first-pass histogram of this CTA's chunk
        __syncthreads();
        if (cb < ce) tl.prefetch(cb, (unsigned)min((unsigned long long)PA_THREADS, ce - cb));
        for (unsigned long long b = cb; b < ce; b += PA_THREADS) {
            const unsigned cnt = (unsigned)min((unsigned long long)PA_THREADS, ce - b);
            const unsigned long long nb = b + PA_THREADS;
            float x = 0.f, y = 0.f, z = 0.f;
            tl.next(b, cnt, nb, nb < ce ? (unsigned)min((unsigned long long)PA_THREADS, ce - nb) : 0u, &x, &y, &z);
            const bool member = t < cnt;
            unsigned long long key = 0;
            if (member) {
                key = voxel_key(geo, x, y, z);
                p.keys[0][b + t] = key;
                kmax = key > kmax ? key : kmax;
            }
            // warp-aggregated histogram of the low digit
            const unsigned digit = member ? (unsigned)(key & (BINS - 1)) : BINS;
            const unsigned peers = match_digit(digit);
            if (member && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&s_tot[digit], __popc(peers));
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, kmax, m);
        kmax = o > kmax ? o : kmax;
    }
    __syncthreads();
    if (lane == 0) s_red64[warp] = kmax;
    __syncthreads();
    if (t == 0) {
        unsigned long long v = 0;
        for (int w = 0; w < PA_WARPS; ++w) v = s_red64[w] > v ? s_red64[w] : v;
        p.pmax[c] = v;
    }
    if (!slab)
        for (unsigned i = t; i < BINS; i += PA_THREADS) p.hist[(size_t)c * BINS + i] = s_tot[i];
    grid_barrier(p.barrier, G, &s_bar_target);
    {
        unsigned long long v = 0;
        for (unsigned b = t; b < G; b += PA_THREADS) {
            const unsigned long long o = __ldcg(p.pmax + b);
            v = o > v ? o : v;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
            v = o > v ? o : v;
        }
        __syncthreads();
        if (lane == 0) s_red64[warp] = v;
        __syncthreads();
        if (t == 0) {
            unsigned long long m = 0;
            for (int w = 0; w < PA_WARPS; ++w) m = s_red64[w] > m ? s_red64[w] : m;
            s_maxkey = m;
        }
        __syncthreads();
    }
    const unsigned long long n = s_n;
    const int key_bits = dev_bits_for_value(s_maxkey);
    if (t == 0) s_mark[2] = gtimer_ns();
    if (c == 0 && t == 0) {
        Meta m;
        for (int d = 0; d < 3; ++d) {
            m.mn[d] = geo.mn[d];
            m.mx[d] = geo.mx[d];
            m.origin[d] = geo.origin[d];
        }
        m.dl = geo.dl;
        m.nX = geo.nX;
        m.nY = geo.nY;
        m.nZ = geo.nZ;
        m.key_bits = key_bits;
        m.error = 0;
        m.max_key = s_maxkey;
        m.n_sel = n;
        m.M = 0;
        m.cur = 0;
        m.n_heavy = 0;
        m.n_deferred = 0;
        for (int k = 0; k < 16; ++k) m.tmark[k] = s_mark[k];
        *p.meta = m;
    }
    if (n == 0) return;  // an empty slab (uniform decision: every CTA leaves)

    // contiguous chunk of the n participating elements for this CTA
    const unsigned long long per = (((n + G - 1) / G) + PA_THREADS - 1) / PA_THREADS * PA_THREADS;
    const unsigned long long sb = min((unsigned long long)c * per, n), se = min(sb + per, n);

    // ---- P2: LSD radix passes
    int cur = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        const unsigned long long* kin = p.keys[cur];
        const unsigned* vin = p.idx[cur];
        unsigned long long* kout = p.keys[cur ^ 1];
        unsigned* vout = p.idx[cur ^ 1];
        const bool implicit_idx = shift == 0 && (!slab || p.rec_bytes != 0);  // first pass of a whole cloud or of packed slab members: value = position
        const bool chunk_matches = shift == 0 && !slab; // the P1 histogram was taken over [cb, ce), which is [sb, se)
        // How many DISTINCT digits does a row of 32 keys hold in this pass?  match.any costs one round per distinct
        // value (cheap on the top digits of a key, where a few values dominate), the ballot form a fixed nine: every
        // warp samples one row of its chunk and picks for the whole pass (speed only -- the masks are the same).
        bool few;
        {
            const unsigned long long i = sb + (unsigned long long)warp * 32u + lane;
            const unsigned dg = i < se ? (unsigned)(__ldcg(kin + i) >> shift) & (BINS - 1) : BINS;
            const unsigned pe = __match_any_sync(0xffffffffu, dg);
            few = __popc(__ballot_sync(0xffffffffu, (pe & ((1u << lane) - 1u)) == 0)) <= 12;
        }
        if (shift > 0 || !chunk_matches) {
            // (slab: P1 counted the members of this CTA's POINT chunk, the sort chunks partition the compacted array)
            for (unsigned i = t; i < BINS; i += PA_THREADS) s_tot[i] = 0;
            __syncthreads();
            for (unsigned long long b = sb; b < se; b += 4ull * PA_THREADS) {  // four rounds of keys in flight
                unsigned long long kk[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long i = b + (unsigned long long)u * PA_THREADS + t;
                    kk[u] = i < se ? __ldcg(kin + i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool in = b + (unsigned long long)u * PA_THREADS + t < se;
                    const unsigned digit = in ? (unsigned)(kk[u] >> shift) & (BINS - 1) : BINS;
                    const unsigned peers = few ? __match_any_sync(0xffffffffu, digit) : match_digit(digit);
                    if (in && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&s_tot[digit], __popc(peers));
                }
            }
            __syncthreads();
            for (unsigned i = t; i < BINS; i += PA_THREADS) p.hist[(size_t)c * BINS + i] = s_tot[i];
            grid_barrier(p.barrier, G, &s_bar_target);
        }
        // digit offsets of this CTA: (sum over all digits below) + (same digit, CTAs before this one)
        {
            const unsigned d = t & (BINS - 1), part = t >> 8;  // 4 x 256 threads: each quarter sums a quarter of the CTAs
            const unsigned q0 = part * ((G + 3) / 4), q1 = min(q0 + (G + 3) / 4, G);
            unsigned pre = 0, all = 0;
            for (unsigned b = q0; b < q1; ++b) {
                const unsigned v = __ldcg(p.hist + (size_t)b * BINS + d);
                all += v;
                if (b < c) pre += v;
            }
            __syncthreads();
            s_cnt[part][d] = pre;
            s_cnt[4 + part][d] = all;
            __syncthreads();
            if (t < BINS) {
                s_pre[t] = s_cnt[0][t] + s_cnt[1][t] + s_cnt[2][t] + s_cnt[3][t];
                s_tot[t] = s_cnt[4][t] + s_cnt[5][t] + s_cnt[6][t] + s_cnt[7][t];
            }
            __syncthreads();
            unsigned tot;
            const unsigned ex = block_excl_scan(t < BINS ? s_tot[t] : 0u, s_w, &tot);
            if (t < BINS) s_base[t] = ex + s_pre[t];
            __syncthreads();
        }
        // stable scatter of the chunk, PA_THREADS consecutive elements per round
        // stable scatter of the chunk: rounds of SC_ITEMS * PA_THREADS consecutive elements, warp w owning the
        // 32 * SC_ITEMS consecutive elements [w * 32 * SC_ITEMS, ...) of the round as SC_ITEMS rows of 32 (so both the
        // loads and the element order inside a warp are coalesced / sequential).  Ranking: match_any inside a row, a
        // per-warp running count per digit across the rows, one prefix over the warps per round.
        for (unsigned long long b = sb; b < se; b += (unsigned long long)SC_ITEMS * PA_THREADS) {
            for (unsigned i = t; i < PA_WARPS * BINS; i += PA_THREADS) (&s_cnt[0][0])[i] = 0;
            unsigned long long key[SC_ITEMS];
            unsigned val[SC_ITEMS], rank[SC_ITEMS];
#pragma unroll
            for (int k = 0; k < SC_ITEMS; ++k) {
                const unsigned long long i = b + (unsigned long long)warp * (32 * SC_ITEMS) + (unsigned)k * 32u + lane;
                key[k] = 0;
                val[k] = 0;
                if (i < se) {
                    key[k] = __ldcg(kin + i);
                    val[k] = implicit_idx ? (unsigned)i : __ldcg(vin + i);
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SC_ITEMS; ++k) {
                const unsigned long long i = b + (unsigned long long)warp * (32 * SC_ITEMS) + (unsigned)k * 32u + lane;
                const bool in = i < se;
                const unsigned digit = in ? (unsigned)(key[k] >> shift) & (BINS - 1) : BINS;
                const unsigned peers = few ? __match_any_sync(0xffffffffu, digit) : match_digit(digit);
                const unsigned r = __popc(peers & ((1u << lane) - 1u));
                const unsigned before = in ? s_cnt[warp][digit] : 0u;  // same digit in the earlier rows of this warp
                rank[k] = before + r;
                __syncwarp();
                if (in && r == 0) s_cnt[warp][digit] = before + __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            // exclusive prefix of every digit over the warps, then of the round's digit totals over the digits: the
            // round is first put in digit order in shared memory and written out from there, so that a warp's stores
            // fall into a few contiguous runs instead of 32 different places
            unsigned total = 0;
            if (t < BINS) {
#pragma unroll 8
                for (int w = 0; w < PA_WARPS; ++w) {
                    const unsigned cnt = s_cnt[w][t];
                    s_cnt[w][t] = total;
                    total += cnt;
                }
            }
            unsigned round_n;
            const unsigned ex = block_excl_scan(t < BINS ? total : 0u, s_w, &round_n);
            if (t < BINS) {
                s_pre[t] = ex;               // where the digit starts inside the round
                s_tot[t] = s_base[t] - ex;   // global position = s_tot[digit] + position inside the round
                s_base[t] += total;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SC_ITEMS; ++k) {
                const unsigned long long i = b + (unsigned long long)warp * (32 * SC_ITEMS) + (unsigned)k * 32u + lane;
                if (i < se) {
                    const unsigned d = (unsigned)(key[k] >> shift) & (BINS - 1);
                    const unsigned local = s_pre[d] + s_cnt[warp][d] + rank[k];
                    s_rkey[local] = key[k];
                    s_rval[local] = val[k];
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SC_ITEMS; ++k) {
                const unsigned j = (unsigned)k * PA_THREADS + t;
                if (j < round_n) {
                    const unsigned long long kk = s_rkey[j];
                    const unsigned pos = s_tot[(unsigned)(kk >> shift) & (BINS - 1)] + j;
                    kout[pos] = kk;
                    vout[pos] = s_rval[j];
                }
            }
            __syncthreads();
        }
        cur ^= 1;
        grid_barrier(p.barrier, G, &s_bar_target);
        if (t == 0 && shift / 8 < 8) s_mark[3 + shift / 8] = gtimer_ns();
    }

    // ---- P3: segment heads -> voxel starts.  Every warp owns a contiguous run of the CTA's chunk: it counts its heads,
    // the counts of all warps of the grid are prefixed after one barrier, and the warp writes its heads' positions
    // with ballot ranks -- no block barrier per tile.
    const unsigned long long* ks = p.keys[cur];
    const unsigned long long hlen = per / PA_WARPS;  // a multiple of 32
    const unsigned long long hb = min(sb + (unsigned long long)warp * hlen, se), he = min(hb + hlen, se);
    auto walk_heads = [&](bool write, unsigned long long first_id) {
        unsigned long long seen = 0;
        for (unsigned long long b0 = hb; b0 < he; b0 += 128) {  // four rows of keys in flight
            unsigned long long kk[4], prev0[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long i = b0 + 32u * (unsigned)u + lane;
                kk[u] = i < he ? __ldcg(ks + i) : 0ull;
                prev0[u] = (lane == 0 && i < he && i > 0) ? __ldcg(ks + i - 1) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long i = b0 + 32u * (unsigned)u + lane;
                unsigned long long pk = __shfl_up_sync(0xffffffffu, kk[u], 1);
                if (lane == 0) pk = prev0[u];
                const bool head = i < he && (i == 0 || kk[u] != pk);
                const unsigned bal = __ballot_sync(0xffffffffu, head);
                if (write && head) p.starts[first_id + seen + __popc(bal & ((1u << lane) - 1u))] = (unsigned)i;
                seen += __popc(bal);
            }
        }
        return seen;
    };
    const unsigned long long my_heads = walk_heads(false, 0ull);
    if (lane == 0) p.wcount[c * PA_WARPS + warp] = (unsigned)my_heads;
    grid_barrier(p.barrier, G, &s_bar_target);
    unsigned long long vbase, M;
    {
        unsigned long long pre = 0, all = 0;  // heads of the CTAs before this one, of all CTAs
        for (unsigned j = t; j < G * PA_WARPS; j += PA_THREADS) {
            const unsigned v = __ldcg(p.wcount + j);
            all += v;
            if (j < c * PA_WARPS) pre += v;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            pre += __shfl_xor_sync(0xffffffffu, pre, m);
            all += __shfl_xor_sync(0xffffffffu, all, m);
        }
        __syncthreads();
        if (lane == 0) s_red64[warp] = pre;
        __syncthreads();
        if (t == 0) {
            unsigned long long v = 0;
            for (int w = 0; w < PA_WARPS; ++w) v += s_red64[w];
            s_maxkey = v;  // borrowed: voxel id of this CTA's first head
        }
        __syncthreads();
        vbase = s_maxkey;
        __syncthreads();
        if (lane == 0) s_red64[warp] = all;
        __syncthreads();
        if (t == 0) {
            unsigned long long v = 0;
            for (int w = 0; w < PA_WARPS; ++w) v += s_red64[w];
            s_n = v;  // borrowed: M
        }
        __syncthreads();
        M = s_n;
        const unsigned mine = __ldcg(p.wcount + c * PA_WARPS + lane);  // heads of the earlier warps of this CTA
        unsigned inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned nn = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += nn;
        }
        vbase += __shfl_sync(0xffffffffu, inc - mine, warp);
    }
    if (t == 0) s_mark[11] = gtimer_ns();
    walk_heads(true, vbase);
    __syncthreads();
    if (c == 0 && t == 0) {
        p.starts[M] = (unsigned)n;
        p.meta->M = M;
        p.meta->cur = cur;
        s_mark[12] = gtimer_ns();
        for (int k = 3; k < 13; ++k) p.meta->tmark[k] = s_mark[k];
    }
}

// ---- label vote: libstdc++ unordered_map<int,int> iteration order (identity hash, unique keys) ----------
// Faithful emulation of _M_insert_unique_node / _M_rehash_aux for up to LABEL_CAP nodes.
__device__ int label_first_in_iteration_order(const int* labels, const int* counts, int n, int maxcount) {
    const int NB[5] = {1, 13, 29, 59, 127};
    short next[LABEL_CAP];
    short bucket[127];  // node BEFORE the first node of the bucket; -2 = before_begin, -1 = empty
    int nb = 1, level = 0, head = -1;
    bucket[0] = -1;
    for (int id = 0; id < n; ++id) {
        if (id + 1 > (level == 0 ? 0 : nb)) {  // _Prime_rehash_policy::_M_need_rehash
            ++level;
            const int nnb = NB[level];
            for (int b = 0; b < nnb; ++b) bucket[b] = -1;
            int p = head, bbegin = 0;
            head = -1;
            while (p >= 0) {
                const int nx = next[p];
                const int b = (int)((unsigned long long)(long long)labels[p] % (unsigned long long)nnb);
                if (bucket[b] == -1) {
                    next[p] = (short)head;
                    head = p;
                    bucket[b] = -2;
                    if (next[p] >= 0) bucket[bbegin] = (short)p;
                    bbegin = b;
                } else {
                    const int before = bucket[b];
                    if (before == -2) {
                        next[p] = (short)head;
                        head = p;
                    } else {
                        next[p] = next[before];
                        next[before] = (short)p;
                    }
                }
                p = nx;
            }
            nb = nnb;
        }
        const int b = (int)((unsigned long long)(long long)labels[id] % (unsigned long long)nb);
        if (bucket[b] != -1) {
            const int before = bucket[b];
            if (before == -2) {
                next[id] = (short)head;
                head = id;
            } else {
                next[id] = next[before];
                next[before] = (short)id;
            }
        } else {
            next[id] = (short)head;
            head = id;
            if (next[id] >= 0)
                bucket[(int)((unsigned long long)(long long)labels[next[id]] % (unsigned long long)nb)] = (short)id;
            bucket[b] = -2;
        }
    }
    for (int p = head; p >= 0; p = next[p])
        if (counts[p] == maxcount) return labels[p];
    return labels[0];
}

// ---- per-voxel sequential reduce: eight lanes per voxel --------------------------------------------------------------
constexpr int RB_THREADS = 256;
constexpr int RB_GROUPS = RB_THREADS / 8;
constexpr int STAGE_STRIDE = 72;  // 8 points x (8 channels + 1 pad): stores (stride 9) and loads of the four groups of a warp are conflict free
constexpr int MAX_CB = 4;         // channel blocks of 8 held in registers per walk over a voxel (32 channels)
constexpr unsigned HEAVY_MIN = 1024;  // voxels with more points than this are reduced by a CTA each
constexpr int RH_THREADS = 256;
constexpr int RH_GATHER = RH_THREADS - 32;     // warps 1..7 gather, the first eight lanes of warp 0 add
constexpr int RH_PER = 2;                      // points per gatherer thread and batch
constexpr int RH_BATCH = RH_GATHER * RH_PER;   // 448 points in flight per batch

struct ReduceParams {
    const float* pts;
    const void* feats;  // float32 or uint8 rows
    const void* cls;    // int32 or uint8 rows
    int feat_u8, cls_u8;
    int fdim, ldim;
    const unsigned long long* keys[2];
    const unsigned* idx[2];
    const unsigned* starts;
    Meta* meta;
    float* out_p;
    float* out_f;
    int* out_c;
    unsigned long long* out_k;
    int* out_n;
    unsigned* heavy;        // voxels with more than heavy_min points, in no particular order (meta->n_heavy of them)
    unsigned heavy_cap;
    unsigned heavy_min;     // voxels with more points than this are deferred to reduce_heavy_kernel
    int direct;             // pts / feats / cls are rows in SORTED order (sorted_rows_kernel ran): row = sorted position
    unsigned pts_stride, feat_stride, cls_stride;  // bytes from one row to the next
    int only_deferred;      // reduce_small_kernel ran: take only the voxels of its list (meta->n_deferred of them)
    unsigned* deferred;     // that list, in no particular order
    unsigned small_max;     // voxels of at most this many points are tried by reduce_small_kernel
};

// Input rows packed into one record each -- xyz | features | labels, padded to a multiple of 16 bytes -- so that the
// reduce's gather through the sorted index costs ONE or two 32-byte sectors per point instead of one or two per ARRAY
// (12-byte rows straddle sector boundaries: 3.75 sectors per point for xyz + rgb + label).  Sequential reads and
// writes, run before the sort when the cloud is larger than the L2 (on an L2-resident cloud the gathers are cheap).
struct PackParams {
    const float* pts;
    const void* feats;
    const void* cls;
    unsigned feat_bytes, cls_bytes;  // per row
    unsigned cls_off, rec_bytes;     // labels start here inside a record; record size
    unsigned long long N;
    unsigned char* rec;
};
constexpr int PK_THREADS = 256;
__global__ void __launch_bounds__(PK_THREADS) pack_rows_kernel(const PackParams p) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const bool fast = p.rec_bytes == 32 && p.feat_bytes == 12 && p.cls_bytes == 4 && p.cls_off == 24;
    const bool fast16 = p.rec_bytes == 16 && p.feat_bytes == 3 && p.cls_bytes == 1 && p.cls_off == 15;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += stride) {
        const float x = __ldg(p.pts + 3ull * i), y = __ldg(p.pts + 3ull * i + 1), z = __ldg(p.pts + 3ull * i + 2);
        unsigned char* r = p.rec + i * p.rec_bytes;
        if (fast) {  // xyz + three float32 features + one int32 label: two 16-byte stores
            const unsigned* f = reinterpret_cast<const unsigned*>(p.feats) + 3ull * i;
            const unsigned f0 = __ldg(f), f1 = __ldg(f + 1), f2 = __ldg(f + 2);
            const unsigned lab = __ldg(reinterpret_cast<const unsigned*>(p.cls) + i);
            reinterpret_cast<uint4*>(r)[0] = make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(z), f0);
            reinterpret_cast<uint4*>(r)[1] = make_uint4(f1, f2, lab, 0u);
            continue;
        }
        if (fast16) {  // xyz + three uint8 features + one uint8 label: one 16-byte store
            const unsigned char* f = reinterpret_cast<const unsigned char*>(p.feats) + 3ull * i;
            const unsigned w = (unsigned)__ldg(f) | ((unsigned)__ldg(f + 1) << 8) | ((unsigned)__ldg(f + 2) << 16) |
                               ((unsigned)__ldg(reinterpret_cast<const unsigned char*>(p.cls) + i) << 24);
            *reinterpret_cast<uint4*>(r) = make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(z), w);
            continue;
        }
        reinterpret_cast<float*>(r)[0] = x;
        reinterpret_cast<float*>(r)[1] = y;
        reinterpret_cast<float*>(r)[2] = z;
        if (p.feat_bytes) {
            const unsigned char* a = reinterpret_cast<const unsigned char*>(p.feats) + i * p.feat_bytes;
            if ((p.feat_bytes & 3u) == 0)
                for (unsigned j = 0; j < p.feat_bytes; j += 4) *reinterpret_cast<unsigned*>(r + 12 + j) = __ldg(reinterpret_cast<const unsigned*>(a + j));
            else
                for (unsigned j = 0; j < p.feat_bytes; ++j) r[12 + j] = __ldg(a + j);
        }
        if (p.cls_bytes) {
            const unsigned char* a = reinterpret_cast<const unsigned char*>(p.cls) + i * p.cls_bytes;
            if ((p.cls_bytes & 3u) == 0)
                for (unsigned j = 0; j < p.cls_bytes; j += 4) *reinterpret_cast<unsigned*>(r + p.cls_off + j) = __ldg(reinterpret_cast<const unsigned*>(a + j));
            else
                for (unsigned j = 0; j < p.cls_bytes; ++j) r[p.cls_off + j] = __ldg(a + j);
        }
    }
}

// Rows of the input arrays in sorted (voxel-major, input order inside a voxel) order, so that the reduce reads
// consecutive records instead of chasing an index per point.  The gathers happen HERE, where every thread has
// SR_ITEMS independent rows in flight and nothing waits on them; the features and labels keep their input type.
constexpr int SR_THREADS = 256;
constexpr int SR_ITEMS = 4;
struct SortedRowsParams {
    const float* pts;
    const void* feats;
    const void* cls;
    int feat_bytes, cls_bytes;  // bytes per ROW (0: absent)
    const unsigned* idx[2];
    const Meta* meta;
    float* out_p;
    unsigned char* out_f;
    unsigned char* out_c;
};
template <typename W>
__device__ __forceinline__ void copy_row(const void* src, void* dst, unsigned long long from, unsigned long long to, int words) {
    const W* a = reinterpret_cast<const W*>(src) + from * (unsigned long long)words;
    W* b = reinterpret_cast<W*>(dst) + to * (unsigned long long)words;
    for (int j = 0; j < words; ++j) b[j] = __ldg(a + j);
}
__global__ void __launch_bounds__(SR_THREADS) sorted_rows_kernel(const SortedRowsParams p) {
    const unsigned long long n = p.meta->n_sel;
    const unsigned* idx = p.meta->cur ? p.idx[1] : p.idx[0];
    const unsigned long long tile = (unsigned long long)SR_THREADS * SR_ITEMS;
    for (unsigned long long b = (unsigned long long)blockIdx.x * tile; b < n; b += (unsigned long long)gridDim.x * tile) {
        unsigned long long row[SR_ITEMS];
#pragma unroll
        for (int u = 0; u < SR_ITEMS; ++u) {
            const unsigned long long i = b + (unsigned)u * SR_THREADS + threadIdx.x;
            row[u] = i < n ? (unsigned long long)__ldg(idx + i) : 0ull;
        }
        float x[SR_ITEMS], y[SR_ITEMS], z[SR_ITEMS];
#pragma unroll
        for (int u = 0; u < SR_ITEMS; ++u) {
            x[u] = __ldg(p.pts + 3ull * row[u]);
            y[u] = __ldg(p.pts + 3ull * row[u] + 1);
            z[u] = __ldg(p.pts + 3ull * row[u] + 2);
        }
#pragma unroll
        for (int u = 0; u < SR_ITEMS; ++u) {
            const unsigned long long i = b + (unsigned)u * SR_THREADS + threadIdx.x;
            if (i < n) {
                p.out_p[3ull * i] = x[u];
                p.out_p[3ull * i + 1] = y[u];
                p.out_p[3ull * i + 2] = z[u];
            }
        }
        if (p.feat_bytes) {
#pragma unroll
            for (int u = 0; u < SR_ITEMS; ++u) {
                const unsigned long long i = b + (unsigned)u * SR_THREADS + threadIdx.x;
                if (i < n) {
                    if ((p.feat_bytes & 3) == 0) copy_row<unsigned>(p.feats, p.out_f, row[u], i, p.feat_bytes >> 2);
                    else copy_row<unsigned char>(p.feats, p.out_f, row[u], i, p.feat_bytes);
                }
            }
        }
        if (p.cls_bytes) {
#pragma unroll
            for (int u = 0; u < SR_ITEMS; ++u) {
                const unsigned long long i = b + (unsigned)u * SR_THREADS + threadIdx.x;
                if (i < n) {
                    if ((p.cls_bytes & 3) == 0) copy_row<unsigned>(p.cls, p.out_c, row[u], i, p.cls_bytes >> 2);
                    else copy_row<unsigned char>(p.cls, p.out_c, row[u], i, p.cls_bytes);
                }
            }
        }
    }
}

// Rows are addressed as base + row * stride (bytes): the caller's own arrays (strides 12, fdim * size, ldim * size) or
// the packed records of pack_rows_kernel (one stride for all three, every row inside one or two 32-byte sectors).
__device__ __forceinline__ float load_coord(const ReduceParams& p, unsigned long long row, int ch) {
    return __ldg(reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p.pts) + row * p.pts_stride) + ch);
}
__device__ __forceinline__ float load_feat(const ReduceParams& p, unsigned long long row, int j) {
    const unsigned char* r = reinterpret_cast<const unsigned char*>(p.feats) + row * p.feat_stride;
    return p.feat_u8 ? (float)__ldg(r + j) : __ldg(reinterpret_cast<const float*>(r) + j);
}
__device__ __forceinline__ int load_label(const ReduceParams& p, unsigned long long row, int col) {
    const unsigned char* r = reinterpret_cast<const unsigned char*>(p.cls) + row * p.cls_stride;
    return p.cls_u8 ? (int)__ldg(r + col) : __ldg(reinterpret_cast<const int*>(r) + col);
}

// The label table of one voxel (first-occurrence order, like the reference's unordered_map insertions); every lane of
// the group calls this with the same arguments.  Returns false on overflow.
__device__ __forceinline__ bool table_add(int* labs, int* cnts, int* nl, int label, int weight, unsigned gmask, int gshift,
                                          int l) {
    int found = -1;
    for (int q0 = 0; q0 < *nl && found < 0; q0 += 8) {
        const int q = q0 + l;
        const bool hit = q < *nl && labs[q] == label;
        const unsigned b = (__ballot_sync(gmask, hit) >> gshift) & 0xFFu;
        if (b) found = q0 + __ffs((int)b) - 1;
    }
    bool ok = true;
    if (found >= 0) {
        if (l == 0) cnts[found] += weight;
    } else if (*nl < LABEL_CAP) {
        if (l == 0) {
            labs[*nl] = label;
            cnts[*nl] = weight;
        }
        *nl += 1;
    } else {
        ok = false;
    }
    __syncwarp(gmask);
    return ok;
}

__global__ void __launch_bounds__(RB_THREADS, 5) reduce_kernel(const ReduceParams p) {  // 5 CTAs per SM: the gathers are latency bound (48 -> 80 registers cost 20 %)
    __shared__ float s_stage[RB_GROUPS][STAGE_STRIDE];
    __shared__ int s_labs[RB_GROUPS][LABEL_CAP], s_cnts[RB_GROUPS][LABEL_CAP];
    const unsigned long long M = p.meta->M;
    const int cur = p.meta->cur;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.meta->tmark[13] = gtimer_ns();
    const unsigned long long* keys = p.keys[cur];
    const unsigned* idx = p.idx[cur];
    const int tid = threadIdx.x, g = tid >> 3, l = tid & 7, lane = tid & 31;
    const int gshift = (lane >> 3) * 8;
    const unsigned gmask = 0xFFu << gshift;
    float* stage = s_stage[g];
    int* labs = s_labs[g];
    int* cnts = s_cnts[g];
    const int CH = 3 + p.fdim;
    bool overflow = false;
    const unsigned long long ngroups = (unsigned long long)gridDim.x * RB_GROUPS;
    for (unsigned long long v = (unsigned long long)blockIdx.x * RB_GROUPS + g; v < M; v += ngroups) {
        const unsigned long long s = p.starts[v], e = p.starts[v + 1];
        const unsigned cnt = (unsigned)(e - s);
        if (cnt > p.heavy_min) {  // a whole CTA takes it (reduce_heavy_kernel): eight lanes would walk it for milliseconds
            if (l == 0) {
                const unsigned pos = atomicAdd(&p.meta->n_heavy, 1u);
                if (pos < p.heavy_cap) p.heavy[pos] = (unsigned)v;
            }
            continue;
        }
        const float a = (float)(1.0 / (double)cnt);  // grid_subsampling.cpp:87: double reciprocal narrowed to float
        const float cf = (float)cnt;
        // walks over the voxel: 32 channels per walk (one walk for every reference caller: xyz + rgb)
        for (int ch0 = 0; ch0 < CH; ch0 += 8 * MAX_CB) {
            float acc[MAX_CB];
#pragma unroll
            for (int cb = 0; cb < MAX_CB; ++cb) acc[cb] = 0.f;
            int nl = 0;
            const bool vote = ch0 == 0 && p.ldim >= 1;
            // the sorted index of the NEXT eight points is fetched while this step's rows are gathered and added
            unsigned long long row_next = (unsigned)l < cnt ? (p.direct ? s + l : (unsigned long long)idx[s + l]) : 0ull;
            for (unsigned c0 = 0; c0 < cnt; c0 += 8) {
                const int m = (int)min(8u, cnt - c0);
                const unsigned long long row = row_next;
                if (c0 + 8 + l < cnt) row_next = p.direct ? s + c0 + 8 + l : (unsigned long long)idx[s + c0 + 8 + l];
#pragma unroll
                for (int cb = 0; cb < MAX_CB; ++cb) {
                    const int cbase = ch0 + 8 * cb;
                    if (cbase < CH) {
                        if (l < m) {  // this lane stages the (up to) eight channels of ITS point
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int ch = cbase + j;
                                if (ch < CH) stage[l * 9 + j] = ch < 3 ? load_coord(p, row, ch) : load_feat(p, row, ch - 3);
                            }
                        }
                        __syncwarp(gmask);
                        if (cbase + l < CH) {  // this lane owns channel cbase + l: add the points in input order
#pragma unroll
                            for (int t = 0; t < 8; ++t)
                                if (t < m) acc[cb] = __fadd_rn(acc[cb], stage[t * 9 + l]);
                        }
                        __syncwarp(gmask);
                    }
                }
                if (vote) {
                    const int lab = l < m ? load_label(p, row, 0) : 0;
                    const int first = __shfl_sync(gmask, lab, gshift);
                    const bool same = l >= m || lab == first;
                    if (((__ballot_sync(gmask, same) >> gshift) & 0xFFu) == 0xFFu) {
                        overflow |= !table_add(labs, cnts, &nl, first, m, gmask, gshift, l);
                    } else {
                        for (int t = 0; t < m; ++t)
                            overflow |= !table_add(labs, cnts, &nl, __shfl_sync(gmask, lab, gshift + t), 1, gmask, gshift, l);
                    }
                }
            }
#pragma unroll
            for (int cb = 0; cb < MAX_CB; ++cb) {
                const int ch = ch0 + 8 * cb + l;
                if (ch < CH) {
                    if (ch < 3) p.out_p[3ull * v + ch] = __fmul_rn(acc[cb], a);
                    else p.out_f[v * (unsigned long long)p.fdim + (ch - 3)] = __fdiv_rn(acc[cb], cf);
                }
            }
            if (vote) {
                for (int col = 0;;) {
                    // winner of the table: largest count, ties by the reference's hash-iteration order
                    int best = -1;
                    for (int q = l; q < nl; q += 8) best = max(best, cnts[q]);
#pragma unroll
                    for (int mm = 1; mm < 8; mm <<= 1) best = max(best, __shfl_xor_sync(gmask, best, mm));
                    int nbest = 0, arg = -1;
                    for (int q0 = 0; q0 < nl; q0 += 8) {
                        const int q = q0 + l;
                        const unsigned b = (__ballot_sync(gmask, q < nl && cnts[q] == best) >> gshift) & 0xFFu;
                        if (b && arg < 0) arg = q0 + __ffs((int)b) - 1;
                        nbest += __popc(b);
                    }
                    if (l == 0)
                        p.out_c[v * (unsigned long long)p.ldim + col] =
                            nbest == 1 ? labs[arg] : label_first_in_iteration_order(labs, cnts, nl, best);
                    __syncwarp(gmask);
                    if (++col >= p.ldim) break;
                    // further label columns: one more walk each (rare: ldim is 1 for every reference caller)
                    nl = 0;
                    for (unsigned c0 = 0; c0 < cnt; c0 += 8) {
                        const int m = (int)min(8u, cnt - c0);
                        const int lab = l < m ? load_label(p, p.direct ? s + c0 + l : (unsigned long long)idx[s + c0 + l], col) : 0;
                        for (int t = 0; t < m; ++t)
                            overflow |= !table_add(labs, cnts, &nl, __shfl_sync(gmask, lab, gshift + t), 1, gmask, gshift, l);
                    }
                }
            }
        }
        if (l == 0) {
            p.out_k[v] = keys[s];
            p.out_n[v] = (int)cnt;
        }
    }
    if (overflow) p.meta->error = 2;
    if (threadIdx.x == 0) atomicMax(&p.meta->tmark[14], gtimer_ns());
}

__device__ __forceinline__ uint4 ldg_gather16(const void* ptr) { return __ldg(reinterpret_cast<const uint4*>(ptr)); }
// (an ld.global.nc.L2::64B hint on these gathers changed nothing: 3.09 vs 3.06 ms for the 80 M-point scan)

// The same reduce for the two layouts every reference caller has -- xyz + three colour channels + one label column --
// read from PACKED records (pack_rows_kernel): REC = 32: float32 colours, int32 label (device callers); REC = 16: uint8
// colours and label (the data-preparation scripts).  A point is ONE or two 16-byte loads, everything else is compile-time
// constant: about a third of the generic kernel's instructions per voxel.
template <int REC>
__global__ void __launch_bounds__(RB_THREADS, 5) reduce_rec_kernel(const ReduceParams p) {
    __shared__ __align__(16) float s_stage[RB_GROUPS][STAGE_STRIDE];  // 8 points x 8 words (+ 8 pad: the four groups of a warp read different banks)
    __shared__ int s_labs[RB_GROUPS][LABEL_CAP], s_cnts[RB_GROUPS][LABEL_CAP];
    const unsigned long long M = p.meta->M;
    const int cur = p.meta->cur;
    if (blockIdx.x == 0 && threadIdx.x == 0 && !p.only_deferred) p.meta->tmark[13] = gtimer_ns();
    const unsigned long long* keys = p.keys[cur];
    const unsigned* idx = p.idx[cur];
    const unsigned char* rec = reinterpret_cast<const unsigned char*>(p.pts);
    const int tid = threadIdx.x, g = tid >> 3, l = tid & 7, lane = tid & 31;
    const int gshift = (lane >> 3) * 8;
    const unsigned gmask = 0xFFu << gshift;
    float* stage = s_stage[g];
    int* labs = s_labs[g];
    int* cnts = s_cnts[g];
    bool overflow = false;
    const unsigned long long ngroups = (unsigned long long)gridDim.x * RB_GROUPS;
    const unsigned long long n_work = p.only_deferred ? (unsigned long long)p.meta->n_deferred : M;
    for (unsigned long long w = (unsigned long long)blockIdx.x * RB_GROUPS + g; w < n_work; w += ngroups) {
        const unsigned long long v = p.only_deferred ? (unsigned long long)__ldcg(p.deferred + w) : w;
        const unsigned long long s = p.starts[v], e = p.starts[v + 1];
        const unsigned cnt = (unsigned)(e - s);
        if (cnt > p.heavy_min) {
            if (l == 0) {
                const unsigned pos = atomicAdd(&p.meta->n_heavy, 1u);
                if (pos < p.heavy_cap) p.heavy[pos] = (unsigned)v;
            }
            continue;
        }
        const float a = (float)(1.0 / (double)cnt);  // grid_subsampling.cpp:87
        const float cf = (float)cnt;
        float acc = 0.f;
        int nl = 0;
        unsigned long long row_next = (unsigned)l < cnt ? (unsigned long long)idx[s + l] : 0ull;
        for (unsigned c0 = 0; c0 < cnt; c0 += 8) {
            const int m = (int)min(8u, cnt - c0);
            const unsigned long long row = row_next;
            if (c0 + 8 + l < cnt) row_next = (unsigned long long)idx[s + c0 + 8 + l];
            int lab = 0;
            if (l < m) {
                const uint4 r0 = ldg_gather16(rec + row * REC);
                float4 lo, hi;
                lo.x = __uint_as_float(r0.x);
                lo.y = __uint_as_float(r0.y);
                lo.z = __uint_as_float(r0.z);
                if (REC == 32) {
                    const uint4 r1 = ldg_gather16(rec + row * REC + 16);
                    lo.w = __uint_as_float(r0.w);
                    hi.x = __uint_as_float(r1.x);
                    hi.y = __uint_as_float(r1.y);
                    lab = (int)r1.z;
                } else {
                    lo.w = (float)(r0.w & 0xFFu);
                    hi.x = (float)((r0.w >> 8) & 0xFFu);
                    hi.y = (float)((r0.w >> 16) & 0xFFu);
                    lab = (int)(r0.w >> 24);
                }
                hi.z = hi.w = 0.f;
                reinterpret_cast<float4*>(stage + l * 8)[0] = lo;
                reinterpret_cast<float4*>(stage + l * 8)[1] = hi;
            }
            __syncwarp(gmask);
            if (l < 6) {  // this lane owns channel l: add the points in input order
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t < m) acc = __fadd_rn(acc, stage[t * 8 + l]);
            }
            __syncwarp(gmask);
            const int first = __shfl_sync(gmask, lab, gshift);
            const bool same = l >= m || lab == first;
            if (((__ballot_sync(gmask, same) >> gshift) & 0xFFu) == 0xFFu) {
                overflow |= !table_add(labs, cnts, &nl, first, m, gmask, gshift, l);
            } else {
                for (int t = 0; t < m; ++t)
                    overflow |= !table_add(labs, cnts, &nl, __shfl_sync(gmask, lab, gshift + t), 1, gmask, gshift, l);
            }
        }
        if (l < 3) p.out_p[3ull * v + l] = __fmul_rn(acc, a);
        else if (l < 6) p.out_f[3ull * v + (l - 3)] = __fdiv_rn(acc, cf);
        {   // winner of the table: largest count, ties by the reference's hash-iteration order
            int best = -1;
            for (int q = l; q < nl; q += 8) best = max(best, cnts[q]);
#pragma unroll
            for (int mm = 1; mm < 8; mm <<= 1) best = max(best, __shfl_xor_sync(gmask, best, mm));
            int nbest = 0, arg = -1;
            for (int q0 = 0; q0 < nl; q0 += 8) {
                const int q = q0 + l;
                const unsigned b = (__ballot_sync(gmask, q < nl && cnts[q] == best) >> gshift) & 0xFFu;
                if (b && arg < 0) arg = q0 + __ffs((int)b) - 1;
                nbest += __popc(b);
            }
            if (l == 0) p.out_c[v] = nbest == 1 ? labs[arg] : label_first_in_iteration_order(labs, cnts, nl, best);
            __syncwarp(gmask);
        }
        if (l == 0) {
            p.out_k[v] = keys[s];
            p.out_n[v] = (int)cnt;
        }
    }
    if (overflow) p.meta->error = 2;
    if (threadIdx.x == 0) atomicMax(&p.meta->tmark[14], gtimer_ns());
}

// Most voxels are small and hold one label (6.5 points per voxel in a 0.06 m scan, 9 in a 0.04 m room): ONE THREAD takes
// such a voxel -- its records are fetched four at a time (eight independent 16-byte loads in flight), the six sums
// stay in registers and are added in input order, no shared memory, no shuffles, all 32 lanes of a warp busy with 32
// voxels.  A voxel with more than SMALL_MAX points or a second label is appended to a list (one atomic per warp) and
// left to the eight-lane groups of reduce_rec_kernel, which then walk that list instead of all voxels.
constexpr unsigned SMALL_MAX = 24;
constexpr int RS_THREADS = 128;
template <int REC>
__global__ void __launch_bounds__(RS_THREADS) reduce_small_kernel(const ReduceParams p) {
    const unsigned long long M = p.meta->M;
    const int cur = p.meta->cur;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.meta->tmark[13] = gtimer_ns();
    const unsigned long long* keys = p.keys[cur];
    const unsigned* idx = p.idx[cur];
    const unsigned char* rec = reinterpret_cast<const unsigned char*>(p.pts);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    // (every lane of a warp makes the same number of trips, so the warp can append its deferred voxels together)
    for (unsigned long long vb = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x - lane); vb < M; vb += stride) {
        const unsigned long long v = vb + lane;
        const bool valid = v < M;
        const unsigned s = valid ? p.starts[v] : 0u, e = valid ? p.starts[v + 1] : 0u;
        const unsigned cnt = e - s;
        const bool small = valid && cnt <= p.small_max;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
        int lab0 = 0;
        bool same = true;
        for (unsigned t0 = 0; small && t0 < cnt; t0 += 4) {
            uint4 lo[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (t0 + j < cnt) {
                    const unsigned char* r = rec + (unsigned long long)__ldg(idx + s + t0 + j) * REC;
                    lo[j] = ldg_gather16(r);
                    if (REC == 32) hi[j] = ldg_gather16(r + 16);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (t0 + j < cnt) {
                    float f0, f1, f2;
                    int lab;
                    if (REC == 32) {
                        f0 = __uint_as_float(lo[j].w);
                        f1 = __uint_as_float(hi[j].x);
                        f2 = __uint_as_float(hi[j].y);
                        lab = (int)hi[j].z;
                    } else {
                        f0 = (float)(lo[j].w & 0xFFu);
                        f1 = (float)((lo[j].w >> 8) & 0xFFu);
                        f2 = (float)((lo[j].w >> 16) & 0xFFu);
                        lab = (int)(lo[j].w >> 24);
                    }
                    a0 = __fadd_rn(a0, __uint_as_float(lo[j].x));
                    a1 = __fadd_rn(a1, __uint_as_float(lo[j].y));
                    a2 = __fadd_rn(a2, __uint_as_float(lo[j].z));
                    a3 = __fadd_rn(a3, f0);
                    a4 = __fadd_rn(a4, f1);
                    a5 = __fadd_rn(a5, f2);
                    if (t0 + j == 0) lab0 = lab;
                    same = same && lab == lab0;
                }
            }
        }
        // larger voxels, and voxels that need a vote, go to the groups: one list append per warp
        const bool defer = valid && (!small || !same);
        const unsigned dbal = __ballot_sync(0xffffffffu, defer);
        if (dbal) {
            unsigned base = 0;
            if (lane == (unsigned)(__ffs((int)dbal) - 1)) base = atomicAdd(&p.meta->n_deferred, (unsigned)__popc(dbal));
            base = __shfl_sync(0xffffffffu, base, __ffs((int)dbal) - 1);
            if (defer) p.deferred[base + __popc(dbal & ((1u << lane) - 1u))] = (unsigned)v;
        }
        if (!valid || defer) continue;
        const float a = (float)(1.0 / (double)cnt);  // grid_subsampling.cpp:87
        const float cf = (float)cnt;
        p.out_p[3ull * v] = __fmul_rn(a0, a);
        p.out_p[3ull * v + 1] = __fmul_rn(a1, a);
        p.out_p[3ull * v + 2] = __fmul_rn(a2, a);
        p.out_f[3ull * v] = __fdiv_rn(a3, cf);
        p.out_f[3ull * v + 1] = __fdiv_rn(a4, cf);
        p.out_f[3ull * v + 2] = __fdiv_rn(a5, cf);
        p.out_c[v] = lab0;
        p.out_k[v] = keys[s];
        p.out_n[v] = (int)cnt;
    }
}

// Heavy voxels (a terrestrial scan puts tens of thousands of points into the voxels next to the scanner): the sums
// must still be the reference's sequential += in input order, so the ADDS of a channel stay one dependent chain -- but
// the gathers need not wait for them.  One CTA per voxel: warps 1..7 gather batch b+1 (sorted index -> xyz, features,
// label; 448 points in flight) into one half of a double buffer while the first eight lanes of warp 0 -- one lane per
// channel, exactly the group of reduce_kernel -- add batch b out of the other half and count its labels.  On one GPU
// the reduce is bound by its gather throughput and this path changes little (80 M-point scan: 11.0 vs 11.4 ms); it
// matters when the cloud is sharded: the slab next to the scanner holds few, huge voxels and its eight-lane chains
// (17 600 points: 4.4 ms) were the critical path of the whole sharded call.
__global__ void __launch_bounds__(RH_THREADS) reduce_heavy_kernel(const ReduceParams p) {
    __shared__ float s_val[2][RH_BATCH * 9];
    __shared__ int s_lab[2][RH_BATCH];
    __shared__ int s_labs[LABEL_CAP], s_cnts[LABEL_CAP];
    __shared__ int s_mixed[2];  // batch in buffer b holds a label other than the voxel's first one
    const unsigned nheavy = min(p.meta->n_heavy, p.heavy_cap);
    const int cur = p.meta->cur;
    const unsigned long long* keys = p.keys[cur];
    const unsigned* idx = p.idx[cur];
    const int tid = threadIdx.x, l = tid & 7;
    const unsigned gmask = 0xFFu;
    const int CH = 3 + p.fdim;
    bool overflow = false;
    for (unsigned h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const unsigned long long v = p.heavy[h];
        const unsigned long long s = p.starts[v], e = p.starts[v + 1];
        const unsigned cnt = (unsigned)(e - s);
        const unsigned nb = (cnt + RH_BATCH - 1) / RH_BATCH;
        const float a = (float)(1.0 / (double)cnt);
        const float cf = (float)cnt;
        for (int ch0 = 0; ch0 < CH; ch0 += 8) {
            const bool vote = ch0 == 0 && p.ldim >= 1;
            float acc = 0.f;
            int nl = 0;
            // labels are piecewise constant in space: a batch whose labels all equal the voxel's first label is counted
            // with ONE table update instead of one per eight points (the adder lanes are the bottleneck otherwise)
            const int lab0 = vote ? load_label(p, p.direct ? s : (unsigned long long)idx[s], 0) : 0;
            auto gather = [&](unsigned b, int buf) {
#pragma unroll
                for (int k = 0; k < RH_PER; ++k) {
                    const unsigned il = (unsigned)k * RH_GATHER + (unsigned)(tid - 32);
                    const unsigned long long i = (unsigned long long)b * RH_BATCH + il;
                    if (i < cnt) {
                        const unsigned long long row = p.direct ? s + i : (unsigned long long)idx[s + i];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int ch = ch0 + j;
                            if (ch < CH) s_val[buf][il * 9 + j] = ch < 3 ? load_coord(p, row, ch) : load_feat(p, row, ch - 3);
                        }
                        if (vote) {
                            const int lab = load_label(p, row, 0);
                            s_lab[buf][il] = lab;
                            if (lab != lab0) s_mixed[buf] = 1;
                        }
                    }
                }
            };
            __syncthreads();  // the previous walk's / voxel's buffers are free
            if (tid < 2) s_mixed[tid] = 0;
            __syncthreads();
            if (tid >= 32) gather(0, 0);
            __syncthreads();
            for (unsigned b = 0; b < nb; ++b) {
                const int buf = (int)(b & 1u);
                if (tid >= 32) {
                    if (b + 1 < nb) gather(b + 1, buf ^ 1);
                } else if (tid < 8) {
                    const unsigned mt = min((unsigned)RH_BATCH, cnt - b * RH_BATCH);
                    const bool mixed = vote && s_mixed[buf] != 0;
                    if (vote && !mixed) overflow |= !table_add(s_labs, s_cnts, &nl, lab0, (int)mt, gmask, 0, l);
                    for (unsigned c0 = 0; c0 < mt; c0 += 8) {
                        const int m = (int)min(8u, mt - c0);
                        if (ch0 + l < CH) {
#pragma unroll
                            for (int t = 0; t < 8; ++t)
                                if (t < m) acc = __fadd_rn(acc, s_val[buf][(c0 + t) * 9 + l]);
                        }
                        if (mixed) {
                            const int lab = l < m ? s_lab[buf][c0 + l] : 0;
                            const int first = __shfl_sync(gmask, lab, 0);
                            const bool same = l >= m || lab == first;
                            if ((__ballot_sync(gmask, same) & 0xFFu) == 0xFFu) {
                                overflow |= !table_add(s_labs, s_cnts, &nl, first, m, gmask, 0, l);
                            } else {
                                for (int t = 0; t < m; ++t)
                                    overflow |= !table_add(s_labs, s_cnts, &nl, __shfl_sync(gmask, lab, t), 1, gmask, 0, l);
                            }
                        }
                    }
                    if (l == 0) s_mixed[buf] = 0;  // the buffer is refilled in the next iteration, behind the barrier
                }
                __syncthreads();
            }
            if (tid < 8) {
                const int ch = ch0 + l;
                if (ch < CH) {
                    if (ch < 3) p.out_p[3ull * v + ch] = __fmul_rn(acc, a);
                    else p.out_f[v * (unsigned long long)p.fdim + (ch - 3)] = __fdiv_rn(acc, cf);
                }
                if (vote) {
                    for (int col = 0;;) {
                        int best = -1;
                        for (int q = l; q < nl; q += 8) best = max(best, s_cnts[q]);
#pragma unroll
                        for (int mm = 1; mm < 8; mm <<= 1) best = max(best, __shfl_xor_sync(gmask, best, mm));
                        int nbest = 0, arg = -1;
                        for (int q0 = 0; q0 < nl; q0 += 8) {
                            const int q = q0 + l;
                            const unsigned bb = __ballot_sync(gmask, q < nl && s_cnts[q] == best) & 0xFFu;
                            if (bb && arg < 0) arg = q0 + __ffs((int)bb) - 1;
                            nbest += __popc(bb);
                        }
                        if (l == 0)
                            p.out_c[v * (unsigned long long)p.ldim + col] =
                                nbest == 1 ? s_labs[arg] : label_first_in_iteration_order(s_labs, s_cnts, nl, best);
                        __syncwarp(gmask);
                        if (++col >= p.ldim) break;
                        nl = 0;  // further label columns (no reference caller has any): straight from global memory
                        for (unsigned c0 = 0; c0 < cnt; c0 += 8) {
                            const int m = (int)min(8u, cnt - c0);
                            const int lab = l < m ? load_label(p, p.direct ? s + c0 + l : (unsigned long long)idx[s + c0 + l], col) : 0;
                            for (int t = 0; t < m; ++t)
                                overflow |= !table_add(s_labs, s_cnts, &nl, __shfl_sync(gmask, lab, t), 1, gmask, 0, l);
                        }
                    }
                }
                if (l == 0 && ch0 == 0) {
                    p.out_k[v] = keys[s];
                    p.out_n[v] = (int)cnt;
                }
            }
        }
    }
    if (overflow) p.meta->error = 2;
    if (threadIdx.x == 0) atomicMax(&p.meta->tmark[14], gtimer_ns());
}

// ---- 7. reference row order: libstdc++ unordered_map<size_t,...> iteration order, epoch by epoch ----------------
// The map is rehashed through a fixed prime sequence; within one bucket count ("epoch") the list is the sequence of
// bucket runs in REVERSE order of bucket creation, each run in REVERSE insertion order, where the epoch's insertion
// sequence is: the previous list (a rehash re-inserts it in list order), then the new nodes (SURVEY.md A.3;
// tools/proto_hash_order.py checks this closed form against the sequential emulation).  Each epoch is therefore an
// atomicMin (bucket creation time) plus one stable radix sort.
__global__ void first_index_kernel(const unsigned* __restrict__ idx_sorted, const unsigned* __restrict__ starts,
                                   unsigned long long M, unsigned long long* __restrict__ fkey,
                                   unsigned* __restrict__ vox) {
    const unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    fkey[v] = idx_sorted[starts[v]];  // a voxel is inserted when its first point (smallest input index) arrives
    vox[v] = (unsigned)v;
}
__global__ void node_key_kernel(const unsigned long long* __restrict__ vkeys, const unsigned* __restrict__ ins,
                                unsigned long long M, unsigned long long* __restrict__ nkey) {
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < M) nkey[r] = vkeys[ins[r]];
}
__global__ void epoch_bucket_kernel(const unsigned long long* __restrict__ nkey, const unsigned* __restrict__ cur,
                                    unsigned prev, unsigned hi, unsigned long long nb, unsigned* __restrict__ node,
                                    unsigned* __restrict__ bkt, unsigned* __restrict__ cmin) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hi) return;
    const unsigned n = j < prev ? cur[j] : j;
    const unsigned b = (unsigned)(nkey[n] % nb);
    node[j] = n;
    bkt[j] = b;
    atomicMin(&cmin[b], j);
}
__global__ void epoch_key_kernel(const unsigned* __restrict__ bkt, const unsigned* __restrict__ cmin, unsigned hi,
                                 int sh, unsigned long long* __restrict__ comp) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < hi) comp[j] = ((unsigned long long)(hi - 1 - cmin[bkt[j]]) << sh) | (unsigned long long)(hi - 1 - j);
}
__global__ void compose_perm_kernel(const unsigned* __restrict__ cur, const unsigned* __restrict__ ins,
                                    unsigned long long M, unsigned* __restrict__ perm) {
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < M) perm[r] = ins[cur[r]];
}
template <typename T>
__global__ void permute_rows_kernel(const T* __restrict__ in, T* __restrict__ out, const unsigned* __restrict__ perm,
                                    unsigned long long M, int width) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * (unsigned long long)width) return;
    const unsigned long long r = i / width;
    out[i] = in[(unsigned long long)perm[r] * width + (i - r * width)];
}

static const unsigned long long kBuckets[] = {13ull,        29ull,        59ull,        127ull,       257ull,
                                              541ull,       1109ull,      2357ull,      5087ull,      10273ull,
                                              20753ull,     42043ull,     85229ull,     172933ull,    351061ull,
                                              712697ull,    1447153ull,   2938679ull,   5967347ull,   12117689ull,
                                              24607243ull,  49969847ull,  101473717ull, 206062531ull, 418451333ull,
                                              849749479ull, 1725587117ull, 3504151727ull};

struct Handle {
    size_t M = 0, fdim = 0, ldim = 0;
    float* d_p = nullptr;
    float* d_f = nullptr;
    int* d_c = nullptr;
    unsigned long long* d_k = nullptr;
    int* d_n = nullptr;
    cudaStream_t stream = nullptr;
    int device = 0;
};

static void free_handle(Handle* h) {  // stream-ordered pool: no device-wide synchronisation, memory is recycled
    if (!h) return;
    if (h->d_p) cudaFreeAsync(h->d_p, h->stream);
    if (h->d_f) cudaFreeAsync(h->d_f, h->stream);
    if (h->d_c) cudaFreeAsync(h->d_c, h->stream);
    if (h->d_k) cudaFreeAsync(h->d_k, h->stream);
    if (h->d_n) cudaFreeAsync(h->d_n, h->stream);
    delete h;
}

// Rows are in ascending-key order in the handle; permute them into the reference's hash-iteration order.
// Reuses the sort workspaces (stream ordered after reduce_kernel): WS_KEYS/WS_KEYS2 (u64), WS_IDX/WS_IDX2 (u32).
static int reorder_reference(Ctx* c, cudaStream_t s, Handle* h, const unsigned* idx_sorted, const unsigned* starts,
                             size_t N, unsigned* scratch) {
    const size_t M = h->M;
    const unsigned mb = (unsigned)((M + 255) / 256);
    // idx_sorted lives in WS_IDX or WS_IDX2, starts in WS_STARTS: copy what we need before those buffers are recycled
    SSDR_TRY(c->ws[WS_HASH].reserve(M * (8 + 4 + 4 + 4 + 4)));
    unsigned long long* nkey = c->ws[WS_HASH].as<unsigned long long>();  // [M] voxel key by insertion rank
    unsigned* ins = reinterpret_cast<unsigned*>(nkey + M);               // [M] insertion rank -> voxel (key order)
    unsigned* cur = ins + M;                                             // [M] list order (node = insertion rank)
    unsigned* node = cur + M;                                            // [M] epoch sequence -> node
    unsigned* bkt = node + M;                                            // [M]
    unsigned long long* ka = c->ws[WS_KEYS].as<unsigned long long>();
    unsigned long long* kb = c->ws[WS_KEYS2].as<unsigned long long>();
    unsigned* va = c->ws[WS_IDX].as<unsigned>();
    unsigned* vb = c->ws[WS_IDX2].as<unsigned>();
    // 1. insertion order of the voxels = ascending first input index.  first_index_kernel reads idx_sorted/starts and
    //    writes into the OTHER ping-pong pair, so nothing it still needs is overwritten.
    const bool sorted_in_b = (idx_sorted == vb);
    unsigned long long* fk = sorted_in_b ? ka : kb;
    unsigned* fv = sorted_in_b ? va : vb;
    first_index_kernel<<<mb, 256, 0, s>>>(idx_sorted, starts, M, fk, fv);
    // sort (first index, voxel) -- after this the old sorted arrays are dead and both pairs are scratch
    int curbuf = 0;
    unsigned long long* k0 = fk;
    unsigned* v0 = fv;
    unsigned long long* k1 = sorted_in_b ? kb : ka;
    unsigned* v1 = sorted_in_b ? vb : va;
    SSDR_TRY(prim::radix_sort_pairs(k0, v0, k1, v1, M, bits_for_value(N), scratch, &curbuf, s));
    SSDR_CHECK_CUDA(cudaMemcpyAsync(ins, curbuf ? v1 : v0, M * sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
    node_key_kernel<<<mb, 256, 0, s>>>(h->d_k, ins, M, nkey);
    // 2. epochs
    size_t prev = 0;
    for (size_t e = 0; e < sizeof(kBuckets) / sizeof(kBuckets[0]) && prev < M; ++e) {
        const unsigned long long nb = kBuckets[e];
        const size_t hi = nb < M ? (size_t)nb : M;
        SSDR_TRY(c->ws[WS_CMIN].reserve((size_t)nb * sizeof(unsigned)));
        unsigned* cmin = c->ws[WS_CMIN].as<unsigned>();
        SSDR_CHECK_CUDA(cudaMemsetAsync(cmin, 0xFF, (size_t)nb * sizeof(unsigned), s));
        const unsigned eb = (unsigned)((hi + 255) / 256);
        epoch_bucket_kernel<<<eb, 256, 0, s>>>(nkey, cur, (unsigned)prev, (unsigned)hi, nb, node, bkt, cmin);
        const int sh = bits_for_value(hi);
        epoch_key_kernel<<<eb, 256, 0, s>>>(bkt, cmin, (unsigned)hi, sh, ka);
        SSDR_CHECK_CUDA(cudaMemcpyAsync(va, node, hi * sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
        SSDR_TRY(prim::radix_sort_pairs(ka, va, kb, vb, hi, 2 * sh, scratch, &curbuf, s));
        SSDR_CHECK_CUDA(cudaMemcpyAsync(cur, curbuf ? vb : va, hi * sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
        prev = hi;
    }
    SSDR_REQUIRE(prev >= M, SSDR_ERR_UNSUPPORTED, "too many voxels for the reference-order emulation");
    // 3. permute the rows
    unsigned* perm = node;
    compose_perm_kernel<<<mb, 256, 0, s>>>(cur, ins, M, perm);
    auto permute = [&](auto** buf, int width) -> int {
        typedef typename std::remove_pointer<typename std::remove_pointer<decltype(buf)>::type>::type T;
        if (!*buf) return SSDR_OK;
        T* nbuf = nullptr;
        cudaError_t e2 = cudaMallocAsync((void**)&nbuf, M * width * sizeof(T), s);
        if (e2 != cudaSuccess) {
            cudaGetLastError();
            return set_error(SSDR_ERR_NOMEM, "cudaMallocAsync failed: %s", cudaGetErrorString(e2));
        }
        const unsigned long long tot = (unsigned long long)M * width;
        permute_rows_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(*buf, nbuf, perm, M, width);
        cudaFreeAsync(*buf, s);
        *buf = nbuf;
        return SSDR_OK;
    };
    SSDR_TRY(permute(&h->d_p, 3));
    SSDR_TRY(permute(&h->d_f, (int)h->fdim));
    SSDR_TRY(permute(&h->d_c, (int)h->ldim));
    SSDR_TRY(permute(&h->d_k, 1));
    SSDR_TRY(permute(&h->d_n, 1));
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// min/max of the points (or the caller's bounding box of a larger cloud this chunk belongs to) -> Meta
static int geometry(Ctx* c, cudaStream_t s, const float* d_p, size_t N, float dl, const float* bbox, Meta* meta) {
    const int nparts = c->sm_count * 4;
    SSDR_TRY(c->ws[WS_PART].reserve((size_t)nparts * 6 * sizeof(float)));
    float* part = c->ws[WS_PART].as<float>();
    if (bbox) {
        for (int d = 0; d < 3; ++d)
            SSDR_REQUIRE(bbox[d] <= bbox[3 + d], SSDR_ERR_INVALID, "bbox min exceeds max (or NaN) on axis %d", d);
        bbox_fill_kernel<<<1, 1, 0, s>>>(part, bbox[0], bbox[1], bbox[2], bbox[3], bbox[4], bbox[5]);
        setup_kernel<<<1, 192, 0, s>>>(part, 1, dl, meta);
    } else {
        minmax_kernel<<<nparts, MM_BLOCK, 0, s>>>(d_p, N, part);
        setup_kernel<<<1, 192, 0, s>>>(part, nparts, dl, meta);
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

static thread_local Meta g_last_meta;  // of the calling thread's last run (ssdr_grid_debug_timing)

struct Slab {
    int axis = -1;  // -1: whole cloud
    unsigned long long lo = 0, hi = 0;
};

struct Inputs {
    const float* p = nullptr;
    const void* f = nullptr;  // float32 rows, or uint8 rows when f_u8
    const void* c = nullptr;  // int32 rows, or uint8 rows when c_u8
    bool f_u8 = false, c_u8 = false;
};

struct Outputs {  // device arrays with room for `capacity` rows (N in the worst case); nullable where noted
    float* p = nullptr;
    float* f = nullptr;               // needed when fdim > 0
    int* c = nullptr;                 // needed when ldim > 0
    unsigned long long* k = nullptr;  // voxel keys
    int* n = nullptr;                 // points per voxel
};

// Everything up to the single host round trip: sort_kernel, the reduce, the read-back of Meta (voxel count, error flag).
static int run_core(Ctx* c, cudaStream_t s, const Inputs& in, size_t N, size_t fdim, size_t ldim, float dl, Slab slab,
                    const float* bbox, const Outputs& out, Meta* hm_out, SortParams* sp_out) {
    typedef unsigned long long KeyT;
    const int G = c->sm_count;
    SSDR_TRY(c->ws[WS_META].reserve(sizeof(Meta)));
    SSDR_TRY(c->ws[WS_KEYS].reserve(N * sizeof(KeyT)));
    SSDR_TRY(c->ws[WS_KEYS2].reserve(N * sizeof(KeyT)));
    SSDR_TRY(c->ws[WS_IDX].reserve(N * sizeof(unsigned)));
    SSDR_TRY(c->ws[WS_IDX2].reserve(N * sizeof(unsigned)));
    SSDR_TRY(c->ws[WS_STARTS].reserve((N + 1) * sizeof(unsigned)));
    SSDR_TRY(c->ws[WS_HEADS].reserve((N / HEAVY_MIN + 2 + N) * sizeof(unsigned)));  // voxels left to reduce_heavy_kernel | to the groups
    // control block: barrier counter | per-CTA partials, largest keys, counts | histogram matrix
    const size_t ctl_bytes = 256 + (size_t)G * (6 * sizeof(float) + sizeof(KeyT) + sizeof(unsigned)) + 256 +
                             (size_t)G * BINS * sizeof(unsigned) + (size_t)G * PA_WARPS * sizeof(unsigned);
    SSDR_TRY(c->ws[WS_CTL].reserve(ctl_bytes));
    char* ctl = c->ws[WS_CTL].as<char>();
    Meta* meta = c->ws[WS_META].as<Meta>();

    SortParams sp;
    sp.pts = in.p;
    sp.N = N;
    sp.dl = dl;
    sp.has_bbox = bbox ? 1 : 0;
    for (int d = 0; d < 6; ++d) sp.bbox[d] = bbox ? bbox[d] : 0.f;
    sp.slab_axis = slab.axis;
    sp.slab_lo = slab.lo;
    sp.slab_hi = slab.hi;
    sp.meta = meta;
    sp.keys[0] = c->ws[WS_KEYS].as<KeyT>();
    sp.keys[1] = c->ws[WS_KEYS2].as<KeyT>();
    sp.idx[0] = c->ws[WS_IDX].as<unsigned>();
    sp.idx[1] = c->ws[WS_IDX2].as<unsigned>();
    sp.barrier = reinterpret_cast<unsigned*>(ctl);
    sp.pmax = reinterpret_cast<KeyT*>(ctl + 256);
    sp.partials = reinterpret_cast<float*>(ctl + 256 + (size_t)G * sizeof(KeyT));
    sp.cta_count = reinterpret_cast<unsigned*>(ctl + 256 + (size_t)G * (sizeof(KeyT) + 6 * sizeof(float)));
    sp.hist = reinterpret_cast<unsigned*>(ctl + (256 + (size_t)G * (6 * sizeof(float) + sizeof(KeyT) + sizeof(unsigned)) + 255) / 256 * 256);
    sp.wcount = sp.hist + (size_t)G * BINS;
    // slab + one of the two reference layouts: the members are packed into records while they are compacted
    // (SSDR_GRID_PACK=0 keeps the gather through the input index, for A/B runs)
    const char* pack_e = getenv("SSDR_GRID_PACK");  // read per call: the tests flip it inside one process
    const int pack_env = pack_e && pack_e[0] ? (pack_e[0] == '0' ? 0 : 1) : -1;
    const bool layout32 = fdim == 3 && ldim == 1 && !in.f_u8 && !in.c_u8, layout16 = fdim == 3 && ldim == 1 && in.f_u8 && in.c_u8;
    const bool slab_pack = slab.axis >= 0 && layout32 && pack_env != 0;  // (device callers pass float32 / int32 rows)
    sp.feats = in.f;
    sp.cls = in.c;
    sp.rec = nullptr;
    sp.rec_bytes = 0;
    if (slab_pack) {
        sp.rec_bytes = layout32 ? 32 : 16;
        SSDR_TRY(c->ws[WS_REC].reserve(N * (size_t)sp.rec_bytes));
        sp.rec = c->ws[WS_REC].as<unsigned char>();
    }
    sp.starts = c->ws[WS_STARTS].as<unsigned>();
    SSDR_CHECK_CUDA(cudaMemsetAsync(sp.barrier, 0, 256, s));
    {
        void* args[] = {(void*)&sp};
        size_t dyn = (size_t)SC_ITEMS * PA_THREADS * (sizeof(unsigned long long) + sizeof(unsigned));
        const size_t ring = (size_t)PA_WARPS * SLAB_STAGES * 384 * sizeof(float);  // the slab pass's cp.async rings
        if (slab.axis >= 0 && ring > dyn) dyn = ring;
        SSDR_CHECK_CUDA(cudaFuncSetAttribute((void*)sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        SSDR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)sort_kernel, dim3(G), dim3(PA_THREADS), args, dyn, s));
    }
    {
        ReduceParams rp;
        rp.pts = in.p;
        rp.feats = fdim ? in.f : nullptr;
        rp.cls = ldim ? in.c : nullptr;
        rp.feat_u8 = in.f_u8 ? 1 : 0;
        rp.cls_u8 = in.c_u8 ? 1 : 0;
        rp.fdim = (int)fdim;
        rp.ldim = (int)ldim;
        rp.keys[0] = sp.keys[0];
        rp.keys[1] = sp.keys[1];
        rp.idx[0] = sp.idx[0];
        rp.idx[1] = sp.idx[1];
        rp.starts = sp.starts;
        rp.meta = meta;
        rp.out_p = out.p;
        rp.out_f = out.f;
        rp.out_c = out.c;
        rp.out_k = out.k;
        rp.out_n = out.n;
        size_t want = (N + RB_GROUPS - 1) / RB_GROUPS;  // M <= N voxels, one group each at most
        const size_t cap = (size_t)c->sm_count * 8;
        const unsigned blocks = (unsigned)(want < cap ? (want ? want : 1) : cap);
        rp.heavy = c->ws[WS_HEADS].as<unsigned>();
        rp.heavy_cap = (unsigned)(N / HEAVY_MIN + 1);
        rp.deferred = rp.heavy + (N / HEAVY_MIN + 2);
        static const bool heavy_on = [] {
            const char* e = getenv("SSDR_GRID_HEAVY");  // SSDR_GRID_HEAVY=0: every voxel by its eight-lane group (A/B runs)
            return !(e && e[0] == '0');
        }();
        rp.heavy_min = heavy_on ? HEAVY_MIN : 0xFFFFFFFFu;
        const size_t fbytes = fdim ? fdim * (in.f_u8 ? 1 : 4) : 0, cbytes = ldim ? ldim * (in.c_u8 ? 1 : 4) : 0;
        rp.pts_stride = 12;
        rp.feat_stride = (unsigned)fbytes;
        rp.cls_stride = (unsigned)cbytes;
        // SSDR_GRID_PACK=0 / 1 forces the packed records off / on (A/B runs); default: whole clouds beyond the L2
        const bool pack = slab.axis < 0 && (fbytes + cbytes) > 0 &&
                          (pack_env >= 0 ? pack_env == 1
                                         : (N * (12 + fbytes + cbytes) >= ((size_t)128 << 20) ||  // beyond the L2 ...
                                            (fdim == 3 && ldim == 1 && in.f_u8 == in.c_u8)));     // ... or a layout with a reduce of its own
        if (pack) {
            PackParams q;
            q.pts = in.p;
            q.feats = rp.feats;
            q.cls = rp.cls;
            q.feat_bytes = (unsigned)fbytes;
            q.cls_bytes = (unsigned)cbytes;
            q.cls_off = (unsigned)((cbytes && !in.c_u8) ? (12 + fbytes + 3) / 4 * 4 : 12 + fbytes);
            q.rec_bytes = (unsigned)((q.cls_off + cbytes + 15) / 16 * 16);
            q.N = N;
            SSDR_TRY(c->ws[WS_REC].reserve(N * (size_t)q.rec_bytes));
            q.rec = c->ws[WS_REC].as<unsigned char>();
            const size_t want_b = (N + PK_THREADS - 1) / PK_THREADS, cap_b = (size_t)c->sm_count * 32;
            pack_rows_kernel<<<(unsigned)(want_b < cap_b ? want_b : cap_b), PK_THREADS, 0, s>>>(q);
            rp.pts = reinterpret_cast<const float*>(q.rec);
            if (fbytes) rp.feats = q.rec + 12;
            if (cbytes) rp.cls = q.rec + q.cls_off;
            rp.pts_stride = rp.feat_stride = rp.cls_stride = q.rec_bytes;
        }
        static const bool sorted_rows_on = [] {
            // =1: rows gathered into sorted order by a kernel of their own, sequential reduce (A/B runs: the gather pass
            // costs more than the reduce saves, profiles/r02_grid_sorted_rows_ab.txt)
            const char* e = getenv("SSDR_GRID_SORTED_ROWS");
            return e && e[0] == '1';
        }();
        rp.direct = 0;
        if (sorted_rows_on && !pack) {
            const size_t fb = fbytes, cb = cbytes;
            const size_t o_f = (N * 12 + 255) / 256 * 256, o_c = o_f + (N * fb + 255) / 256 * 256;
            SSDR_TRY(c->ws[WS_REC].reserve(o_c + N * cb + 256));
            unsigned char* rec = c->ws[WS_REC].as<unsigned char>();
            SortedRowsParams q;
            q.pts = in.p;
            q.feats = rp.feats;
            q.cls = rp.cls;
            q.feat_bytes = (int)fb;
            q.cls_bytes = (int)cb;
            q.idx[0] = sp.idx[0];
            q.idx[1] = sp.idx[1];
            q.meta = meta;
            q.out_p = reinterpret_cast<float*>(rec);
            q.out_f = rec + o_f;
            q.out_c = rec + o_c;
            const size_t tiles = (N + (size_t)SR_THREADS * SR_ITEMS - 1) / ((size_t)SR_THREADS * SR_ITEMS);
            const size_t capb = (size_t)c->sm_count * 16;
            sorted_rows_kernel<<<(unsigned)(tiles < capb ? tiles : capb), SR_THREADS, 0, s>>>(q);
            rp.pts = q.out_p;
            if (fb) rp.feats = q.out_f;
            if (cb) rp.cls = q.out_c;
            rp.direct = 1;
        }
        // the two layouts of the reference callers (xyz + rgb + one label column) have a reduce of their own
        const bool rec32 = (pack || slab_pack) && layout32;
        const bool rec16 = (pack || slab_pack) && layout16;
        if (slab_pack) {  // (the heavy-voxel reduce reads the same records through the strides)
            rp.pts = reinterpret_cast<const float*>(sp.rec);
            rp.feats = sp.rec + 12;
            rp.cls = sp.rec + (layout32 ? 24 : 15);
            rp.pts_stride = rp.feat_stride = rp.cls_stride = (unsigned)sp.rec_bytes;
        }
        rp.only_deferred = 0;
        if (rec32 || rec16) {
            const char* se = getenv("SSDR_GRID_SMALL");  // =0: every voxel by its eight-lane group (A/B runs, tests)
            if (!(se && se[0] == '0')) {
                rp.only_deferred = 1;
                const char* sm = getenv("SSDR_GRID_SMALL_MAX");
                rp.small_max = sm && atoi(sm) > 0 ? (unsigned)atoi(sm) : SMALL_MAX;
                const size_t want_s = (N + RS_THREADS - 1) / RS_THREADS, cap_s = (size_t)c->sm_count * 16;
                const unsigned bs = (unsigned)(want_s < cap_s ? (want_s ? want_s : 1) : cap_s);
                if (rec32) reduce_small_kernel<32><<<bs, RS_THREADS, 0, s>>>(rp);
                else reduce_small_kernel<16><<<bs, RS_THREADS, 0, s>>>(rp);
            }
        }
        if (rec32) reduce_rec_kernel<32><<<blocks, RB_THREADS, 0, s>>>(rp);
        else if (rec16) reduce_rec_kernel<16><<<blocks, RB_THREADS, 0, s>>>(rp);
        else reduce_kernel<<<blocks, RB_THREADS, 0, s>>>(rp);
        // the voxels the groups passed over (count on the device; none: the CTAs return at once)
        if (heavy_on) reduce_heavy_kernel<<<(unsigned)c->sm_count * 4, RH_THREADS, 0, s>>>(rp);
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    // the ONE host round trip of the call: voxel count (the caller sizes its arrays with it) and the error flag
    SSDR_TRY(d2h_sync(c, hm_out, meta, sizeof(Meta), s));
    g_last_meta = *hm_out;
    SSDR_REQUIRE(hm_out->error != 2, SSDR_ERR_UNSUPPORTED, "more than %d distinct labels inside one voxel", LABEL_CAP);
    if (sp_out) *sp_out = sp;
    return SSDR_OK;
}

static int check_args(const Inputs& in, size_t N, size_t& fdim, size_t& ldim, float dl, int order, const Slab& slab,
                      const float* bbox) {
    SSDR_REQUIRE(in.p, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_EMPTY, "Error");
    SSDR_REQUIRE(N < 0x7FFFFFFFull, SSDR_ERR_UNSUPPORTED, "N=%zu exceeds 2^31-2 points per call", N);
    SSDR_REQUIRE(order == SSDR_GRID_ORDER_KEY || order == SSDR_GRID_ORDER_REFERENCE, SSDR_ERR_INVALID,
                 "order must be SSDR_GRID_ORDER_KEY or SSDR_GRID_ORDER_REFERENCE");
    SSDR_REQUIRE(dl > 0.0f, SSDR_ERR_INVALID, "sampleDl must be positive");
    if (!in.f) fdim = 0;
    if (!in.c) ldim = 0;
    if (bbox)
        for (int d = 0; d < 3; ++d)
            SSDR_REQUIRE(bbox[d] <= bbox[3 + d], SSDR_ERR_INVALID, "bbox min exceeds max (or NaN) on axis %d", d);
    if (slab.axis >= 0) SSDR_REQUIRE(slab.axis < 3 && slab.lo <= slab.hi, SSDR_ERR_INVALID, "bad slab");
    return SSDR_OK;
}

static int run_dev(Ctx* c, cudaStream_t s, const Inputs& in, size_t N, size_t fdim, size_t ldim, float dl, int order,
                   size_t* M_out, void** handle, Slab slab = Slab(), const float* bbox = nullptr) {
    SSDR_REQUIRE(M_out && handle, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_TRY(check_args(in, N, fdim, ldim, dl, order, slab, bbox));
    // outputs are sized for the worst case (every point its own voxel): M is known only on the device
    Handle* h = new Handle();
    h->stream = s;
    h->device = c->device;
    h->fdim = fdim;
    h->ldim = ldim;
    auto fail = [&](int rc) {
        free_handle(h);
        return rc;
    };
#define SSDR_ALLOC(ptr, bytes)                                                                            \
    do {                                                                                                  \
        cudaError_t _e = cudaMallocAsync((void**)&(ptr), (bytes), s);                                             \
        if (_e != cudaSuccess) {                                                                          \
            cudaGetLastError();                                                                           \
            return fail(set_error(SSDR_ERR_NOMEM, "cudaMallocAsync(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(_e))); \
        }                                                                                                 \
    } while (0)
    SSDR_ALLOC(h->d_p, N * 3 * sizeof(float));
    if (fdim) SSDR_ALLOC(h->d_f, N * fdim * sizeof(float));
    if (ldim) SSDR_ALLOC(h->d_c, N * ldim * sizeof(int));
    SSDR_ALLOC(h->d_k, N * sizeof(unsigned long long));
    SSDR_ALLOC(h->d_n, N * sizeof(int));
#undef SSDR_ALLOC
    Outputs out;
    out.p = h->d_p;
    out.f = h->d_f;
    out.c = h->d_c;
    out.k = h->d_k;
    out.n = h->d_n;
    Meta hm;
    SortParams sp;
    {
        int rc = run_core(c, s, in, N, fdim, ldim, dl, slab, bbox, out, &hm, &sp);
        if (rc != SSDR_OK) return fail(rc);
    }
    h->M = (size_t)hm.M;
    if (slab.axis >= 0 && hm.n_sel == 0) {  // an empty slab is a valid shard: zero rows
        *M_out = 0;
        *handle = h;
        return SSDR_OK;
    }
    if (!(h->M >= 1 && h->M <= N)) return fail(set_error(SSDR_ERR_EMPTY, "Error"));
    if (order == SSDR_GRID_ORDER_REFERENCE && h->M > 1) {
        const size_t n_sel = (size_t)hm.n_sel;
        SSDR_TRY(c->ws[WS_TEMP].reserve((prim::rs_scratch_words(n_sel) + prim::scan_scratch_words(n_sel) + 8) * sizeof(unsigned)));
        unsigned* scratch = c->ws[WS_TEMP].as<unsigned>();
        SSDR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (prim::rs_scratch_words(n_sel) + prim::scan_scratch_words(n_sel) + 8) * sizeof(unsigned), s));
        int rc = reorder_reference(c, s, h, sp.idx[hm.cur], sp.starts, N, scratch);
        if (rc != SSDR_OK) return fail(rc);
    }
    *M_out = h->M;
    *handle = h;
    return SSDR_OK;
}

}  // namespace grid
}  // namespace ssdr

using namespace ssdr;

extern "C" {

int ssdr_grid_subsample_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N,
                            size_t fdim, size_t ldim, float sampleDl, int order, void* stream, size_t* M_out,
                            void** handle) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    grid::Inputs in;
    in.p = d_points;
    in.f = d_feats;
    in.c = d_classes;
    return grid::run_dev(c, (cudaStream_t)stream, in, N, fdim, ldim, sampleDl, order, M_out, handle);
}

int ssdr_grid_subsample_slab_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N,
                                 size_t fdim, size_t ldim, float sampleDl, int order, const float* bbox, int axis,
                                 unsigned long long layer_lo, unsigned long long layer_hi, void* stream,
                                 size_t* M_out, void** handle) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    SSDR_REQUIRE(axis >= -1 && axis <= 2, SSDR_ERR_INVALID, "axis must be -1 (no slab), 0, 1 or 2");
    SSDR_REQUIRE(order == SSDR_GRID_ORDER_KEY || (axis < 0 && !bbox), SSDR_ERR_UNSUPPORTED,
                 "the reference's hash-iteration order is defined for a whole cloud only; slabs come in key order");
    grid::Slab slab;
    slab.axis = axis;
    slab.lo = layer_lo;
    slab.hi = layer_hi;
    grid::Inputs in;
    in.p = d_points;
    in.f = d_feats;
    in.c = d_classes;
    return grid::run_dev(c, (cudaStream_t)stream, in, N, fdim, ldim, sampleDl, order, M_out, handle, slab, bbox);
}

int ssdr_grid_subsample_into_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N,
                                 size_t fdim, size_t ldim, float sampleDl, const float* bbox, int axis,
                                 unsigned long long layer_lo, unsigned long long layer_hi, float* d_points_out,
                                 float* d_feats_out, int32_t* d_classes_out, uint64_t* d_keys_out, int32_t* d_counts_out,
                                 size_t capacity, void* stream, size_t* M_out) {
    SSDR_REQUIRE(d_points_out && M_out, SSDR_ERR_INVALID, "NULL pointer");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    SSDR_REQUIRE(axis >= -1 && axis <= 2, SSDR_ERR_INVALID, "axis must be -1 (no slab), 0, 1 or 2");
    grid::Slab slab;
    slab.axis = axis;
    slab.lo = layer_lo;
    slab.hi = layer_hi;
    grid::Inputs in;
    in.p = d_points;
    in.f = d_feats;
    in.c = d_classes;
    SSDR_TRY(grid::check_args(in, N, fdim, ldim, sampleDl, SSDR_GRID_ORDER_KEY, slab, bbox));
    SSDR_REQUIRE(capacity >= N, SSDR_ERR_INVALID,
                 "output capacity %zu is below the worst case of one voxel per point (%zu rows)", capacity, N);
    SSDR_REQUIRE((!fdim || d_feats_out) && (!ldim || d_classes_out), SSDR_ERR_INVALID, "NULL output array");
    grid::Outputs out;
    out.p = d_points_out;
    out.f = d_feats_out;
    out.c = d_classes_out;
    // keys / counts are optional for the caller; the kernels always write them
    SSDR_TRY(c->ws[grid::WS_TEMP].reserve(N * (sizeof(unsigned long long) + sizeof(int))));
    out.k = d_keys_out ? reinterpret_cast<unsigned long long*>(d_keys_out) : c->ws[grid::WS_TEMP].as<unsigned long long>();
    out.n = d_counts_out ? d_counts_out : reinterpret_cast<int*>(c->ws[grid::WS_TEMP].as<unsigned long long>() + N);
    grid::Meta hm;
    SSDR_TRY(grid::run_core(c, s, in, N, fdim, ldim, sampleDl, slab, bbox, out, &hm, nullptr));
    if (!(slab.axis >= 0 && hm.n_sel == 0)) SSDR_REQUIRE(hm.M >= 1 && hm.M <= N, SSDR_ERR_EMPTY, "Error");
    *M_out = (size_t)hm.M;
    return SSDR_OK;
}

int ssdr_grid_bbox_dev(const float* d_points, size_t N, void* stream, float* bbox_out) {
    SSDR_REQUIRE(d_points && bbox_out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_EMPTY, "Error");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    SSDR_TRY(c->ws[grid::WS_META].reserve(sizeof(grid::Meta)));
    grid::Meta* meta = c->ws[grid::WS_META].as<grid::Meta>();
    SSDR_TRY(grid::geometry(c, s, d_points, N, 1.0f, nullptr, meta));
    grid::Meta hm;
    SSDR_TRY(d2h_sync(c, &hm, meta, sizeof(grid::Meta), s));
    for (int d = 0; d < 3; ++d) {
        bbox_out[d] = hm.mn[d];
        bbox_out[3 + d] = hm.mx[d];
    }
    return SSDR_OK;
}

int ssdr_grid_point_layers_dev(const float* d_points, size_t N, const float* bbox, float sampleDl, int axis,
                               int32_t* d_layers, void* stream, unsigned long long* n_layers_out) {
    SSDR_REQUIRE(d_points && d_layers, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_EMPTY, "Error");
    SSDR_REQUIRE(axis >= 0 && axis <= 2, SSDR_ERR_INVALID, "axis must be 0, 1 or 2");
    SSDR_REQUIRE(sampleDl > 0.0f, SSDR_ERR_INVALID, "sampleDl must be positive");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    SSDR_TRY(c->ws[grid::WS_META].reserve(sizeof(grid::Meta)));
    grid::Meta* meta = c->ws[grid::WS_META].as<grid::Meta>();
    SSDR_TRY(grid::geometry(c, s, d_points, N, sampleDl, bbox, meta));
    grid::point_layers_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(d_points, N, meta, axis, d_layers);
    SSDR_CHECK_CUDA(cudaGetLastError());
    if (n_layers_out) {
        grid::Meta hm;
        SSDR_TRY(d2h_sync(c, &hm, meta, sizeof(grid::Meta), s));
        *n_layers_out = axis == 0 ? hm.nX : axis == 1 ? hm.nY : hm.nZ;
    }
    return SSDR_OK;
}

int ssdr_grid_layer_hist_dev(const float* d_points, size_t N, const float* bbox, float sampleDl, int axis,
                             unsigned long long* d_hist, size_t n_layers, size_t sample_stride, void* stream) {
    SSDR_REQUIRE(d_points && d_hist && bbox, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1 && n_layers >= 1, SSDR_ERR_EMPTY, "Error");
    SSDR_REQUIRE(axis >= 0 && axis <= 2, SSDR_ERR_INVALID, "axis must be 0, 1 or 2");
    SSDR_REQUIRE(sampleDl > 0.0f, SSDR_ERR_INVALID, "sampleDl must be positive");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    SSDR_TRY(c->ws[grid::WS_META].reserve(sizeof(grid::Meta)));
    grid::Meta* meta = c->ws[grid::WS_META].as<grid::Meta>();
    SSDR_TRY(grid::geometry(c, s, d_points, N, sampleDl, bbox, meta));
    SSDR_CHECK_CUDA(cudaMemsetAsync(d_hist, 0, n_layers * sizeof(unsigned long long), s));
    grid::layer_hist_kernel<<<(unsigned)c->sm_count * 8, 256, 0, s>>>(d_points, N, meta, axis, d_hist, n_layers,
                                                                      sample_stride ? sample_stride : 1);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

int ssdr_grid_route_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N, size_t fdim,
                        size_t ldim, float sampleDl, const float* bbox, int axis, const unsigned long long* bounds,
                        int world, float* d_points_out, float* d_feats_out, int32_t* d_classes_out,
                        unsigned long long* counts_out, void* stream) {
    SSDR_REQUIRE(d_points && bbox && bounds && d_points_out && counts_out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(world >= 1 && world <= grid::MAX_ROUTE_WORLD, SSDR_ERR_INVALID, "world must be in [1, %d]",
                 grid::MAX_ROUTE_WORLD);
    SSDR_REQUIRE(axis >= 0 && axis <= 2, SSDR_ERR_INVALID, "axis must be 0, 1 or 2");
    SSDR_REQUIRE(N < 0x7FFFFFFFull, SSDR_ERR_UNSUPPORTED, "N=%zu exceeds 2^31-2 points per call", N);
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    for (int r = 0; r < world; ++r) counts_out[r] = 0;
    if (N == 0) return SSDR_OK;
    if (!d_feats) fdim = 0;
    if (!d_classes) ldim = 0;
    typedef unsigned long long KeyT;
    SSDR_TRY(c->ws[grid::WS_META].reserve(sizeof(grid::Meta)));
    SSDR_TRY(c->ws[grid::WS_KEYS].reserve(N * sizeof(KeyT)));
    SSDR_TRY(c->ws[grid::WS_KEYS2].reserve(N * sizeof(KeyT)));
    SSDR_TRY(c->ws[grid::WS_IDX].reserve(N * sizeof(unsigned)));
    SSDR_TRY(c->ws[grid::WS_IDX2].reserve(N * sizeof(unsigned)));
    SSDR_TRY(c->ws[grid::WS_CTL].reserve(1024));
    const size_t scr = (prim::rs_scratch_words(N) + 8) * sizeof(unsigned);
    SSDR_TRY(c->ws[grid::WS_TEMP].reserve(scr));
    grid::Meta* meta = c->ws[grid::WS_META].as<grid::Meta>();
    SSDR_TRY(grid::geometry(c, s, d_points, N, sampleDl, bbox, meta));
    grid::RouteBounds rb;
    rb.world = world;
    for (int r = 0; r <= world; ++r) rb.b[r] = bounds[r];
    unsigned long long* d_counts = c->ws[grid::WS_CTL].as<unsigned long long>();
    SSDR_CHECK_CUDA(cudaMemsetAsync(d_counts, 0, grid::MAX_ROUTE_WORLD * sizeof(unsigned long long), s));
    SSDR_CHECK_CUDA(cudaMemsetAsync(c->ws[grid::WS_TEMP].p, 0, scr, s));
    KeyT* ka = c->ws[grid::WS_KEYS].as<KeyT>();
    KeyT* kb = c->ws[grid::WS_KEYS2].as<KeyT>();
    unsigned* va = c->ws[grid::WS_IDX].as<unsigned>();
    unsigned* vb = c->ws[grid::WS_IDX2].as<unsigned>();
    grid::route_key_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(d_points, N, meta, axis, rb, ka, va, d_counts);
    int cur = 0;  // one stable 8-bit pass orders the rows by destination and keeps the input order inside each
    SSDR_TRY(prim::radix_sort_pairs(ka, va, kb, vb, N, 8, c->ws[grid::WS_TEMP].as<unsigned>(), &cur, s));
    const unsigned* perm = cur ? vb : va;
    const unsigned long long tot3 = (unsigned long long)N * 3;
    grid::permute_rows_kernel<float><<<(unsigned)((tot3 + 255) / 256), 256, 0, s>>>(d_points, d_points_out, perm, N, 3);
    if (fdim && d_feats_out)
        grid::permute_rows_kernel<float><<<(unsigned)((N * fdim + 255) / 256), 256, 0, s>>>(d_feats, d_feats_out, perm, N, (int)fdim);
    if (ldim && d_classes_out)
        grid::permute_rows_kernel<int><<<(unsigned)((N * ldim + 255) / 256), 256, 0, s>>>(d_classes, d_classes_out, perm, N, (int)ldim);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return d2h_sync(c, counts_out, d_counts, (size_t)world * sizeof(unsigned long long), s);
}

int ssdr_grid_subsample_typed(const float* points, const void* feats, int feats_dtype, const void* classes,
                              int classes_dtype, size_t N, size_t fdim, size_t ldim, float sampleDl, int order,
                              size_t* M_out, void** handle) {
    SSDR_REQUIRE(points && M_out && handle, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_EMPTY, "Error");
    SSDR_REQUIRE(feats_dtype == SSDR_DTYPE_NATIVE || feats_dtype == SSDR_DTYPE_U8, SSDR_ERR_INVALID,
                 "feats_dtype must be SSDR_DTYPE_NATIVE (float32) or SSDR_DTYPE_U8");
    SSDR_REQUIRE(classes_dtype == SSDR_DTYPE_NATIVE || classes_dtype == SSDR_DTYPE_U8, SSDR_ERR_INVALID,
                 "classes_dtype must be SSDR_DTYPE_NATIVE (int32) or SSDR_DTYPE_U8");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    if (!feats) fdim = 0;
    if (!classes) ldim = 0;
    cudaStream_t s = c->stream;
    // the bytes travel as they are (uint8 colours / labels: 4x less PCIe than the float32 / int32 the reference's
    // wrapper widens them to on the host, wrapper.cpp:100-106); the reduce kernel widens on the fly (exact)
    grid::Inputs in;
    SSDR_TRY(c->ws[grid::WS_IN_P].reserve(N * 3 * sizeof(float)));
    SSDR_TRY(h2d(c, c->ws[grid::WS_IN_P].p, points, N * 3 * sizeof(float), s));
    in.p = c->ws[grid::WS_IN_P].as<float>();
    if (fdim) {
        in.f_u8 = feats_dtype == SSDR_DTYPE_U8;
        const size_t bytes = N * fdim * (in.f_u8 ? 1 : sizeof(float));
        SSDR_TRY(c->ws[grid::WS_IN_F].reserve(bytes));
        SSDR_TRY(h2d(c, c->ws[grid::WS_IN_F].p, feats, bytes, s));
        in.f = c->ws[grid::WS_IN_F].p;
    }
    if (ldim) {
        in.c_u8 = classes_dtype == SSDR_DTYPE_U8;
        const size_t bytes = N * ldim * (in.c_u8 ? 1 : sizeof(int));
        SSDR_TRY(c->ws[grid::WS_IN_C].reserve(bytes));
        SSDR_TRY(h2d(c, c->ws[grid::WS_IN_C].p, classes, bytes, s));
        in.c = c->ws[grid::WS_IN_C].p;
    }
    return grid::run_dev(c, s, in, N, fdim, ldim, sampleDl, order, M_out, handle);
}

int ssdr_grid_subsample(const float* points, const float* feats, const int32_t* classes, size_t N, size_t fdim,
                        size_t ldim, float sampleDl, int order, size_t* M_out, void** handle) {
    return ssdr_grid_subsample_typed(points, feats, SSDR_DTYPE_NATIVE, classes, SSDR_DTYPE_NATIVE, N, fdim, ldim, sampleDl,
                                     order, M_out, handle);
}

int ssdr_grid_fetch_ex(void* handle, float* points_out, float* feats_out, int32_t* classes_out, uint64_t* keys_out,
                       int32_t* counts_out) {
    SSDR_REQUIRE(handle, SSDR_ERR_INVALID, "NULL handle");
    grid::Handle* h = (grid::Handle*)handle;
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    cudaStream_t s = h->stream;
    if (points_out) SSDR_CHECK_CUDA(cudaMemcpyAsync(points_out, h->d_p, h->M * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (feats_out && h->d_f)
        SSDR_CHECK_CUDA(cudaMemcpyAsync(feats_out, h->d_f, h->M * h->fdim * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (classes_out && h->d_c)
        SSDR_CHECK_CUDA(cudaMemcpyAsync(classes_out, h->d_c, h->M * h->ldim * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (keys_out) SSDR_CHECK_CUDA(cudaMemcpyAsync(keys_out, h->d_k, h->M * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (counts_out) SSDR_CHECK_CUDA(cudaMemcpyAsync(counts_out, h->d_n, h->M * sizeof(int), cudaMemcpyDeviceToHost, s));
    SSDR_CHECK_CUDA(cudaStreamSynchronize(s));
    return SSDR_OK;
}

int ssdr_grid_fetch(void* handle, float* points_out, float* feats_out, int32_t* classes_out) {
    return ssdr_grid_fetch_ex(handle, points_out, feats_out, classes_out, nullptr, nullptr);
}

int ssdr_grid_dev_ptrs(void* handle, const float** d_points, const float** d_feats, const int32_t** d_classes) {
    SSDR_REQUIRE(handle, SSDR_ERR_INVALID, "NULL handle");
    grid::Handle* h = (grid::Handle*)handle;
    if (d_points) *d_points = h->d_p;
    if (d_feats) *d_feats = h->d_f;
    if (d_classes) *d_classes = h->d_c;
    return SSDR_OK;
}

int ssdr_grid_debug_timing(uint64_t* marks16, int* key_bits) {
    SSDR_REQUIRE(marks16, SSDR_ERR_INVALID, "NULL pointer");
    for (int k = 0; k < 16; ++k) marks16[k] = grid::g_last_meta.tmark[k];
    if (key_bits) *key_bits = grid::g_last_meta.key_bits;
    return SSDR_OK;
}

int ssdr_grid_free(void* handle) {
    grid::free_handle((grid::Handle*)handle);
    return SSDR_OK;
}
}
