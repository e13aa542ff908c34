// knn.cu -- exact k nearest neighbours with nanoflann-identical results on sm_100a.
//
// Replaces cpp_knn / cpp_knn_omp / cpp_knn_batch / cpp_knn_batch_omp (utils/nearest_neighbors/knn_.cxx:22-135), which
// build a nanoflann KD-tree per cloud (twice) and answer queries one by one on the CPU.
//
// Device pipeline (all batch items in the same launches):
//   A. uniform cell grid per item: bbox -> cell edge from the expected K-NN radius -> counting sort of the points
//      (and of the queries, for warp coherence) into cells (hand-written scan, primitives.cuh).
//   B. main kernel, one thread per (cell-sorted) query: scan the 3x3x3 block, then Chebyshev shells, keeping the
//      exact top-(K+1) under the order (fp32 squared distance, index) in registers as packed 64-bit keys; stop when
//      the (K+1)-th distance is provably inside the scanned block.  Distances use the reference's operation order
//      ((dx*dx + dy*dy) + dz*dz, no FMA: nanoflann.hpp:300-303).
//   C. nanoflann's own order differs from (distance, index) ONLY when the top-(K+1) holds equal or almost equal
//      distances (tree-visit order decides ties, nanoflann.hpp:72-96,1317; pruning compares differently rounded
//      sums).  Such rows are flagged by B and re-resolved by an exact replay of the nanoflann tree (kdtree.cuh).
#include <atomic>
#include <mutex>
#include <vector>

#include <time.h>
#include "common.cuh"
#include "kdtree.cuh"
#include "primitives.cuh"

namespace ssdr {
namespace knn {

struct ItemMeta {
    float lo[3], hi[3];
    float h, inv_h, eps_abs;
    int g[3];
    unsigned cell_base;  // offset of this item's cells in the flat cell arrays
};

enum { WS_ENC = 0, WS_ITEMS = 1, WS_CELL_P = 2, WS_CELL_Q = 3, WS_CNT_P = 4, WS_CNT_Q = 5, WS_START_P = 6,
       WS_START_Q = 7, WS_SORT_P = 8, WS_SORT_Q = 9, WS_FLAGS = 10, WS_TEMP = 11, WS_IN_P = 12, WS_IN_Q = 13,
       WS_OUT = 14, WS_STATS = 15, WS_TREE = 16 /* .. WS_TREE+5 used by kdtree.cuh */, WS_PATCH = 22, WS_ASYNC = 23 };

// Calls enqueued back to back without a host round trip (the pyramid): the tree workspace travels from call to call and
// errors of the tie path accumulate in a persistent device word that ssdr_knn_status reads.
static thread_local unsigned long long g_last_launches = 0;  // kernels launched by this thread's last pyramid call

struct AsyncCtx {
    kdtree::Tree tree;
    bool reuse = false;      // this call's support cloud is the previous call's: trees are built at most once
    unsigned* status = nullptr;
};

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// ---- A1+A2: bounding box per item, then (last block) the cell grid geometry --------------------------------------
// enc holds order-preserving encodings combined with atomicMax and identity 0: [0..2] = ~ord(min), [3..5] = ord(max),
// so one memset prepares it together with the counters.
// Occupancy probe: non-empty cubes of the bbox at two nested resolutions (r2 cubes along the longest edge, and half of
// that).  Their ratio is the cloud's box-counting dimension d at the scale that matters -- ~3 for points filling a
// volume, ~2 for scanned surfaces, which is what the real inputs are -- and n2 anchors the density, so the cell edge
// can be chosen for a target number of points per NON-EMPTY cell instead of per bbox volume (for surfaces the latter
// puts tens of points into every occupied cell and multiplies the distance evaluations per query).
struct Probe {
    unsigned r2;      // 0: no probe (small clouds), volumetric rule
    unsigned n1, n2;  // non-empty cubes at edge 2*E/r2 and E/r2
};
constexpr unsigned PROBE_MIN_POINTS = 4096;
constexpr unsigned PROBE_MAX_R = 64;
constexpr unsigned PROBE_SAMPLE = 1u << 18;  // points looked at per item
constexpr unsigned PROBE_WORDS = PROBE_MAX_R * PROBE_MAX_R * PROBE_MAX_R / 32 + (PROBE_MAX_R / 2) * (PROBE_MAX_R / 2) * (PROBE_MAX_R / 2) / 32;

__host__ __device__ inline unsigned probe_resolution(unsigned N) {
    if (N < PROBE_MIN_POINTS) return 0;
    unsigned r = (unsigned)cbrtf((float)N / 16.f);  // ~16 points per cube if the cloud filled its bbox
    r &= ~1u;                                       // even, so that the coarse cubes are unions of 8 fine ones
    return r < 4 ? 4 : (r > PROBE_MAX_R ? PROBE_MAX_R : r);
}

__device__ void setup_item(const unsigned* __restrict__ enc, ItemMeta* __restrict__ items, int b, unsigned N,
                           float occupancy, unsigned cell_cap, unsigned cstride, Probe pr = Probe{0, 0, 0}) {
    ItemMeta m;
    float ext[3], E = 0.f, amax = 0.f;
    for (int d = 0; d < 3; ++d) {
        m.lo[d] = ord2f(~__ldcg(&enc[b * 6 + d]));
        m.hi[d] = ord2f(__ldcg(&enc[b * 6 + 3 + d]));
        ext[d] = m.hi[d] - m.lo[d];
        E = fmaxf(E, ext[d]);
        amax = fmaxf(amax, fmaxf(fabsf(m.lo[d]), fabsf(m.hi[d])));
    }
    if (!(E > 0.f) || !isfinite(E)) {
        m.h = 1.f;
        m.g[0] = m.g[1] = m.g[2] = 1;
    } else {
        float vol = 1.f;
        for (int d = 0; d < 3; ++d) vol *= fmaxf(ext[d], E * 1e-3f);
        float h = cbrtf(vol * occupancy / (float)N);
        if (pr.r2 && pr.n1 >= 1 && pr.n2 >= pr.n1) {
            float dim = log2f((float)pr.n2 / (float)pr.n1);
            dim = fminf(fmaxf(dim, 1.f), 3.f);
            const float per_cube = (float)N / (float)pr.n2;          // points per non-empty cube of edge E / r2
            const float target = occupancy * (1.f + 0.75f * (3.f - dim));  // flatter clouds: fewer occupied neighbours
            h = (E / (float)pr.r2) * powf(target / per_cube, 1.f / dim);
        }
        h = fmaxf(h, E * (1.f / 1000.f));
        for (;;) {
            double cells = 1.0;
            for (int d = 0; d < 3; ++d) {
                int g = (int)floorf(ext[d] / h) + 1;
                m.g[d] = g < 1 ? 1 : g;
                cells *= (double)m.g[d];
            }
            if (cells <= (double)cell_cap) break;
            h *= 1.26f;
        }
        m.h = h;
    }
    m.inv_h = 1.f / m.h;
    m.eps_abs = 1e-3f * m.h + 1e-6f * amax;
    m.cell_base = (unsigned)b * cstride;
    items[b] = m;
}

__global__ void bbox_setup_kernel(const float* __restrict__ pts, unsigned N, unsigned* __restrict__ enc,
                                  unsigned* __restrict__ ticket, ItemMeta* __restrict__ items, int B, float occupancy,
                                  unsigned cell_cap, unsigned cstride, bool defer_setup) {
    const unsigned b = blockIdx.y;
    const float* p = pts + (size_t)b * N * 3;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = __ldg(p + 3 * (size_t)i + d);
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (mn[d] != INFINITY) atomicMax(&enc[b * 6 + d], ~f2ord(mn[d]));
            if (mx[d] != -INFINITY) atomicMax(&enc[b * 6 + 3 + d], f2ord(mx[d]));
        }
    }
    if (defer_setup) return;  // probe_kernel derives the geometry
    // last block done -> geometry of every item
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int it = threadIdx.x; it < B; it += blockDim.x)
            setup_item(enc, items, it, N, occupancy, cell_cap, cstride);
    }
}

__device__ __forceinline__ int cell_coord(float x, float lo, float inv_h, int g) {
    int c = (int)floorf((x - lo) * inv_h);
    return c < 0 ? 0 : (c >= g ? g - 1 : c);
}

// ---- A3: counting sort into cells; points and (when they differ) queries in the same launches ---------------------
struct SortJob {
    const float* xyz;     // (B, n, 3)
    unsigned n, total;    // per item, all items
    unsigned* cell_of;    // [total]
    unsigned* counts;     // [ncell]
    const unsigned* starts;  // [ncell] exclusive scan of counts (this job's block of the concatenated scan)
    float4* sorted;       // [total]
    unsigned start_bias;  // subtracted from starts (the queries' counts are scanned behind the points')
};

__global__ void cell_count_kernel(SortJob jp, SortJob jq, const ItemMeta* __restrict__ items) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool isq = i >= jp.total;
    const SortJob& j = isq ? jq : jp;
    if (isq) i -= jp.total;
    if (i >= j.total) return;
    const ItemMeta& m = items[i / j.n];
    const float x = __ldg(j.xyz + 3 * (size_t)i), y = __ldg(j.xyz + 3 * (size_t)i + 1), z = __ldg(j.xyz + 3 * (size_t)i + 2);
    const int cx = cell_coord(x, m.lo[0], m.inv_h, m.g[0]);
    const int cy = cell_coord(y, m.lo[1], m.inv_h, m.g[1]);
    const int cz = cell_coord(z, m.lo[2], m.inv_h, m.g[2]);
    const unsigned cell = m.cell_base + (unsigned)((cz * m.g[1] + cy) * m.g[0] + cx);
    j.cell_of[i] = cell;
    atomicAdd(&j.counts[cell], 1u);
}
__global__ void cell_scatter_kernel(SortJob jp, SortJob jq) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool isq = i >= jp.total;
    const SortJob& j = isq ? jq : jp;
    if (isq) i -= jp.total;
    if (i >= j.total) return;
    const unsigned cell = j.cell_of[i];
    // the counts are consumed as slot tickets (no separate cursor array to zero): order inside a cell is arbitrary
    const unsigned pos = j.starts[cell] - j.start_bias + (atomicSub(&j.counts[cell], 1u) - 1u);
    float4 v;
    v.x = __ldg(j.xyz + 3 * (size_t)i);
    v.y = __ldg(j.xyz + 3 * (size_t)i + 1);
    v.z = __ldg(j.xyz + 3 * (size_t)i + 2);
    v.w = __int_as_float((int)(i % j.n));  // index inside the item
    j.sorted[pos] = v;
}

// cube index of a point at the probe's fine resolution (edge = E / r2, anchored at the bbox minimum)
__device__ __forceinline__ void probe_cubes(float x, float y, float z, const float lo[3], float inv_edge, unsigned r2,
                                            unsigned* fine, unsigned* coarse) {
    const unsigned cx = min((unsigned)fmaxf((x - lo[0]) * inv_edge, 0.f), r2 - 1);
    const unsigned cy = min((unsigned)fmaxf((y - lo[1]) * inv_edge, 0.f), r2 - 1);
    const unsigned cz = min((unsigned)fmaxf((z - lo[2]) * inv_edge, 0.f), r2 - 1);
    const unsigned r1 = r2 >> 1;
    *fine = (cz * r2 + cy) * r2 + cx;
    *coarse = ((cz >> 1) * r1 + (cy >> 1)) * r1 + (cx >> 1);
}

// multi-launch path: marks the cubes of every item in global bitmaps; new bits are counted per block, the last block
// to finish derives every item's geometry (what bbox_setup_kernel's last block does when there is no probe)
__global__ void probe_kernel(const float* __restrict__ pts, unsigned N, const unsigned* __restrict__ enc,
                             unsigned* __restrict__ bitmaps /* [B][PROBE_WORDS] */,
                             unsigned* __restrict__ counts /* [B][2] */, unsigned* __restrict__ ticket,
                             ItemMeta* __restrict__ items, int B, float occupancy, unsigned cell_cap, unsigned cstride,
                             unsigned r2, unsigned stride /* probe every stride-th point (large clouds) */) {
    const unsigned b = blockIdx.y;
    const float* p = pts + (size_t)b * N * 3;
    float lo[3], E = 0.f;
    for (int d = 0; d < 3; ++d) {
        lo[d] = ord2f(~__ldcg(&enc[b * 6 + d]));
        E = fmaxf(E, ord2f(__ldcg(&enc[b * 6 + 3 + d])) - lo[d]);
    }
    unsigned* fine_bm = bitmaps + (size_t)b * PROBE_WORDS;
    unsigned* coarse_bm = fine_bm + PROBE_MAX_R * PROBE_MAX_R * PROBE_MAX_R / 32;
    unsigned new1 = 0, new2 = 0;
    if (E > 0.f && isfinite(E)) {
        const float inv_edge = (float)r2 / E;
        const unsigned ns = (N + stride - 1) / stride;
        for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < ns; j += gridDim.x * blockDim.x) {
            const size_t i = (size_t)j * stride;
            unsigned f, c2;
            probe_cubes(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2), lo, inv_edge, r2, &f, &c2);
            const unsigned fb = 1u << (f & 31), cb = 1u << (c2 & 31);
            if (!(__ldcg(&fine_bm[f >> 5]) & fb)) new2 += (atomicOr(&fine_bm[f >> 5], fb) & fb) ? 0u : 1u;
            if (!(__ldcg(&coarse_bm[c2 >> 5]) & cb)) new1 += (atomicOr(&coarse_bm[c2 >> 5], cb) & cb) ? 0u : 1u;
        }
    }
    __shared__ unsigned s_new[2];
    __shared__ bool s_last;
    if (threadIdx.x < 2) s_new[threadIdx.x] = 0;
    __syncthreads();
    new1 = __reduce_add_sync(0xffffffffu, new1);
    new2 = __reduce_add_sync(0xffffffffu, new2);
    if ((threadIdx.x & 31) == 0) {
        if (new1) atomicAdd(&s_new[0], new1);
        if (new2) atomicAdd(&s_new[1], new2);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_new[0]) atomicAdd(&counts[b * 2], s_new[0]);
        if (s_new[1]) atomicAdd(&counts[b * 2 + 1], s_new[1]);
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int it = threadIdx.x; it < B; it += blockDim.x)
            setup_item(enc, items, it, N, occupancy, cell_cap, cstride,
                       Probe{r2, __ldcg(&counts[it * 2]), __ldcg(&counts[it * 2 + 1])});
    }
}

// ---- A (small clouds): the whole cell-grid build of one item in ONE CTA -- bbox, geometry, cell histogram and scan
// in shared memory, scatter -- instead of six launches; the lower pyramid levels are pure launch latency otherwise.
constexpr int SG_THREADS = 512;
constexpr unsigned SG_MAX_CELLS = 40960;  // cstride limit (160 KB of dynamic shared memory)
constexpr unsigned SG_MAX_POINTS = 16384;

__device__ __forceinline__ unsigned local_cell(const ItemMeta& m, float x, float y, float z) {
    const int cx = cell_coord(x, m.lo[0], m.inv_h, m.g[0]);
    const int cy = cell_coord(y, m.lo[1], m.inv_h, m.g[1]);
    const int cz = cell_coord(z, m.lo[2], m.inv_h, m.g[2]);
    return (unsigned)((cz * m.g[1] + cy) * m.g[0] + cx);
}

// exclusive scan of a[0..n) in shared memory by the whole CTA (every thread owns a contiguous chunk)
__device__ void block_exclusive_scan(unsigned* a, unsigned n, unsigned* warp_tot /* [SG_THREADS/32 + 1] */) {
    const unsigned per = (n + SG_THREADS - 1) / SG_THREADS;
    const unsigned lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    unsigned sum = 0;
    for (unsigned i = lo; i < hi; ++i) sum += a[i];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (unsigned)o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned w = threadIdx.x < SG_THREADS / 32 ? warp_tot[threadIdx.x] : 0u;
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= (unsigned)o) wi += v;
        }
        if (threadIdx.x < SG_THREADS / 32) warp_tot[threadIdx.x] = wi - w;
    }
    __syncthreads();
    unsigned run = warp_tot[threadIdx.x >> 5] + incl - sum;
    for (unsigned i = lo; i < hi; ++i) {
        const unsigned v = a[i];
        a[i] = run;
        run += v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SG_THREADS) small_grid_kernel(const float* __restrict__ pts, unsigned N,
                                                                const float* __restrict__ qs, unsigned Q,
                                                                unsigned* __restrict__ enc, ItemMeta* __restrict__ items,
                                                                int B, float occupancy, unsigned cell_cap,
                                                                unsigned cstride, unsigned* __restrict__ starts_p,
                                                                float4* __restrict__ sort_p, float4* __restrict__ sort_q,
                                                                unsigned r2) {
    extern __shared__ unsigned sg_cells[];  // [max(cstride, PROBE_WORDS when probing)]
    __shared__ float s_red[6][SG_THREADS / 32];
    __shared__ unsigned s_wtot[SG_THREADS / 32 + 1];
    __shared__ ItemMeta s_m;
    __shared__ float s_box[4];
    __shared__ unsigned s_new[2];
    const unsigned b = blockIdx.x;
    const float* p = pts + (size_t)b * N * 3;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = threadIdx.x; i < N; i += SG_THREADS) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = __ldg(p + 3 * (size_t)i + d);
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
        }
    if ((threadIdx.x & 31) == 0)
        for (int d = 0; d < 3; ++d) {
            s_red[d][threadIdx.x >> 5] = mn[d];
            s_red[3 + d][threadIdx.x >> 5] = mx[d];
        }
    if (r2)
        for (unsigned i = threadIdx.x; i < PROBE_WORDS; i += SG_THREADS) sg_cells[i] = 0;  // the probe's two bitmaps
    __syncthreads();
    if (threadIdx.x == 0) {
        float E = 0.f;
        for (int d = 0; d < 3; ++d) {
            float a = s_red[d][0], z = s_red[3 + d][0];
            for (int w = 1; w < SG_THREADS / 32; ++w) {
                a = fminf(a, s_red[d][w]);
                z = fmaxf(z, s_red[3 + d][w]);
            }
            enc[b * 6 + d] = ~f2ord(a);  // same encoding the multi-launch path reduces with atomicMax
            enc[b * 6 + 3 + d] = f2ord(z);
            s_box[d] = a;
            E = fmaxf(E, z - a);
        }
        s_box[3] = E;
        s_new[0] = s_new[1] = 0;
    }
    __syncthreads();
    Probe pr{0, 0, 0};
    if (r2 && s_box[3] > 0.f && isfinite(s_box[3])) {  // occupied cubes at two resolutions (see Probe)
        const float lo[3] = {s_box[0], s_box[1], s_box[2]};
        const float inv_edge = (float)r2 / s_box[3];
        unsigned* fine_bm = sg_cells;
        unsigned* coarse_bm = sg_cells + PROBE_MAX_R * PROBE_MAX_R * PROBE_MAX_R / 32;
        unsigned new1 = 0, new2 = 0;
        for (unsigned i = threadIdx.x; i < N; i += SG_THREADS) {
            unsigned f, c2;
            probe_cubes(__ldg(p + 3 * (size_t)i), __ldg(p + 3 * (size_t)i + 1), __ldg(p + 3 * (size_t)i + 2), lo, inv_edge,
                        r2, &f, &c2);
            const unsigned fb = 1u << (f & 31), cb = 1u << (c2 & 31);
            new2 += (atomicOr(&fine_bm[f >> 5], fb) & fb) ? 0u : 1u;
            new1 += (atomicOr(&coarse_bm[c2 >> 5], cb) & cb) ? 0u : 1u;
        }
        new1 = __reduce_add_sync(0xffffffffu, new1);
        new2 = __reduce_add_sync(0xffffffffu, new2);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&s_new[0], new1);
            atomicAdd(&s_new[1], new2);
        }
        __syncthreads();
        pr = Probe{r2, s_new[0], s_new[1]};
    }
    if (threadIdx.x == 0) {
        setup_item(enc, items, (int)b, N, occupancy, cell_cap, cstride, pr);
        s_m = items[b];
    }
    __syncthreads();
    const ItemMeta m = s_m;
    const unsigned ncell = (unsigned)(m.g[0] * m.g[1] * m.g[2]) + 1u;  // cells in use (+1: end of the last one), <= cstride
    for (unsigned i = threadIdx.x; i < ncell; i += SG_THREADS) sg_cells[i] = 0;
    __syncthreads();
    // support points: histogram -> scan -> global starts (offset by the item's base) -> scatter
    for (unsigned i = threadIdx.x; i < N; i += SG_THREADS)
        atomicAdd(&sg_cells[local_cell(m, __ldg(p + 3 * (size_t)i), __ldg(p + 3 * (size_t)i + 1), __ldg(p + 3 * (size_t)i + 2))], 1u);
    __syncthreads();
    block_exclusive_scan(sg_cells, ncell, s_wtot);
    for (unsigned i = threadIdx.x; i < ncell; i += SG_THREADS) starts_p[(size_t)b * cstride + i] = b * N + sg_cells[i];
    if (b == (unsigned)B - 1 && threadIdx.x == 0) starts_p[(size_t)B * cstride] = (unsigned)B * N;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < N; i += SG_THREADS) {
        float4 v;
        v.x = __ldg(p + 3 * (size_t)i);
        v.y = __ldg(p + 3 * (size_t)i + 1);
        v.z = __ldg(p + 3 * (size_t)i + 2);
        v.w = __int_as_float((int)i);
        sort_p[(size_t)b * N + atomicAdd(&sg_cells[local_cell(m, v.x, v.y, v.z)], 1u)] = v;
    }
    if (!qs) return;
    // queries (a different array): only their cell-sorted order is needed
    __syncthreads();
    const float* q = qs + (size_t)b * Q * 3;
    for (unsigned i = threadIdx.x; i < ncell; i += SG_THREADS) sg_cells[i] = 0;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < Q; i += SG_THREADS)
        atomicAdd(&sg_cells[local_cell(m, __ldg(q + 3 * (size_t)i), __ldg(q + 3 * (size_t)i + 1), __ldg(q + 3 * (size_t)i + 2))], 1u);
    __syncthreads();
    block_exclusive_scan(sg_cells, ncell, s_wtot);
    for (unsigned i = threadIdx.x; i < Q; i += SG_THREADS) {
        float4 v;
        v.x = __ldg(q + 3 * (size_t)i);
        v.y = __ldg(q + 3 * (size_t)i + 1);
        v.z = __ldg(q + 3 * (size_t)i + 2);
        v.w = __int_as_float((int)i);
        sort_q[(size_t)b * Q + atomicAdd(&sg_cells[local_cell(m, v.x, v.y, v.z)], 1u)] = v;
    }
}

// ---- B: main query kernel -------------------------------------------------------------------------------------
// Top list: distances and indices in separate register arrays, ordered by DISTANCE ONLY (equal distances keep
// arrival order).  That is enough: a row whose top-(K+1) holds two equal distances is flagged and re-resolved by the
// exact nanoflann replay anyway, and without such a tie the order by distance is the order by (distance, index).
// The branch-free insertion costs 2 FMNMX + 1 FSETP + 2 SEL per slot.
template <int KCAP>
__device__ __forceinline__ void list_insert(float (&dist)[KCAP], unsigned (&idx)[KCAP], float d, unsigned id) {
    bool p_prev = true;  // d < dist[j] for the slot above (the caller checked d < dist[KCAP-1])
#pragma unroll
    for (int j = KCAP - 1; j > 0; --j) {
        const bool p = d < dist[j - 1];
        const float nd = fmaxf(dist[j - 1], fminf(dist[j], d));
        idx[j] = p ? idx[j - 1] : (p_prev ? id : idx[j]);
        dist[j] = nd;
        p_prev = p;
    }
    dist[0] = fminf(dist[0], d);
    idx[0] = p_prev ? id : idx[0];
}

// conservative squared distance from coordinate v to the slab of cell c along one axis (0 inside, minus the slack
// that covers fp32 rounding in the cell assignment)
__device__ __forceinline__ float slab_gap(float v, float lo, float h, int c, float eps) {
    const float a = lo + (float)c * h, b = lo + (float)(c + 1) * h;
    float g = fmaxf(fmaxf(a - v, v - b), 0.f) - eps;
    return g > 0.f ? g : 0.f;
}

// The kernel is WARP-SYNCHRONOUS: the 32 queries of a warp (neighbours in the cell-sorted order) walk the same
// sequence of cell rows together; a lane whose row is pruned or shorter simply idles.  That makes warp votes legal
// inside the scan, which is what the deferred insertion needs: a candidate that beats the lane's threshold is parked
// in a one-deep pending slot, and the (expensive, branch-free) list insertion runs only when SOME lane would overflow
// its slot -- then every lane with a parked candidate inserts at once.  Lanes therefore insert together instead of
// one at a time, which is where the thread-per-query version lost two thirds of its issue slots to divergence.
template <int KCAP, typename OutT>
__global__ void __launch_bounds__(128) query_kernel(const float4* __restrict__ spts, const float4* __restrict__ sq,
                                                    const unsigned* __restrict__ cell_start,
                                                    const ItemMeta* __restrict__ items, unsigned N, unsigned Q,
                                                    unsigned q_begin, unsigned q_end, int K, OutT* __restrict__ out,
                                                    unsigned* __restrict__ flag_count, unsigned* __restrict__ flag_list,
                                                    unsigned long long* __restrict__ evals) {
    constexpr unsigned FULL = 0xffffffffu;
    const unsigned t = q_begin + blockIdx.x * blockDim.x + threadIdx.x;  // this launch covers sorted rows [begin, end)
    const bool valid = t < q_end;
    const unsigned tq = valid ? t : q_end - 1;  // idle lanes shadow the last query and never emit
    unsigned my_evals = 0;
    const unsigned b = tq / Q;
    const ItemMeta m = items[b];
    const float4 q = __ldg(sq + tq);
    const unsigned qid = (unsigned)__float_as_int(q.w);
    const int gx = m.g[0], gy = m.g[1], gz = m.g[2];
    const int cx = cell_coord(q.x, m.lo[0], m.inv_h, gx);
    const int cy = cell_coord(q.y, m.lo[1], m.inv_h, gy);
    const int cz = cell_coord(q.z, m.lo[2], m.inv_h, gz);
    const unsigned* cs = cell_start + m.cell_base;
    float dist[KCAP];
    unsigned idx[KCAP];
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
        dist[j] = INFINITY;
        idx[j] = 0xFFFFFFFFu;
    }
    float pend_d = 0.f;
    unsigned pend_id = 0;
    bool has_pend = false;
    bool done = !valid;

    for (int r = 1; __any_sync(FULL, !done); ++r) {
        // rows of the shell at Chebyshev radius r (for r == 1: the whole 3x3 block, nearest rows first: own row, the
        // 4 face rows, the 4 diagonal rows); r, the row order and the segment count are warp-uniform
        const int w = 2 * r + 1, nrow = w * w;
        for (int k = 0; k < nrow; ++k) {
            int dz, dy;
            if (r == 1) {
                dz = (int)((0x2402049u >> (3 * k)) & 7u) - 1;  // packed (offset + 1), 3 bits per entry, order:
                dy = (int)((0x2081281u >> (3 * k)) & 7u) - 1;  // (0,0)(0,-1)(0,1)(-1,0)(1,0)(-1,-1)(-1,1)(1,-1)(1,1)
            } else {
                dz = k / w - r;
                dy = k % w - r;
            }
            const bool whole = (r == 1) || dz == -r || dz == r || dy == -r || dy == r;
            const int nseg = whole ? 1 : 2;
            for (int sg = 0; sg < nseg; ++sg) {
                // this lane's slice of the cell-sorted array for the row (empty when done, outside the grid, or when
                // even the row's nearest possible point cannot enter the list)
                unsigned s0 = 0, len = 0;
                if (!done) {
                    const int z = cz + dz, y = cy + dy;
                    int xa = whole ? cx - r : (sg == 0 ? cx - r : cx + r);
                    int xb = whole ? cx + r : xa;
                    xa = max(xa, 0);
                    xb = min(xb, gx - 1);
                    if (z >= 0 && z < gz && y >= 0 && y < gy && xa <= xb) {
                        const float gy_ = slab_gap(q.y, m.lo[1], m.h, y, m.eps_abs);
                        const float gz_ = slab_gap(q.z, m.lo[2], m.h, z, m.eps_abs);
                        float gx_ = 0.f;
                        if (cx < xa) gx_ = slab_gap(q.x, m.lo[0], m.h, xa, m.eps_abs);
                        else if (cx > xb) gx_ = slab_gap(q.x, m.lo[0], m.h, xb, m.eps_abs);
                        const float lower = (gx_ * gx_ + gy_ * gy_ + gz_ * gz_) * 0.99999f;
                        if (lower < dist[KCAP - 1]) {
                            const unsigned row = (unsigned)((z * gy + y) * gx);
                            s0 = cs[row + xa];
                            len = cs[row + xb + 1] - s0;
                        }
                    }
                }
                my_evals += len;
                const unsigned maxlen = __reduce_max_sync(FULL, len);
                for (unsigned i = 0; i < maxlen; ++i) {
                    const bool act = i < len;
                    float d = INFINITY;
                    unsigned id = 0;
                    if (act) {
                        const float4 p = __ldg(spts + s0 + i);
                        const float dx = __fsub_rn(q.x, p.x), dy_ = __fsub_rn(q.y, p.y), dz_ = __fsub_rn(q.z, p.z);
                        d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy_, dy_)), __fmul_rn(dz_, dz_));
                        id = (unsigned)__float_as_int(p.w);
                    }
                    bool pass = d < dist[KCAP - 1];  // inactive lanes carry +inf
                    if (__any_sync(FULL, pass && has_pend)) {
                        if (has_pend) {
                            list_insert<KCAP>(dist, idx, pend_d, pend_id);
                            has_pend = false;
                        }
                        pass = d < dist[KCAP - 1];
                    }
                    if (pass) {
                        pend_d = d;
                        pend_id = id;
                        has_pend = true;
                    }
                }
            }
        }
        // the shell is complete: settle the parked candidates, then test termination per lane
        if (__any_sync(FULL, has_pend)) {
            if (has_pend) {
                if (pend_d < dist[KCAP - 1]) list_insert<KCAP>(dist, idx, pend_d, pend_id);
                has_pend = false;
            }
        }
        if (!done) {
            const int x0 = max(cx - r, 0), x1 = min(cx + r, gx - 1);
            const int y0 = max(cy - r, 0), y1 = min(cy + r, gy - 1);
            const int z0 = max(cz - r, 0), z1 = min(cz + r, gz - 1);
            if (x0 == 0 && x1 == gx - 1 && y0 == 0 && y1 == gy - 1 && z0 == 0 && z1 == gz - 1) {
                done = true;  // the block covers the whole grid
            } else {
                // every unscanned point lies beyond a face of the block: lower-bound its distance
                float R = INFINITY;
                if (cx - r >= 0) R = fminf(R, q.x - (m.lo[0] + (float)(cx - r) * m.h));
                if (cx + r <= gx - 1) R = fminf(R, (m.lo[0] + (float)(cx + r + 1) * m.h) - q.x);
                if (cy - r >= 0) R = fminf(R, q.y - (m.lo[1] + (float)(cy - r) * m.h));
                if (cy + r <= gy - 1) R = fminf(R, (m.lo[1] + (float)(cy + r + 1) * m.h) - q.y);
                if (cz - r >= 0) R = fminf(R, q.z - (m.lo[2] + (float)(cz - r) * m.h));
                if (cz + r <= gz - 1) R = fminf(R, (m.lo[2] + (float)(cz + r + 1) * m.h) - q.z);
                R -= m.eps_abs;
                // the last kept entry (rank KCAP >= K+1) must be strictly inside the guaranteed radius (+inf while
                // fewer than KCAP candidates were seen)
                if (R > 0.f && dist[KCAP - 1] < R * R * 0.99999f) done = true;
            }
        }
    }

    // ---- emit + tie flags
    if (valid) {
        const int nvalid = (unsigned)K < N ? K : (int)N;
        OutT* o = out + ((size_t)b * Q + qid) * (size_t)K;
        bool flag = false;
#pragma unroll
        for (int j = 0; j < KCAP - 1; ++j) {
            if (j < nvalid) o[j] = (OutT)idx[j];
            if (j + 1 < nvalid && dist[j] == dist[j + 1]) flag = true;
            // boundary tie or near tie between rank K and K+1 (pruning-rounding hazard)
            if (j + 1 == K && N > (unsigned)K && dist[j + 1] <= dist[j] * 1.00001f) flag = true;
        }
        if (KCAP - 1 < nvalid) o[KCAP - 1] = (OutT)idx[KCAP - 1];
        if (flag) flag_list[atomicAdd(flag_count, 1u)] = b * Q + qid;
    }
#pragma unroll
    for (int mm = 16; mm > 0; mm >>= 1) my_evals += __shfl_xor_sync(FULL, my_evals, mm);
    if ((threadIdx.x & 31) == 0 && my_evals) atomicAdd(evals, (unsigned long long)my_evals);
}

// ---- B2: main query kernel for 8 < K <= 16: batched selection --------------------------------------------------
// Still one thread per query and warp-synchronous, but no per-candidate list insertion any more.  A candidate under
// the lane's threshold is appended to the lane's column of a shared-memory queue (one store); when some lane has 16
// queued, ALL lanes sort their batch with a 60-comparator network and fold it into their sorted 16-entry list with a
// bitonic merge -- straight-line code that every lane executes together, so no issue slot is lost to divergence
// (the one-deep deferred insertion above fires ~90 times per warp for ~10 lanes each; here a warp runs 5 merges).
// The list holds the 16 nearest; `m17` tracks the smallest distance among everything else that was seen, i.e. the
// 17th smallest: it is the queue's admission threshold, the row pruning bound, the termination bound and the
// boundary-tie detector, exactly what dist[16] is in query_kernel<17>.
// The candidates of a lane are a FLAT stream: the (start, length) pairs of up to Q16_DESC cell rows are computed by the
// whole warp together (all cell-start loads of the chunk in flight at once, every row's lines prefetched into L1) and
// parked in shared memory, then every lane walks its own pairs, four candidates per step, so that a lane with a short
// or pruned row does not idle until the longest row of the warp is done.
constexpr int Q16_L = 16;      // list entries kept sorted = batch size of a merge
constexpr int Q16_STEP = 4;    // candidates per lane and scan step
constexpr int Q16_CAP = Q16_L + Q16_STEP;  // queue rows: a step may overshoot a full batch by STEP - 1 entries
constexpr int Q16_DESC = 8;    // row descriptors per lane and chunk
constexpr int Q16_THREADS = 128;

#define SSDR_CE(a, b)                                   \
    {                                                   \
        const float x_ = d[a], y_ = d[b];               \
        const bool p_ = y_ < x_;                        \
        const unsigned ia_ = i[a], ib_ = i[b];          \
        d[a] = fminf(x_, y_);                           \
        d[b] = fmaxf(x_, y_);                           \
        i[a] = p_ ? ib_ : ia_;                          \
        i[b] = p_ ? ia_ : ib_;                          \
    }
// 60 compare-exchanges, 10 layers (the smallest known 16-input network; 0-1 principle checked exhaustively)
__device__ __forceinline__ void sort16(float (&d)[16], unsigned (&i)[16]) {
    SSDR_CE(0, 13) SSDR_CE(1, 12) SSDR_CE(2, 15) SSDR_CE(3, 14) SSDR_CE(4, 8) SSDR_CE(5, 6) SSDR_CE(7, 11) SSDR_CE(9, 10)
    SSDR_CE(0, 5) SSDR_CE(1, 7) SSDR_CE(2, 9) SSDR_CE(3, 4) SSDR_CE(6, 13) SSDR_CE(8, 14) SSDR_CE(10, 15) SSDR_CE(11, 12)
    SSDR_CE(0, 1) SSDR_CE(2, 3) SSDR_CE(4, 5) SSDR_CE(6, 8) SSDR_CE(7, 9) SSDR_CE(10, 11) SSDR_CE(12, 13) SSDR_CE(14, 15)
    SSDR_CE(0, 2) SSDR_CE(1, 3) SSDR_CE(4, 10) SSDR_CE(5, 11) SSDR_CE(6, 7) SSDR_CE(8, 9) SSDR_CE(12, 14) SSDR_CE(13, 15)
    SSDR_CE(1, 2) SSDR_CE(3, 12) SSDR_CE(4, 6) SSDR_CE(5, 7) SSDR_CE(8, 10) SSDR_CE(9, 11) SSDR_CE(13, 14)
    SSDR_CE(1, 4) SSDR_CE(2, 6) SSDR_CE(5, 8) SSDR_CE(7, 10) SSDR_CE(9, 13) SSDR_CE(11, 14)
    SSDR_CE(2, 4) SSDR_CE(3, 6) SSDR_CE(9, 12) SSDR_CE(11, 13)
    SSDR_CE(3, 5) SSDR_CE(6, 8) SSDR_CE(7, 9) SSDR_CE(10, 12)
    SSDR_CE(3, 4) SSDR_CE(5, 6) SSDR_CE(7, 8) SSDR_CE(9, 10) SSDR_CE(11, 12)
    SSDR_CE(6, 7) SSDR_CE(8, 9)
}
// ascending order of a bitonic 16-sequence: 4 layers of 8
__device__ __forceinline__ void bitonic_merge16(float (&d)[16], unsigned (&i)[16]) {
#pragma unroll
    for (int st = 8; st > 0; st >>= 1)
#pragma unroll
        for (int a = 0; a < 16; ++a)
            if ((a & st) == 0) SSDR_CE(a, a + st)
}
#undef SSDR_CE

// Conservative distance from a coordinate to the slab [u0, u0 + h) of a cell row, all measured from the lower face of
// the coordinate's own cell: max(u0 - f0, f0 - u0 - h, 0) - eps, never negative.  na = -f0 - eps and nb = f0 - h - eps
// are per query and axis, so a row costs two adds and one three-input max.
__device__ __forceinline__ float slab_gap_rel(float u0, float na, float nb) {
    return fmaxf(fmaxf(u0 + na, nb - u0), 0.f);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// p = *src where `on`, else p keeps whatever it held (an inactive slot is never looked at, so it needs no value)
__device__ __forceinline__ void ldg_if(float4& p, const float4* src, bool on) {
    asm volatile(
        "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %4, 0;\n\t@pp ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%5];\n\t}"
        : "+f"(p.x), "+f"(p.y), "+f"(p.z), "+f"(p.w)
        : "r"((int)on), "l"(src));
}

template <typename OutT>
__global__ void __launch_bounds__(Q16_THREADS, 5) query16_kernel(const float4* __restrict__ spts,
                                                                 const float4* __restrict__ sq,
                                                                 const unsigned* __restrict__ cell_start,
                                                                 const ItemMeta* __restrict__ items, unsigned N, unsigned Q,
                                                                 unsigned q_begin, unsigned q_end, int K,
                                                                 OutT* __restrict__ out, unsigned* __restrict__ flag_count,
                                                                 unsigned* __restrict__ flag_list,
                                                                 unsigned long long* __restrict__ evals) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint2 s_queue[Q16_CAP][Q16_THREADS];  // (distance bits, index); a lane only ever touches its own column
    __shared__ uint2 s_list[Q16_L][Q16_THREADS];     // the lane's sorted list, parked here between merges
    __shared__ uint2 s_desc[Q16_DESC][Q16_THREADS];  // (first position, length) of the lane's rows of the current chunk
    const unsigned tid = threadIdx.x;
    const unsigned t = q_begin + blockIdx.x * blockDim.x + tid;
    const bool valid = t < q_end;
    const unsigned tq = valid ? t : q_end - 1;
    unsigned my_evals = 0;
    const unsigned b = tq / Q;
    const ItemMeta m = items[b];
    const float4 q = __ldg(sq + tq);
    const unsigned qid = (unsigned)__float_as_int(q.w);
    const int gx = m.g[0], gy = m.g[1], gz = m.g[2];
    const int cx = cell_coord(q.x, m.lo[0], m.inv_h, gx);
    const int cy = cell_coord(q.y, m.lo[1], m.inv_h, gy);
    const int cz = cell_coord(q.z, m.lo[2], m.inv_h, gz);
    const float h = m.h;
    // offsets of the query above the lower faces of its own cell, folded with the slack (see slab_gap_rel)
    const float fx0 = q.x - (m.lo[0] + (float)cx * h), fy0 = q.y - (m.lo[1] + (float)cy * h),
                fz0 = q.z - (m.lo[2] + (float)cz * h);
    const float nax = -fx0 - m.eps_abs, nbx = fx0 - h - m.eps_abs;
    const float nay = -fy0 - m.eps_abs, nby = fy0 - h - m.eps_abs;
    const float naz = -fz0 - m.eps_abs, nbz = fz0 - h - m.eps_abs;
    const unsigned* cs = cell_start + m.cell_base;
#pragma unroll
    for (int j = 0; j < Q16_L; ++j) s_list[j][tid] = make_uint2(0x7F800000u, 0xFFFFFFFFu);
    float m17 = INFINITY;
    unsigned cnt = 0;
    bool done = !valid;
    float4 p[Q16_STEP];
#pragma unroll
    for (int u = 0; u < Q16_STEP; ++u) p[u] = q;

    int r = 1, nseg = 9;            // shell radius and its row segments (warp-uniform)
    int base = 0;                   // segments of this shell already scanned
    int it_dz = 0, it_dy = 0, it_lim = 0, it_ph = 0;  // row iterator of shells r >= 2: phase 0 = every row of the
                                                      // (2r+1)^2 block (whole on the shell's faces, else its left end),
                                                      // phase 1 = right ends of the interior rows
    for (;;) {
        // ---- descriptors of segments [base, base + nk) of shell r.  The two cell-start loads of a segment are used one
        // iteration later, so a load is in flight while the next segment's geometry is worked out.
        const int nk = min(Q16_DESC, nseg - base);
        unsigned pa = 0, pb = 0;  // loads of the previous segment (0, 0 = empty)
        for (int j = 0; j <= nk; ++j) {
            unsigned na_ = 0, nb_ = 0;
            if (j < nk) {
                int dz, dy, dxa, dxb;  // row offsets; cell offsets of the segment's ends
                if (r == 1) {
                    dz = (int)((0x2402049u >> (3 * (base + j))) & 7u) - 1;  // nearest rows first, see query_kernel
                    dy = (int)((0x2081281u >> (3 * (base + j))) & 7u) - 1;
                    dxa = -1;
                    dxb = 1;
                } else {
                    dz = it_dz;
                    dy = it_dy;
                    const bool whole = it_ph == 0 && (dz == -r || dz == r || dy == -r || dy == r);
                    dxa = it_ph ? r : -r;
                    dxb = whole ? r : dxa;
                    if (++it_dy > it_lim) {
                        it_dy = -it_lim;
                        if (++it_dz > it_lim) {
                            it_ph = 1;
                            it_lim = r - 1;
                            it_dz = it_dy = -it_lim;
                        }
                    }
                }
                const int z = cz + dz, y = cy + dy;
                const int xa = max(cx + dxa, 0), xb = min(cx + dxb, gx - 1);
                if (!done && (unsigned)z < (unsigned)gz && (unsigned)y < (unsigned)gy && xa <= xb) {
                    const float gy_ = slab_gap_rel((float)dy * h, nay, nby);
                    const float gz_ = slab_gap_rel((float)dz * h, naz, nbz);
                    // a segment that spans the query's own cell column has no gap along x
                    const float gx_ = dxa != dxb ? 0.f : slab_gap_rel((float)dxa * h, nax, nbx);
                    const float lower = (gx_ * gx_ + gy_ * gy_ + gz_ * gz_) * 0.99999f;
                    if (lower < m17) {
                        const unsigned row = (unsigned)((z * gy + y) * gx);
                        na_ = __ldg(cs + row + (unsigned)xa);
                        nb_ = __ldg(cs + row + (unsigned)xb + 1u);
                    }
                }
            }
            if (j > 0) {
                const unsigned len = pb - pa;
                my_evals += len;
                s_desc[j - 1][tid] = make_uint2(pa, len);
                for (unsigned e = 0; e < len; e += 8) prefetch_l1(spts + pa + e);  // 8 records = one 128-byte line
                if (len) prefetch_l1(spts + pa + len - 1);
            }
            pa = na_;
            pb = nb_;
        }
        // ---- flat scan of the chunk; a merge runs whenever some lane has a full batch, and at the end of the shell
        const bool last_chunk = base + nk >= nseg;
        int k = 0;
        unsigned rem = 0, ptr = 0;
        for (;;) {
            bool exhausted;
            for (;;) {
                while (rem == 0 && k < nk) {
                    const uint2 dsc = s_desc[k][tid];
                    ++k;
                    ptr = dsc.x;
                    rem = dsc.y;
                }
                exhausted = !__any_sync(FULL, rem != 0);
                if (exhausted) break;
                const unsigned n = rem < (unsigned)Q16_STEP ? rem : (unsigned)Q16_STEP;
#pragma unroll
                for (int u = 0; u < Q16_STEP; ++u) ldg_if(p[u], spts + ptr + u, (unsigned)u < n);
#pragma unroll
                for (int u = 0; u < Q16_STEP; ++u) {
                    const float dx = __fsub_rn(q.x, p[u].x), dy_ = __fsub_rn(q.y, p[u].y), dz_ = __fsub_rn(q.z, p[u].z);
                    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy_, dy_)), __fmul_rn(dz_, dz_));
                    if ((unsigned)u < n && d < m17) {
                        s_queue[cnt][tid] = make_uint2(__float_as_uint(d), __float_as_uint(p[u].w));
                        ++cnt;
                    }
                }
                ptr += n;
                rem -= n;
                if (__any_sync(FULL, cnt >= (unsigned)Q16_L)) break;
            }
            const bool tail = exhausted && last_chunk && __any_sync(FULL, cnt != 0);
            if (!exhausted || tail) {
                // ---- merge: sort the batch, keep the 16 smallest of list + batch, fold the rest into m17
                float d[Q16_L];
                unsigned i[Q16_L];
#pragma unroll
                for (int j = 0; j < Q16_L; ++j) {
                    const uint2 e = s_queue[j][tid];
                    d[j] = (unsigned)j < cnt ? __uint_as_float(e.x) : INFINITY;
                    i[j] = e.y;
                }
#pragma unroll
                for (int u = 0; u < Q16_STEP - 1; ++u)  // entries beyond the batch move to the front of the queue
                    if ((unsigned)(Q16_L + u) < cnt) s_queue[u][tid] = s_queue[Q16_L + u][tid];
                cnt = cnt > (unsigned)Q16_L ? cnt - (unsigned)Q16_L : 0u;
                sort16(d, i);
                // list ascending against batch descending: the minima are the 16 smallest (a bitonic sequence, here in
                // the batch's registers, i.e. reversed -- still bitonic), the maxima only matter through their minimum
                float hmin = INFINITY;
#pragma unroll
                for (int j = 0; j < Q16_L; ++j) {
                    const uint2 e = s_list[j][tid];
                    const float a = __uint_as_float(e.x), c2 = d[Q16_L - 1 - j];
                    const bool pa_ = a < c2;
                    d[Q16_L - 1 - j] = fminf(a, c2);
                    hmin = fminf(hmin, fmaxf(a, c2));
                    i[Q16_L - 1 - j] = pa_ ? e.y : i[Q16_L - 1 - j];
                }
                m17 = fminf(m17, hmin);
                bitonic_merge16(d, i);
#pragma unroll
                for (int j = 0; j < Q16_L; ++j) s_list[j][tid] = make_uint2(__float_as_uint(d[j]), i[j]);
            }
            if (exhausted && !(tail && __any_sync(FULL, cnt != 0))) break;  // (left-overs of a full batch: once more)
        }
        if (!last_chunk) {  // next chunk of the same shell
            base += nk;
            continue;
        }
        // ---- the shell is complete: test termination per lane (the queues are empty)
        if (!done) {
            const int x0 = max(cx - r, 0), x1 = min(cx + r, gx - 1);
            const int y0 = max(cy - r, 0), y1 = min(cy + r, gy - 1);
            const int z0 = max(cz - r, 0), z1 = min(cz + r, gz - 1);
            if (x0 == 0 && x1 == gx - 1 && y0 == 0 && y1 == gy - 1 && z0 == 0 && z1 == gz - 1) {
                done = true;
            } else {
                float R = INFINITY;
                if (cx - r >= 0) R = fminf(R, q.x - (m.lo[0] + (float)(cx - r) * m.h));
                if (cx + r <= gx - 1) R = fminf(R, (m.lo[0] + (float)(cx + r + 1) * m.h) - q.x);
                if (cy - r >= 0) R = fminf(R, q.y - (m.lo[1] + (float)(cy - r) * m.h));
                if (cy + r <= gy - 1) R = fminf(R, (m.lo[1] + (float)(cy + r + 1) * m.h) - q.y);
                if (cz - r >= 0) R = fminf(R, q.z - (m.lo[2] + (float)(cz - r) * m.h));
                if (cz + r <= gz - 1) R = fminf(R, (m.lo[2] + (float)(cz + r + 1) * m.h) - q.z);
                R -= m.eps_abs;
                // the 17th smallest distance must be strictly inside the guaranteed radius
                if (R > 0.f && m17 < R * R * 0.99999f) done = true;
            }
        }
        if (!__any_sync(FULL, !done)) break;
        ++r;
        const int w = 2 * r + 1;
        nseg = w * w + (w - 2) * (w - 2);
        base = 0;
        it_dz = it_dy = -r;
        it_lim = r;
        it_ph = 0;
    }

    // ---- emit + tie flags (dist[16] of the 17-slot kernel is m17 here)
    if (valid) {
        const int nvalid = (unsigned)K < N ? K : (int)N;
        OutT* o = out + ((size_t)b * Q + qid) * (size_t)K;
        float Ld[Q16_L];
        unsigned Li[Q16_L];
#pragma unroll
        for (int j = 0; j < Q16_L; ++j) {
            const uint2 e = s_list[j][tid];
            Ld[j] = __uint_as_float(e.x);
            Li[j] = e.y;
        }
        bool flag = false;
        if (K == Q16_L && nvalid == Q16_L && (reinterpret_cast<size_t>(o) & 15) == 0) {  // full, aligned rows: 16-byte stores
            if (sizeof(OutT) == 8) {
#pragma unroll
                for (int j = 0; j < Q16_L; j += 2)
                    *reinterpret_cast<longlong2*>(o + j) = make_longlong2((long long)Li[j], (long long)Li[j + 1]);
            } else {
#pragma unroll
                for (int j = 0; j < Q16_L; j += 4)
                    *reinterpret_cast<int4*>(o + j) = make_int4((int)Li[j], (int)Li[j + 1], (int)Li[j + 2], (int)Li[j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < Q16_L; ++j)
                if (j < nvalid) o[j] = (OutT)Li[j];
        }
#pragma unroll
        for (int j = 0; j < Q16_L; ++j) {
            const float nxt = j + 1 < Q16_L ? Ld[(j + 1) & (Q16_L - 1)] : m17;
            if (j + 1 < nvalid && Ld[j] == nxt) flag = true;
            if (j + 1 == K && N > (unsigned)K && nxt <= Ld[j] * 1.00001f) flag = true;
        }
        if (flag) flag_list[atomicAdd(flag_count, 1u)] = b * Q + qid;
    }
#pragma unroll
    for (int mm = 16; mm > 0; mm >>= 1) my_evals += __shfl_xor_sync(FULL, my_evals, mm);
    if ((tid & 31) == 0 && my_evals) atomicAdd(evals, (unsigned long long)my_evals);
}

struct DevStats {
    unsigned flag_count;
    unsigned pad;
    unsigned long long evals;
};

static float g_occupancy_scale = 0.3f;  // points per cell ~= scale * K (tunable: SSDR_KNN_OCCUPANCY)
static bool g_probe_on = true;
static bool g_query_v2 = true;  // SSDR_KNN_QUERY=1 selects the insertion kernel for 8 < K <= 16 (A/B runs)

// Rows the tie path rewrote, compacted for a small second read-back (the bulk read-back of all rows runs on the copy
// stream while the tie path works; the host then overwrites these rows).
template <typename OutT>
__global__ void gather_rows_kernel(const unsigned* __restrict__ flag_list, const unsigned* __restrict__ n_flag,
                                   const OutT* __restrict__ out, int K, unsigned cap, OutT* __restrict__ patch) {
    unsigned n = *n_flag;
    n = n < cap ? n : cap;
    const unsigned long long total = (unsigned long long)n * K;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned r = (unsigned)(i / K), j = (unsigned)(i % K);
        patch[i] = out[(size_t)flag_list[r] * K + j];
    }
}

// ---- B' (tiny clouds): brute force, one warp per query -------------------------------------------------------------
// Up to 1024 support points per item (the lowest pyramid levels): no cell grid at all.  The CTA keeps the item's points
// in shared memory, every lane of a warp holds the distances to SLOTS of them in registers (same fp32 expression as
// everywhere), and min(K+1, N) rounds of a warp-wide arg-min on (distance, index) peel off the neighbours in order.
// Rows with equal distances inside the top K, or a (near) tie between rank K and K+1, are flagged for the tie path
// exactly like in query_kernel, so the contract is the same: unflagged rows are final.
constexpr int TINY_MAX_POINTS = 1024;
constexpr int TINY_WARPS = 16;
constexpr int TINY_Q_PER_WARP = 1;

template <int SLOTS, typename OutT>
__global__ void __launch_bounds__(TINY_WARPS * 32) tiny_query_kernel(const float* __restrict__ pts_all,
                                                                    const float* __restrict__ q_all, unsigned N,
                                                                    unsigned Q, int K, OutT* __restrict__ out,
                                                                    unsigned* __restrict__ flag_count,
                                                                    unsigned* __restrict__ flag_list,
                                                                    unsigned long long* __restrict__ evals) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ float sx[TINY_MAX_POINTS], sy[TINY_MAX_POINTS], sz[TINY_MAX_POINTS];
    const unsigned qblocks = (Q + TINY_WARPS * TINY_Q_PER_WARP - 1) / (TINY_WARPS * TINY_Q_PER_WARP);
    const unsigned b = blockIdx.x / qblocks, qblk = blockIdx.x % qblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* p = pts_all + (size_t)b * N * 3;
    for (unsigned i = threadIdx.x; i < N; i += blockDim.x) {
        sx[i] = __ldg(p + 3 * (size_t)i);
        sy[i] = __ldg(p + 3 * (size_t)i + 1);
        sz[i] = __ldg(p + 3 * (size_t)i + 2);
    }
    __syncthreads();
    const int nvalid = (unsigned)K < N ? K : (int)N;
    const int rounds = N > (unsigned)K ? K + 1 : (int)N;
    const unsigned q0 = (qblk * TINY_WARPS + warp) * TINY_Q_PER_WARP;
    for (unsigned qi = q0; qi < q0 + TINY_Q_PER_WARP && qi < Q; ++qi) {
        const float* qp = q_all + ((size_t)b * Q + qi) * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        float d[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const unsigned i = (unsigned)s * 32u + (unsigned)lane;  // slot order == index order inside a lane
            d[s] = INFINITY;
            if (i < N) {
                const float dx = __fsub_rn(qx, sx[i]), dy = __fsub_rn(qy, sy[i]), dz = __fsub_rn(qz, sz[i]);
                d[s] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            }
        }
        OutT* o = out + ((size_t)b * Q + qi) * (size_t)K;
        bool flag = false;
        float prev = 0.f;
        for (int j = 0; j < rounds; ++j) {
            // lane-local minimum, lowest index first (strict '<' over ascending slots)
            float bd = d[0];
            int bs = 0;
#pragma unroll
            for (int s = 1; s < SLOTS; ++s)
                if (d[s] < bd) {
                    bd = d[s];
                    bs = s;
                }
            unsigned bi = (unsigned)bs * 32u + (unsigned)lane;
            // warp arg-min on (distance, index)
            float wd = bd;
            unsigned wi = bi;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                const float od = __shfl_xor_sync(FULL, wd, m);
                const unsigned oi = __shfl_xor_sync(FULL, wi, m);
                if (od < wd || (od == wd && oi < wi)) {
                    wd = od;
                    wi = oi;
                }
            }
            if (wi == bi) {  // the owner retires the slot
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
                    if (s == bs) d[s] = INFINITY;
            }
            if (j < nvalid) {
                if (lane == 0) o[j] = (OutT)wi;
                if (j > 0 && wd == prev) flag = true;
            } else if (wd <= prev * 1.00001f) {  // j == K: (near) tie between rank K and K+1
                flag = true;
            }
            prev = wd;
        }
        if (flag && lane == 0) flag_list[atomicAdd(flag_count, 1u)] = b * Q + qi;
    }
    if (threadIdx.x == 0) {
        const unsigned first = qblk * TINY_WARPS * TINY_Q_PER_WARP;
        const unsigned nq = first < Q ? min(Q - first, (unsigned)(TINY_WARPS * TINY_Q_PER_WARP)) : 0u;
        if (nq) atomicAdd(evals, (unsigned long long)nq * N);
    }
}

constexpr unsigned MAX_CHUNKS = 2;
constexpr unsigned PATCH_CAP = 1u << 15;  // rows; more flagged rows than this fall back to a second full read-back

// h_out (nullable, host entry points only, K <= N): the results are also delivered to this host buffer; the bulk
// device->host copy starts right behind the main kernel and overlaps the tie path.
template <typename OutT>
static int run_dev(Ctx* c, cudaStream_t s, const float* d_pts, size_t B, size_t N, const float* d_q, size_t Q, size_t K,
                   OutT* d_out, ssdr_knn_stats* stats, OutT* h_out = nullptr, AsyncCtx* ac = nullptr) {
    SSDR_REQUIRE(d_pts && d_q && d_out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_INVALID, "npts must be >= 1 (the reference asserts npts != 0)");
    SSDR_REQUIRE(K >= 1, SSDR_ERR_INVALID, "K must be >= 1");
    SSDR_REQUIRE(K <= 64, SSDR_ERR_UNSUPPORTED, "K=%zu > 64 is not supported", K);
    SSDR_REQUIRE(B * N < 0x7FFFFFFFull && B * Q < 0x7FFFFFFFull, SSDR_ERR_UNSUPPORTED, "batch too large");
    if (B == 0 || Q == 0) return SSDR_OK;
    static std::once_flag env_once;  // the library is re-entrant per calling thread: read the knobs exactly once
    std::call_once(env_once, [] {
        const char* e = getenv("SSDR_KNN_OCCUPANCY");
        if (e && atof(e) > 0) g_occupancy_scale = (float)atof(e);
        const char* pr = getenv("SSDR_KNN_PROBE");  // SSDR_KNN_PROBE=0 switches the occupancy probe off (A/B runs)
        g_probe_on = !(pr && pr[0] == '0');
        const char* qk = getenv("SSDR_KNN_QUERY");  // SSDR_KNN_QUERY=1: per-candidate insertion kernel for 8 < K <= 16
        g_query_v2 = !(qk && qk[0] == '1');
    });
    float occupancy = g_occupancy_scale * (float)K;
    occupancy = occupancy < 0.6f ? 0.6f : (occupancy > 24.f ? 24.f : occupancy);
    // cells per item: for a cloud that fills its bbox h gives cells ~= N / occupancy (3x head-room for the +1 per axis);
    // a cloud of surfaces needs the same number of OCCUPIED cells inside a mostly empty bbox, hence up to 4 cells per
    // point when the occupancy probe runs (setup_item grows h until the grid fits, so the bound costs speed at worst,
    // never correctness)
    const bool probe_on = g_probe_on;
    // large clouds are probed on a strided sample (the cube counts only need occupancies well above one)
    const unsigned pstride = (unsigned)((N + PROBE_SAMPLE - 1) / PROBE_SAMPLE);
    const unsigned r2 = probe_on ? probe_resolution((unsigned)((N + pstride - 1) / pstride)) : 0u;
    size_t cap = (size_t)(3.0 * (double)N / occupancy) + 64;
    {
        // cells a probed cloud may use per point: 4, or 8 for one large cloud queried with its own points (the query
        // job shares the point job's arrays then, so the larger grid costs one scan, not two).  Config 3's barycentres
        // (profiles/r02_cfg3_knn_cells.txt): 7.6 ms at 4, 6.5 at 8, 6.4 at 16, 6.9 at 32 cells per point on one GPU;
        // as one of eight query blocks 2.4 / 2.8 / 4.0 / 4.9 ms.  SSDR_KNN_CELLS_PER_POINT overrides (A/B runs).
        static const size_t cpp_env = [] {
            const char* e = getenv("SSDR_KNN_CELLS_PER_POINT");
            const int v = e ? atoi(e) : 0;
            return (size_t)(v < 0 ? 0 : (v > 64 ? 64 : v));
        }();
        const size_t cpp = cpp_env ? cpp_env : ((d_q == d_pts && Q == N && N >= ((size_t)1 << 20)) ? 8 : 4);
        if (r2 && cap < cpp * N) cap = cpp * N;
    }
    // (a scan of surfaces inside a 200 m bbox wants far more cells than points: the old 2^24 bound put 90 barycentres
    // into every occupied cell of config 3's cloud -- 772 distance evaluations per query instead of ~150)
    if (cap > ((size_t)1 << 28)) cap = (size_t)1 << 28;
    if (cap > ((size_t)0xFFFFFFF0u / B) - 2) cap = ((size_t)0xFFFFFFF0u / B) - 2;  // 32-bit cell indices over all items
    if (N <= SG_MAX_POINTS && cap > SG_MAX_CELLS - 1 && (size_t)(3.0 * (double)N / occupancy) + 64 <= SG_MAX_CELLS - 1)
        cap = SG_MAX_CELLS - 1;  // stay within the shared memory of the single-CTA build
    const unsigned cstride = (unsigned)cap + 1;
    const size_t ncell = B * (size_t)cstride + 1;
    const bool self = (d_q == d_pts && Q == N);
    const unsigned totalP = (unsigned)(B * N), totalQ = (unsigned)(B * Q);
    const int njobs = self ? 1 : 2;

    // one zero-initialised control slab: stats | ticket | bbox encodings | scan scratch | probe | counts (P,Q)
    const size_t nscan_words = prim::scan_scratch_words((size_t)njobs * ncell);
    const size_t probe_words = r2 ? B * (size_t)(PROBE_WORDS + 2) + 2 : 0;  // multi-launch path: bitmaps, counts, ticket
    const size_t hdr_words = 16 + B * 6 + nscan_words + probe_words;  // stats, ticket, bbox, scan scratch, probe (zeroed)
    const size_t ctl_words = hdr_words + (size_t)njobs * ncell;
    SSDR_TRY(c->ws[WS_CNT_P].reserve(ctl_words * 4));
    SSDR_TRY(c->ws[WS_ITEMS].reserve(B * sizeof(ItemMeta)));
    SSDR_TRY(c->ws[WS_CELL_P].reserve(((size_t)totalP + (self ? 0 : totalQ)) * 4));
    SSDR_TRY(c->ws[WS_START_P].reserve((size_t)njobs * ncell * 4));
    SSDR_TRY(c->ws[WS_SORT_P].reserve((size_t)totalP * sizeof(float4)));
    SSDR_TRY(c->ws[WS_FLAGS].reserve((size_t)totalQ * 4 + 16));
    if (!self) SSDR_TRY(c->ws[WS_SORT_Q].reserve((size_t)totalQ * sizeof(float4)));
    unsigned* ctl = c->ws[WS_CNT_P].as<unsigned>();
    DevStats* dstats = reinterpret_cast<DevStats*>(ctl);
    unsigned* ticket = ctl + 8;
    unsigned* enc = ctl + 16;
    unsigned* counts = ctl + hdr_words;                       // [njobs][ncell]
    unsigned* starts = c->ws[WS_START_P].as<unsigned>();      // [njobs][ncell], one concatenated exclusive scan
    ItemMeta* items = c->ws[WS_ITEMS].as<ItemMeta>();
    unsigned* start_p = starts;
    float4* sort_p = c->ws[WS_SORT_P].as<float4>();
    unsigned* flag_list = c->ws[WS_FLAGS].as<unsigned>();

    unsigned long long n_launch = 0;
    if (stats) SSDR_CHECK_CUDA(cudaEventRecord(c->tev[0], s));
    // tiny clouds: brute force beats any index (and needs none) while the evaluations stay in the tens of millions
    // (its cost is ~220 warp instructions per arg-min round and query; beyond ~160k rounds the indexed kernel's better
    // instruction economy wins over its latency)
    const bool tiny = N <= (size_t)TINY_MAX_POINTS && (size_t)totalQ * ((K < N ? K + 1 : N)) <= (size_t)160 * 1024;
    const bool small = !tiny && cstride <= SG_MAX_CELLS && N <= SG_MAX_POINTS && Q <= 8 * SG_MAX_POINTS;
    float4* sort_q_buf = self ? nullptr : c->ws[WS_SORT_Q].as<float4>();
    if (tiny) {
        SSDR_CHECK_CUDA(cudaMemsetAsync(ctl, 0, 16 * 4, s));  // counters only: no cell grid
    } else if (small) {  // one CTA per item does the whole grid build
        SSDR_CHECK_CUDA(cudaMemsetAsync(ctl, 0, 16 * 4, s));
        static std::atomic<bool> attr_set_dev[64];  // per device; setting the attribute twice is harmless
        std::atomic<bool>& attr_set = attr_set_dev[c->device & 63];
        if (!attr_set.load(std::memory_order_acquire)) {
            SSDR_CHECK_CUDA(cudaFuncSetAttribute(small_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(SG_MAX_CELLS * sizeof(unsigned))));
            attr_set.store(true, std::memory_order_release);
        }
        const size_t sg_words = (r2 && cstride < PROBE_WORDS) ? (size_t)PROBE_WORDS : (size_t)cstride;
        small_grid_kernel<<<(unsigned)B, SG_THREADS, sg_words * sizeof(unsigned), s>>>(
            d_pts, (unsigned)N, self ? nullptr : d_q, (unsigned)Q, enc, items, (int)B, occupancy, (unsigned)cap, cstride,
            starts, sort_p, sort_q_buf, r2);
        n_launch += 1;
    } else {
        SSDR_REQUIRE(B <= 65535, SSDR_ERR_UNSUPPORTED, "more than 65535 batch items of more than %u points each",
                     SG_MAX_POINTS);
        SSDR_CHECK_CUDA(cudaMemsetAsync(ctl, 0, ctl_words * 4, s));
        unsigned bx = (unsigned)((N + 1023) / 1024);
        if (bx > 1024) bx = 1024;
        bbox_setup_kernel<<<dim3(bx, (unsigned)B), 256, 0, s>>>(d_pts, (unsigned)N, enc, ticket, items, (int)B, occupancy,
                                                               (unsigned)cap, cstride, r2 != 0);
        if (r2) {
            unsigned* pbase = ctl + 16 + B * 6 + nscan_words;  // [ticket, pad][B][2 counts][B][PROBE_WORDS]
            probe_kernel<<<dim3(bx, (unsigned)B), 256, 0, s>>>(d_pts, (unsigned)N, enc, pbase + 2 + 2 * B, pbase + 2, pbase,
                                                              items, (int)B, occupancy, (unsigned)cap, cstride, r2, pstride);
            n_launch += 1;
        }
        SortJob jp, jq;
        jp.xyz = d_pts;
        jp.n = (unsigned)N;
        jp.total = totalP;
        jp.cell_of = c->ws[WS_CELL_P].as<unsigned>();
        jp.counts = counts;
        jp.starts = starts;
        jp.sorted = sort_p;
        jp.start_bias = 0;
        jq = jp;
        jq.total = 0;
        if (!self) {
            jq.xyz = d_q;
            jq.n = (unsigned)Q;
            jq.total = totalQ;
            jq.cell_of = jp.cell_of + totalP;
            jq.counts = counts + ncell;
            jq.starts = starts + ncell;
            jq.sorted = c->ws[WS_SORT_Q].as<float4>();
            jq.start_bias = totalP;  // the concatenated scan continues behind the points' total
        }
        const unsigned tot = totalP + jq.total;
        cell_count_kernel<<<(tot + 255) / 256, 256, 0, s>>>(jp, jq, items);
        SSDR_TRY(prim::exclusive_scan_u32(counts, starts, (size_t)njobs * ncell, ctl + 16 + B * 6, nullptr, s));
        cell_scatter_kernel<<<(tot + 255) / 256, 256, 0, s>>>(jp, jq);
        n_launch += 5;  // bbox_setup, cell_count, scan_sums, scan_apply, cell_scatter
    }
    const float4* sort_q = self ? sort_p : sort_q_buf;
    SSDR_CHECK_CUDA(cudaGetLastError());
    if (K > N) SSDR_CHECK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)totalQ * K * sizeof(OutT), s));

    if (stats) SSDR_CHECK_CUDA(cudaEventRecord(c->tev[1], s));
    // Host entry points with a large result: the items are queried in a few launches and every finished chunk of rows
    // starts its device->host copy at once (copy stream), so PCIe runs under the remaining launches.
    const size_t out_bytes = (size_t)totalQ * K * sizeof(OutT);
    struct CopyFence {  // no return path, error or not, may leave a copy into the caller's buffer in flight
        cudaStream_t cs;
        bool armed;
        ~CopyFence() {
            if (armed) cudaStreamSynchronize(cs);
        }
    } fence{c->copy_stream, false};
    // (two chunks only, and only for big results: a launch over fewer queries is hardly shorter -- its duration is set
    // by the slowest warps -- so more chunks cost more kernel time than the copy overlap returns; measured)
    const unsigned nchunk = (!tiny && h_out && B >= 2 && out_bytes >= ((size_t)16 << 20)) ? MAX_CHUNKS : 1u;
    if (tiny) {
        const unsigned per_cta = TINY_WARPS * TINY_Q_PER_WARP;
        const unsigned tblocks = (unsigned)B * (unsigned)((Q + per_cta - 1) / per_cta);
#define SSDR_TINY(SL)                                                                                              \
    tiny_query_kernel<SL, OutT><<<tblocks, TINY_WARPS * 32, 0, s>>>(d_pts, d_q, (unsigned)N, (unsigned)Q, (int)K, d_out, \
                                                                    &dstats->flag_count, flag_list, &dstats->evals)
        if (N <= 128) SSDR_TINY(4);
        else if (N <= 256) SSDR_TINY(8);
        else if (N <= 512) SSDR_TINY(16);
        else SSDR_TINY(32);
#undef SSDR_TINY
        SSDR_CHECK_CUDA(cudaGetLastError());
        n_launch += 1;
    }
    for (unsigned ch = 0; !tiny && ch < nchunk; ++ch) {
        const unsigned b0 = (unsigned)(B * ch / nchunk), b1 = (unsigned)(B * (ch + 1) / nchunk);
        const unsigned qb = b0 * (unsigned)Q, qe = b1 * (unsigned)Q;
        const unsigned qblocks = (qe - qb + 127) / 128;
#define SSDR_QUERY(KC)                                                                                          \
    query_kernel<KC, OutT><<<qblocks, 128, 0, s>>>(sort_p, sort_q, start_p, items, (unsigned)N, (unsigned)Q, qb, qe, \
                                                   (int)K, d_out, &dstats->flag_count, flag_list, &dstats->evals)
        if (K == 1) SSDR_QUERY(2);
        else if (K <= 4) SSDR_QUERY(5);
        else if (K <= 8) SSDR_QUERY(9);
        else if (K <= 16 && g_query_v2)
            query16_kernel<OutT><<<qblocks, Q16_THREADS, 0, s>>>(sort_p, sort_q, start_p, items, (unsigned)N, (unsigned)Q, qb,
                                                                 qe, (int)K, d_out, &dstats->flag_count, flag_list,
                                                                 &dstats->evals);
        else if (K <= 16) SSDR_QUERY(17);
        else if (K <= 32) SSDR_QUERY(33);
        else SSDR_QUERY(65);
#undef SSDR_QUERY
        SSDR_CHECK_CUDA(cudaGetLastError());
        n_launch += 1;
        if (nchunk > 1) {
            fence.armed = true;
            SSDR_CHECK_CUDA(cudaEventRecord(c->ev_chunk[ch], s));
            SSDR_CHECK_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_chunk[ch], 0));
            SSDR_CHECK_CUDA(cudaMemcpyAsync(h_out + (size_t)qb * K, d_out + (size_t)qb * K,
                                            (size_t)(qe - qb) * K * sizeof(OutT), cudaMemcpyDeviceToHost, c->copy_stream));
        }
    }
    if (stats) SSDR_CHECK_CUDA(cudaEventRecord(c->tev[2], s));

    // ---- C: exact nanoflann replay of the flagged rows.  Boundary near-ties alone flag ~1.5e-5*K of the rows, so a
    // call with K * queries >= 131072 has flagged rows almost surely and the tie path is enqueued right behind the
    // main kernel without a host round trip (its kernels return at once if the device-side count is zero).
    // Otherwise ties are rare: read the count first and skip the (cooperative, whole-GPU) launch when it is zero.
    if (ac) {  // asynchronous flavour: the tie path is enqueued unconditionally, nothing is read back
        SSDR_TRY((kdtree::enqueue_tie_path_async<OutT>(c, s, d_pts, B, N, d_q, Q, K, d_out, flag_list, &dstats->flag_count,
                                                       &ac->tree, ac->reuse, ac->status, &n_launch)));
        g_last_launches += n_launch;
        return SSDR_OK;
    }
    kdtree::Tree tree;
    tree.error = nullptr;
    DevStats hs;
    unsigned h_err = 0;
    const bool speculate = (size_t)totalQ * K >= 131072;  // expected flagged rows ~ 1.5e-5 * K * queries >= 2
    const bool bulk_copy = h_out && nchunk == 1;  // otherwise the rows are already on their way, chunk by chunk
    // small results are not worth a second stream (event + wait + extra synchronisation cost more than they hide)
    const bool side_copy = bulk_copy && out_bytes >= ((size_t)1 << 20);
    cudaStream_t cs = side_copy ? c->copy_stream : s;
    OutT* d_patch = nullptr;
    if (h_out) {
        SSDR_TRY(c->ws[WS_PATCH].reserve((size_t)PATCH_CAP * K * sizeof(OutT)));
        d_patch = c->ws[WS_PATCH].as<OutT>();
        if (side_copy) {
            fence.armed = true;
            SSDR_CHECK_CUDA(cudaEventRecord(c->ev_main, s));
            SSDR_CHECK_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
        }
    }
    bool tree_reused = false;
    auto tie_path = [&]() -> int {
        SSDR_TRY((kdtree::enqueue_tie_path<OutT>(c, s, d_pts, B, N, d_q, Q, K, d_out, flag_list, &dstats->flag_count,
                                                 &tree, stats ? c->tev[4] : nullptr, &n_launch, &tree_reused)));
        if (h_out) {
            gather_rows_kernel<OutT><<<64, 256, 0, s>>>(flag_list, &dstats->flag_count, d_out, (int)K, PATCH_CAP, d_patch);
            SSDR_CHECK_CUDA(cudaGetLastError());
            n_launch += 1;
        }
        return SSDR_OK;
    };
    if (!speculate) {
        // (a pageable h_out makes this copy block the host; the count read below then simply follows it)
        if (bulk_copy) SSDR_CHECK_CUDA(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, cs));
        SSDR_TRY(d2h_sync(c, &hs, dstats, sizeof(DevStats), s));
        if (hs.flag_count) SSDR_TRY(tie_path());
    } else {
        SSDR_TRY(tie_path());
        if (bulk_copy) SSDR_CHECK_CUDA(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, cs));
    }
    if (stats) SSDR_CHECK_CUDA(cudaEventRecord(c->tev[3], s));
    if (speculate) SSDR_CHECK_CUDA(cudaMemcpyAsync(&hs, dstats, sizeof(DevStats), cudaMemcpyDeviceToHost, s));
    if (tree.error) SSDR_CHECK_CUDA(cudaMemcpyAsync(&h_err, tree.error, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    SSDR_CHECK_CUDA(cudaStreamSynchronize(s));
    if (h_out) {
        int rc = kdtree::tree_error_to_status(h_err);
        const unsigned nf = hs.flag_count < PATCH_CAP ? (unsigned)hs.flag_count : PATCH_CAP;
        std::vector<unsigned> rows(nf);
        std::vector<OutT> patch((size_t)nf * K);
        if (rc == SSDR_OK && nf && hs.flag_count <= PATCH_CAP) {
            SSDR_CHECK_CUDA(cudaMemcpyAsync(rows.data(), flag_list, (size_t)nf * 4, cudaMemcpyDeviceToHost, s));
            SSDR_CHECK_CUDA(cudaMemcpyAsync(patch.data(), d_patch, (size_t)nf * K * sizeof(OutT), cudaMemcpyDeviceToHost, s));
            SSDR_CHECK_CUDA(cudaStreamSynchronize(s));
        }
        if (fence.armed) SSDR_CHECK_CUDA(cudaStreamSynchronize(c->copy_stream));  // all rows have landed
        SSDR_TRY(rc);
        if (hs.flag_count > PATCH_CAP) {
            SSDR_TRY(d2h_sync(c, h_out, d_out, out_bytes, s));
        } else {
            for (unsigned i = 0; i < nf; ++i)
                memcpy(h_out + (size_t)rows[i] * K, patch.data() + (size_t)i * K, K * sizeof(OutT));
        }
    }
    kdtree::settle_tree_cache(c, hs.flag_count > 0 && h_err == 0);
    SSDR_TRY(kdtree::tree_error_to_status(h_err));
    const unsigned long long builds = (hs.flag_count && !tree_reused) ? B : 0;
    if (stats) {
        float ms = 0.f;
        SSDR_CHECK_CUDA(cudaEventElapsedTime(&ms, c->tev[0], c->tev[1]));
        stats->grid_build_ms = ms;
        SSDR_CHECK_CUDA(cudaEventElapsedTime(&ms, c->tev[1], c->tev[2]));
        stats->main_kernel_ms = ms;
        SSDR_CHECK_CUDA(cudaEventElapsedTime(&ms, c->tev[2], c->tev[3]));
        stats->tie_path_ms = hs.flag_count ? ms : 0.0;
        stats->tree_build_ms = 0.0;
        if (hs.flag_count) {
            SSDR_CHECK_CUDA(cudaEventElapsedTime(&ms, c->tev[2], c->tev[4]));
            stats->tree_build_ms = ms;  // includes the flag-count read-back that precedes the build
        }
        stats->queries = totalQ;
        stats->tie_rows = hs.flag_count;
        stats->tree_builds = builds;
        stats->dist_evals = hs.evals;
        stats->kernel_launches = n_launch;
    }
    return SSDR_OK;
}

template <typename OutT>
static int run_host(const float* pts, size_t B, size_t N, size_t dim, const float* q, size_t Q, size_t K, OutT* out,
                    size_t pts_stride = 0, size_t q_stride = 0) {  // floats between batch items (0 = dense)
    SSDR_REQUIRE(pts && q && out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(dim == 3, SSDR_ERR_UNSUPPORTED, "dim=%zu: only 3-D points are supported", dim);
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    if (B == 0 || Q == 0) return SSDR_OK;
    SSDR_REQUIRE(N >= 1, SSDR_ERR_INVALID, "npts must be >= 1 (the reference asserts npts != 0)");
    if (pts_stride == 0) pts_stride = N * 3;
    if (q_stride == 0) q_stride = Q * 3;
    SSDR_REQUIRE(pts_stride >= N * 3 && q_stride >= Q * 3, SSDR_ERR_INVALID, "item stride shorter than one item");
    const bool self = (q == pts && Q == N && q_stride == pts_stride);
    // items that are slices of a larger array (the pyramid's xyz[:, :N/ratio, :]) are packed by the copy engine
    auto upload = [&](void* dst, const float* src, size_t n, size_t stride) -> int {
        if (stride == n * 3 || B == 1) return h2d(c, dst, src, B * n * 3 * sizeof(float), c->stream);
        SSDR_CHECK_CUDA(cudaMemcpy2DAsync(dst, n * 3 * sizeof(float), src, stride * sizeof(float), n * 3 * sizeof(float), B,
                                          cudaMemcpyHostToDevice, c->stream));
        return SSDR_OK;
    };
    SSDR_TRY(c->ws[WS_IN_P].reserve(B * N * 3 * sizeof(float)));
    SSDR_TRY(upload(c->ws[WS_IN_P].p, pts, N, pts_stride));
    const float* d_q = c->ws[WS_IN_P].as<float>();
    if (!self) {
        SSDR_TRY(c->ws[WS_IN_Q].reserve(B * Q * 3 * sizeof(float)));
        SSDR_TRY(upload(c->ws[WS_IN_Q].p, q, Q, q_stride));
        d_q = c->ws[WS_IN_Q].as<float>();
    }
    SSDR_TRY(c->ws[WS_OUT].reserve(B * Q * K * sizeof(OutT)));
    SSDR_TRY((run_dev<OutT>(c, c->stream, c->ws[WS_IN_P].as<float>(), B, N, d_q, Q, K, c->ws[WS_OUT].as<OutT>(), nullptr,
                            K > N ? nullptr : out)));
    if (K > N) {  // only the first npts slots of each row are defined (knn_.cxx:59-67): leave the rest untouched
        SSDR_CHECK_CUDA(cudaMemcpy2DAsync(out, K * sizeof(OutT), c->ws[WS_OUT].p, K * sizeof(OutT), N * sizeof(OutT),
                                          B * Q, cudaMemcpyDeviceToHost, c->stream));
        SSDR_CHECK_CUDA(cudaStreamSynchronize(c->stream));
        return SSDR_OK;
    }
    return SSDR_OK;  // delivered by run_dev (bulk read-back overlapped with the tie path)
}

static int async_status_word(Ctx* c, cudaStream_t s, unsigned** out) {
    const void* before = c->ws[WS_ASYNC].p;
    SSDR_TRY(c->ws[WS_ASYNC].reserve(64));
    if (c->ws[WS_ASYNC].p != before) SSDR_CHECK_CUDA(cudaMemsetAsync(c->ws[WS_ASYNC].p, 0, 64, s));
    *out = c->ws[WS_ASYNC].as<unsigned>();
    return SSDR_OK;
}

// The loop of s3dis_dataset.py:164-177 in one call, every level enqueued behind the previous one.
static int pyramid_dev(Ctx* c, cudaStream_t s, const float* d_points, size_t B, size_t npts, const int32_t* ratios,
                       size_t n_levels, size_t K, long long* const* d_neigh, long long* const* d_up, bool mark_async,
                       long long* const* h_neigh = nullptr, long long* const* h_up = nullptr) {
    SSDR_REQUIRE(d_points && ratios && d_neigh && d_up, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(n_levels >= 1 && n_levels <= 16, SSDR_ERR_INVALID, "n_levels must be in [1, 16]");
    SSDR_REQUIRE(npts >= 1 && K >= 1, SSDR_ERR_INVALID, "npts and K must be >= 1");
    if (B == 0) return SSDR_OK;
    size_t n[17];
    n[0] = npts;
    size_t packed = 0;
    for (size_t l = 0; l < n_levels; ++l) {
        SSDR_REQUIRE(ratios[l] >= 1, SSDR_ERR_INVALID, "sub-sampling ratio %d of level %zu is not positive", ratios[l], l);
        n[l + 1] = n[l] / (size_t)ratios[l];
        SSDR_REQUIRE(n[l + 1] >= 1, SSDR_ERR_INVALID, "level %zu would be empty", l + 1);
        SSDR_REQUIRE(d_neigh[l] && d_up[l], SSDR_ERR_INVALID, "NULL output of level %zu", l);
        packed += B * n[l + 1] * 3;
    }
    SSDR_TRY(c->ws[WS_IN_Q].reserve(packed * sizeof(float)));
    // level l+1 = the first N_{l+1} points of every item of level l (the loader's random order makes a prefix a random
    // sub-sample, s3dis_dataset.py:166): packed by the copy engine
    const float* level[17];
    level[0] = d_points;
    float* w = c->ws[WS_IN_Q].as<float>();
    for (size_t l = 0; l < n_levels; ++l) {
        SSDR_CHECK_CUDA(cudaMemcpy2DAsync(w, n[l + 1] * 3 * sizeof(float), level[l], n[l] * 3 * sizeof(float),
                                          n[l + 1] * 3 * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
        level[l + 1] = w;
        w += B * n[l + 1] * 3;
    }
    unsigned* status = nullptr;
    SSDR_TRY(async_status_word(c, s, &status));
    g_last_launches = 0;
    static const bool branches_on = [] {
        const char* e = getenv("SSDR_KNN_BRANCHES");  // SSDR_KNN_BRANCHES=0: every call behind the previous one (A/B runs)
        return !(e && e[0] == '0');
    }();
    if (!branches_on) {
        AsyncCtx ac;
        ac.status = status;
        for (size_t l = 0; l < n_levels; ++l) {
            ac.reuse = l > 0;  // the trees of this level's cloud belong to the previous level's up-sampling query
            SSDR_TRY((run_dev<long long>(c, s, level[l], B, n[l], level[l], n[l], K, d_neigh[l], nullptr, nullptr, &ac)));
            ac.reuse = false;
            SSDR_TRY((run_dev<long long>(c, s, level[l + 1], B, n[l + 1], level[l], n[l], 1, d_up[l], nullptr, nullptr, &ac)));
            if (h_neigh)
                SSDR_CHECK_CUDA(cudaMemcpyAsync(h_neigh[l], d_neigh[l], B * n[l] * K * sizeof(long long),
                                                cudaMemcpyDeviceToHost, s));
            if (h_up)
                SSDR_CHECK_CUDA(cudaMemcpyAsync(h_up[l], d_up[l], B * n[l] * sizeof(long long), cudaMemcpyDeviceToHost, s));
        }
        return mark_async ? ctx_mark_async(c, s) : SSDR_OK;
    }
    // The calls only depend on each other through the trees of a shared support cloud (the 1-NN query onto level j and
    // the k-NN query of level j itself), so the pyramid is n_levels + 1 independent BRANCHES, one per support cloud:
    // branch j = [1-NN of level j-1 onto level j] -> [k-NN of level j].  Branch 0 stays on the caller's stream, the
    // others run on streams of their own, each in its own workspace bank, forked behind the packing copies and joined
    // before the call returns: the small levels' latency-bound kernels fill the SMs the big level leaves idle.
    struct Restore {
        Ctx* c;
        ~Restore() { ctx_use_bank(c, 0); }
    } restore{c};
    SSDR_TRY(ctx_branch(c, 0, nullptr));
    SSDR_CHECK_CUDA(cudaEventRecord(c->ev_fork, s));
    // A support cloud's branch is split once more by batch items (independent clouds): the parts of a level run side by
    // side, so one part's tree build -- a latency-bound cooperative kernel on a few dozen SMs -- overlaps the other
    // parts' query kernels.  SSDR_KNN_SPLIT = parts per level of 131072 points and more (default 3; 1 = no split).
    static const int split_env = [] {
        const char* e = getenv("SSDR_KNN_SPLIT");
        const int v = e ? atoi(e) : 3;
        return v < 1 ? 1 : (v > 3 ? 3 : v);
    }();
    const size_t S = (size_t)split_env < B ? (size_t)split_env : B;
    // host flavour: part 0 of branch 0 goes LAST and through the synchronous host path of run_dev -- its rows leave in
    // chunks under the next query launch and the tie path, the few rewritten rows follow as a patch -- while the other
    // branches, already enqueued, run beside it
    const size_t n_br = (n_levels + 1) * S;
    for (size_t tt = 0; tt < n_br; ++tt) {
        const size_t t = h_neigh ? (tt + 1) % n_br : tt;
        const size_t j = t / S, part = t % S;
        const size_t Sj = B * n[j] >= 131072 ? S : 1;  // small levels stay whole: a split only adds launches there
        if (part >= Sj) continue;
        const size_t b0 = B * part / Sj, b1 = B * (part + 1) / Sj, Bp = b1 - b0;
        cudaStream_t bs = s;
        if (t > 0) {
            SSDR_TRY(ctx_branch(c, (int)t, &bs));
            SSDR_CHECK_CUDA(cudaStreamWaitEvent(bs, c->ev_fork, 0));
        } else {
            ctx_use_bank(c, 0);
        }
        AsyncCtx ac;
        ac.status = status;
        ac.reuse = false;
        const float* sup = level[j] + b0 * n[j] * 3;  // this part's items of the support cloud
        if (j > 0)
            SSDR_TRY((run_dev<long long>(c, bs, sup, Bp, n[j], level[j - 1] + b0 * n[j - 1] * 3, n[j - 1], 1,
                                         d_up[j - 1] + b0 * n[j - 1], nullptr, nullptr, &ac)));
        if (t == 0 && h_neigh) {
            kdtree::invalidate_tree_cache(c);  // (the synchronous path would first check the previous call's trees)
            SSDR_TRY((run_dev<long long>(c, s, sup, Bp, n[0], sup, n[0], K, d_neigh[0], nullptr, h_neigh[0], nullptr)));
        } else if (j < n_levels) {
            ac.reuse = j > 0;  // the trees of this cloud belong to the up-sampling query just enqueued
            SSDR_TRY((run_dev<long long>(c, bs, sup, Bp, n[j], sup, n[j], K, d_neigh[j] + b0 * n[j] * K, nullptr, nullptr,
                                         &ac)));
        }
        // host flavour: a branch's rows start their way home on the branch's own stream, under the other branches' kernels
        if (h_up && j > 0)
            SSDR_CHECK_CUDA(cudaMemcpyAsync(h_up[j - 1] + b0 * n[j - 1], d_up[j - 1] + b0 * n[j - 1],
                                            Bp * n[j - 1] * sizeof(long long), cudaMemcpyDeviceToHost, bs));
        if (h_neigh && t > 0 && j < n_levels)
            SSDR_CHECK_CUDA(cudaMemcpyAsync(h_neigh[j] + b0 * n[j] * K, d_neigh[j] + b0 * n[j] * K,
                                            Bp * n[j] * K * sizeof(long long), cudaMemcpyDeviceToHost, bs));
        if (t > 0) SSDR_CHECK_CUDA(cudaEventRecord(c->ev_branch[t], bs));
    }
    for (size_t t = 1; t < n_br; ++t) {
        const size_t Sj = B * n[t / S] >= 131072 ? S : 1;
        if (t % S >= Sj) continue;
        SSDR_CHECK_CUDA(cudaStreamWaitEvent(s, c->ev_branch[t], 0));
    }
    ctx_use_bank(c, 0);
    // the call returns with its work in flight: later calls are ordered behind it (a capture records no event)
    return mark_async ? ctx_mark_async(c, s) : SSDR_OK;
}

// A pyramid call launches ~45 kernels, memsets and copies whose arguments do not depend on the data (every decision of
// the tie path is taken on the device), so a caller that repeats the call with the same buffers -- a loader that
// refills one device batch -- gets the whole sequence as ONE captured CUDA graph from the second call on.  The graph is
// dropped when any argument changes or a workspace of this thread has been re-allocated since the capture.
struct PyramidGraph {
    bool seen = false, valid = false;
    const float* pts = nullptr;
    size_t B = 0, npts = 0, n_levels = 0, K = 0;
    int32_t ratios[16] = {};
    const void* neigh[16] = {};
    const void* up[16] = {};
    cudaStream_t stream = nullptr;
    int device = -1;
    unsigned long long generation = 0, launches = 0;
    cudaGraphExec_t exec = nullptr;
};
static thread_local PyramidGraph g_pyr;

static bool pyramid_key_matches(const PyramidGraph& g, Ctx* c, cudaStream_t s, const float* pts, size_t B, size_t npts,
                                const int32_t* ratios, size_t n_levels, size_t K, int64_t* const* neigh,
                                int64_t* const* up) {
    (void)s;  // the graph lives on the library's own stream (the caller's may be the legacy stream, which cannot capture)
    if (!g.seen || g.pts != pts || g.B != B || g.npts != npts || g.n_levels != n_levels || g.K != K ||
        g.device != c->device)
        return false;
    for (size_t l = 0; l < n_levels; ++l)
        if (g.ratios[l] != ratios[l] || g.neigh[l] != neigh[l] || g.up[l] != up[l]) return false;
    return true;
}

}  // namespace knn
}  // namespace ssdr

using namespace ssdr;

extern "C" {
int ssdr_knn_pyramid_dev(const float* d_points, size_t batch_size, size_t npts, const int32_t* ratios, size_t n_levels,
                         size_t K, int64_t* const* d_neigh, int64_t* const* d_up, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    long long* const* neigh = reinterpret_cast<long long* const*>(d_neigh);
    long long* const* up = reinterpret_cast<long long* const*>(d_up);
    static const bool graphs_on = [] {
        const char* e = getenv("SSDR_KNN_GRAPH");
        return !(e && e[0] == '0');
    }();
    knn::PyramidGraph& g = knn::g_pyr;
    const bool args_ok = d_points && ratios && d_neigh && d_up && n_levels >= 1 && n_levels <= 16;
    // replay on the library's stream, forked from and joined to the caller's stream
    auto replay = [&]() -> int {
        SSDR_CHECK_CUDA(cudaEventRecord(c->ev, s));
        SSDR_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev, 0));
        SSDR_CHECK_CUDA(cudaGraphLaunch(g.exec, c->stream));
        SSDR_CHECK_CUDA(cudaEventRecord(c->ev_main, c->stream));
        SSDR_CHECK_CUDA(cudaStreamWaitEvent(s, c->ev_main, 0));
        knn::g_last_launches = g.launches;
        return ctx_mark_async(c, s);
    };
    if (graphs_on && args_ok &&
        knn::pyramid_key_matches(g, c, s, d_points, batch_size, npts, ratios, n_levels, K, d_neigh, d_up)) {
        if (g.valid && g.generation == ws_generation()) return replay();
        // second call with these arguments (the first one sized every workspace): capture it
        if (g.exec) {
            cudaGraphExecDestroy(g.exec);
            g.exec = nullptr;
        }
        g.valid = false;
        const unsigned long long gen0 = ws_generation();
        SSDR_CHECK_CUDA(cudaStreamSynchronize(c->stream));
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            const int rc = knn::pyramid_dev(c, c->stream, d_points, batch_size, npts, ratios, n_levels, K, neigh, up, false);
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc == SSDR_OK && e == cudaSuccess && graph && gen0 == ws_generation() &&
                cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                g.valid = true;
                g.generation = gen0;
                g.launches = knn::g_last_launches;
            }
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (g.valid) return replay();
        }
        cudaGetLastError();
        // capture not possible here: fall through to the eager path
    } else if (graphs_on && args_ok) {
        g.seen = true;
        g.valid = false;
        g.pts = d_points;
        g.B = batch_size;
        g.npts = npts;
        g.n_levels = n_levels;
        g.K = K;
        g.device = c->device;
        for (size_t l = 0; l < n_levels; ++l) {
            g.ratios[l] = ratios[l];
            g.neigh[l] = d_neigh[l];
            g.up[l] = d_up[l];
        }
    }
    return knn::pyramid_dev(c, s, d_points, batch_size, npts, ratios, n_levels, K, neigh, up, true);
}
unsigned long long ssdr_knn_pyramid_launches(void) { return knn::g_last_launches; }
int ssdr_knn_pyramid(const float* batch_xyz, size_t batch_size, size_t npts, size_t dim, const int32_t* ratios,
                     size_t n_levels, size_t K, int64_t* const* neigh, int64_t* const* up) {
    using namespace knn;
    SSDR_REQUIRE(batch_xyz && ratios && neigh && up, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(dim == 3, SSDR_ERR_UNSUPPORTED, "dim=%zu: only 3-D points are supported", dim);
    SSDR_REQUIRE(n_levels >= 1 && n_levels <= 16, SSDR_ERR_INVALID, "n_levels must be in [1, 16]");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    if (batch_size == 0) return SSDR_OK;
    const size_t B = batch_size;
    size_t n = npts, rows = 0;
    for (size_t l = 0; l < n_levels; ++l) {
        SSDR_REQUIRE(ratios[l] >= 1 && n / (size_t)ratios[l] >= 1, SSDR_ERR_INVALID, "bad sub-sampling ratio at level %zu", l);
        SSDR_REQUIRE(K <= n, SSDR_ERR_UNSUPPORTED, "K=%zu exceeds the %zu points of level %zu", K, n, l);
        rows += B * n;
        n /= (size_t)ratios[l];
    }
    cudaStream_t s = c->stream;
    static const bool trace = getenv("SSDR_TRACE") != nullptr;
    auto now = [] {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    const double t0 = trace ? now() : 0.0;
    SSDR_TRY(c->ws[WS_IN_P].reserve(B * npts * 3 * sizeof(float)));
    SSDR_TRY(c->ws[WS_OUT].reserve(rows * (K + 1) * sizeof(long long)));
    SSDR_TRY(h2d(c, c->ws[WS_IN_P].p, batch_xyz, B * npts * 3 * sizeof(float), s));
    const double t1 = trace ? now() : 0.0;
    long long* dn[16];
    long long* du[16];
    long long* w = c->ws[WS_OUT].as<long long>();
    n = npts;
    for (size_t l = 0; l < n_levels; ++l) {
        dn[l] = w;
        w += B * n * K;
        du[l] = w;
        w += B * n;
        n /= (size_t)ratios[l];
    }
    // no return path may leave a copy into the caller's arrays in flight
    struct Fence {
        Ctx* c;
        ~Fence() {
            cudaStreamSynchronize(c->stream);
            for (int k = 1; k < Ctx::MAX_BRANCH; ++k)
                if (c->branch_stream[k]) cudaStreamSynchronize(c->branch_stream[k]);
        }
    };
    int rc;
    double t2 = 0.0;
    {
        Fence fence{c};
        rc = pyramid_dev(c, s, c->ws[WS_IN_P].as<float>(), B, npts, ratios, n_levels, K, dn, du, false,
                         reinterpret_cast<long long* const*>(neigh), reinterpret_cast<long long* const*>(up));
        if (trace) t2 = now();
    }
    if (trace)
        fprintf(stderr, "ssdr_knn_pyramid: upload %.3f ms, enqueue + branch 0 %.3f ms, other branches %.3f ms\n", t1 - t0,
                t2 - t1, now() - t2);
    SSDR_TRY(rc);
    unsigned* word = nullptr;
    SSDR_TRY(async_status_word(c, s, &word));
    unsigned h = 0;
    SSDR_TRY(d2h_sync(c, &h, word, sizeof(h), s));
    if (h) SSDR_CHECK_CUDA(cudaMemsetAsync(word, 0, sizeof(unsigned), s));
    return kdtree::tree_error_to_status(h);
}
int ssdr_knn_status(void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* w = nullptr;
    SSDR_TRY(knn::async_status_word(c, s, &w));
    unsigned h = 0;
    SSDR_TRY(d2h_sync(c, &h, w, sizeof(h), s));
    if (h) SSDR_CHECK_CUDA(cudaMemsetAsync(w, 0, sizeof(unsigned), s));
    return kdtree::tree_error_to_status(h);
}
int ssdr_knn(const float* points, size_t npts, size_t dim, const float* queries, size_t nqueries, size_t K,
             int64_t* indices) {
    return knn::run_host<long long>(points, 1, npts, dim, queries, nqueries, K, reinterpret_cast<long long*>(indices));
}
int ssdr_knn_batch(const float* batch_data, size_t batch_size, size_t npts, size_t dim, const float* queries,
                   size_t nqueries, size_t K, int64_t* batch_indices) {
    return knn::run_host<long long>(batch_data, batch_size, npts, dim, queries, nqueries, K,
                                    reinterpret_cast<long long*>(batch_indices));
}
int ssdr_knn_batch_strided(const float* batch_data, size_t batch_size, size_t npts, size_t dim, size_t data_item_stride,
                           const float* queries, size_t nqueries, size_t query_item_stride, size_t K,
                           int64_t* batch_indices) {
    return knn::run_host<long long>(batch_data, batch_size, npts, dim, queries, nqueries, K,
                                    reinterpret_cast<long long*>(batch_indices), data_item_stride, query_item_stride);
}
int ssdr_knn_batch_dev(const float* d_points, size_t batch_size, size_t npts, const float* d_queries, size_t nqueries,
                       size_t K, int64_t* d_indices, void* stream, ssdr_knn_stats* stats) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return knn::run_dev<long long>(c, (cudaStream_t)stream, d_points, batch_size, npts, d_queries,
                                   nqueries, K, reinterpret_cast<long long*>(d_indices), stats);
}
// Diagnostic: build the nanoflann-identical tree of one cloud on the device and copy it back (tests compare it with
// a sequential CPU build).  Node arrays must hold 3*npts+64 entries.
// Diagnostic: build B trees of npts points each (device timing only) and return %globaltimer marks in ns:
// [0] start, [1] roots done, [2..9] top level k done, [10] top done, [11] CTA 0 done, [12] last CTA done.
int ssdr_knn_debug_build_timing(const float* points, size_t B, size_t npts, uint64_t* marks16) {
    SSDR_REQUIRE(points && npts >= 1 && B >= 1 && marks16, SSDR_ERR_INVALID, "bad argument");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    cudaStream_t s = c->stream;
    SSDR_TRY(c->ws[knn::WS_IN_P].reserve(B * npts * 3 * sizeof(float)));
    SSDR_TRY(h2d(c, c->ws[knn::WS_IN_P].p, points, B * npts * 3 * sizeof(float), s));
    SSDR_TRY(c->ws[knn::WS_STATS].reserve(16 * sizeof(unsigned long long) + 64));
    unsigned long long* d_marks = reinterpret_cast<unsigned long long*>(c->ws[knn::WS_STATS].as<char>() + 64);
    for (int rep = 0; rep < 3; ++rep) {  // the last repetition is the warm one
        kdtree::Tree t;
        SSDR_TRY(kdtree::alloc_tree(c, s, B, npts, &t));
        SSDR_CHECK_CUDA(cudaMemsetAsync(d_marks, 0, 16 * sizeof(unsigned long long), s));
        t.tstamps = d_marks;
        SSDR_TRY(kdtree::launch_build(c, s, c->ws[knn::WS_IN_P].as<float>(), t));
        {
            unsigned h_err2 = 0;
            SSDR_TRY(d2h_sync(c, &h_err2, t.error, sizeof(unsigned), s));
            SSDR_TRY(kdtree::tree_error_to_status(h_err2));
        }
    }
    return d2h_sync(c, marks16, d_marks, 16 * sizeof(unsigned long long), s);
}

int ssdr_knn_debug_tree(const float* points, size_t npts, uint32_t* vind_out, uint32_t* n_nodes_out, uint32_t* left,
                        uint32_t* right, int32_t* child1, int32_t* child2, int32_t* divfeat, float* divlow,
                        float* divhigh) {
    SSDR_REQUIRE(points && npts >= 1 && vind_out && n_nodes_out, SSDR_ERR_INVALID, "bad argument");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    cudaStream_t s = c->stream;
    SSDR_TRY(c->ws[knn::WS_IN_P].reserve(npts * 3 * sizeof(float)));
    SSDR_TRY(h2d(c, c->ws[knn::WS_IN_P].p, points, npts * 3 * sizeof(float), s));
    kdtree::Tree t;
    SSDR_TRY(kdtree::alloc_tree(c, s, 1, npts, &t));
    SSDR_TRY(kdtree::launch_build(c, s, c->ws[knn::WS_IN_P].as<float>(), t));
    {
            unsigned h_err2 = 0;
            SSDR_TRY(d2h_sync(c, &h_err2, t.error, sizeof(unsigned), s));
            SSDR_TRY(kdtree::tree_error_to_status(h_err2));
        }
    SSDR_TRY(d2h_sync(c, n_nodes_out, t.node_count, sizeof(unsigned), s));
    const size_t nn = *n_nodes_out;
    SSDR_REQUIRE(nn <= t.cap, SSDR_ERR_CUDA, "node count %zu exceeds capacity", nn);
    {
        float4* hp = (float4*)malloc(npts * sizeof(float4));
        SSDR_REQUIRE(hp, SSDR_ERR_NOMEM, "host allocation failed");
        int rc0 = d2h_sync(c, hp, t.pp, npts * sizeof(float4), s);
        for (size_t i = 0; rc0 == SSDR_OK && i < npts; ++i) memcpy(&vind_out[i], &hp[i].w, 4);
        free(hp);
        SSDR_TRY(rc0);
    }
    kdtree::NodeRec* h = (kdtree::NodeRec*)malloc(nn * sizeof(kdtree::NodeRec));
    SSDR_REQUIRE(h, SSDR_ERR_NOMEM, "host allocation failed");
    int rc = d2h_sync(c, h, t.nodes, nn * sizeof(kdtree::NodeRec), s);
    // node ids are handed out in blocks (gaps are never written): export the reachable nodes in pre-order
    size_t n_out = 0;
    if (rc == SSDR_OK) {
        std::vector<unsigned> stack, order;
        std::vector<int> remap(nn, -1);
        stack.push_back(0);
        while (!stack.empty()) {
            const unsigned n = stack.back();
            stack.pop_back();
            remap[n] = (int)order.size();
            order.push_back(n);
            if (h[n].c1 >= 0) {
                if ((size_t)h[n].c1 >= nn || (size_t)h[n].c2 >= nn) {
                    rc = set_error(SSDR_ERR_CUDA, "corrupt tree: child id out of range");
                    break;
                }
                stack.push_back((unsigned)h[n].c2);
                stack.push_back((unsigned)h[n].c1);
            }
        }
        n_out = order.size();
        for (size_t k = 0; rc == SSDR_OK && k < n_out; ++k) {
            const kdtree::NodeRec& r = h[order[k]];
            if (left) left[k] = r.l;
            if (right) right[k] = r.r;
            if (child1) child1[k] = r.c1 >= 0 ? remap[r.c1] : -1;
            if (child2) child2[k] = r.c2 >= 0 ? remap[r.c2] : -1;
            if (divfeat) divfeat[k] = r.feat;
            if (divlow) divlow[k] = r.divlow;
            if (divhigh) divhigh[k] = r.divhigh;
        }
    }
    free(h);
    *n_nodes_out = (uint32_t)n_out;
    return rc;
}
int ssdr_knn_batch_dev_i32(const float* d_points, size_t batch_size, size_t npts, const float* d_queries,
                           size_t nqueries, size_t K, int32_t* d_indices, void* stream, ssdr_knn_stats* stats) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return knn::run_dev<int>(c, (cudaStream_t)stream, d_points, batch_size, npts, d_queries,
                             nqueries, K, d_indices, stats);
}
}
