// primitives.cuh -- hand-written device-wide building blocks used by the hot path:
//   * exclusive_scan_u32 : two launches (block sums + last-block-done scan of the sums, then the per-block scan)
//   * radix_sort_pairs   : stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass, over the significant
//                          key bits only; per pass: block histograms -> digit-major exclusive scan -> stable scatter
//                          (warp match_any ranking keeps equal keys in input order)
// They replace the cub::DeviceScan / DeviceRadixSort / DeviceSelect calls of the first version.
#pragma once
#include "common.cuh"

namespace ssdr {
namespace prim {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 8;                               // consecutive elements per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;        // 8192 elements per block

// inclusive block scan of one u32 per thread (1024 threads); returns the inclusive value, *total = block sum
__device__ __forceinline__ unsigned block_scan_u32(unsigned v, unsigned* s_warp, unsigned* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < nw ? s_warp[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned base = warp ? s_warp[warp - 1] : 0u;
    *total = s_warp[nw - 1];
    __syncthreads();
    return v + base;
}

// phase 1: per-tile sums; the last block to finish turns them into exclusive tile offsets (and the grand total)
static __global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(const unsigned* __restrict__ in, size_t n,
                                                                 unsigned* __restrict__ tile_off, unsigned ntiles,
                                                                 unsigned* __restrict__ ticket,
                                                                 unsigned* __restrict__ total_out) {
    __shared__ unsigned s_warp[32];
    __shared__ bool s_last;
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) sum += in[base + k];
    unsigned tot;
    block_scan_u32(sum, s_warp, &tot);
    if (threadIdx.x == 0) {
        tile_off[blockIdx.x] = tot;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan of the tile sums by this one block, SCAN_THREADS sums per round
    unsigned carry = 0;
    for (unsigned t0 = 0; t0 < ntiles; t0 += SCAN_THREADS) {
        const unsigned i = t0 + threadIdx.x;
        const unsigned v = i < ntiles ? __ldcg(&tile_off[i]) : 0u;
        unsigned rt;
        const unsigned inc = block_scan_u32(v, s_warp, &rt);
        if (i < ntiles) tile_off[i] = carry + inc - v;
        carry += rt;
    }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = carry;
        *ticket = 0;  // ready for the next use
    }
}

// phase 2: exclusive scan inside each tile + tile offset
static __global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const unsigned* __restrict__ in,
                                                                  unsigned* __restrict__ out, size_t n,
                                                                  const unsigned* __restrict__ tile_off) {
    __shared__ unsigned s_warp[32];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0u;
        sum += v[k];
    }
    unsigned tot;
    const unsigned inc = block_scan_u32(sum, s_warp, &tot);
    unsigned run = tile_off[blockIdx.x] + inc - sum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

// scratch: [0] ticket (must be zero on entry; it is reset on exit), [2 ..] tile offsets.  The ticket sits at a FIXED
// position so that scans of different sizes can share one scratch area.
static inline size_t scan_scratch_words(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 2; }

static int exclusive_scan_u32(const unsigned* d_in, unsigned* d_out, size_t n, unsigned* d_scratch,
                              unsigned* d_total /* nullable */, cudaStream_t s) {
    if (n == 0) return SSDR_OK;
    const unsigned ntiles = (unsigned)((n + SCAN_TILE - 1) / SCAN_TILE);
    unsigned* ticket = d_scratch;
    unsigned* tile_off = d_scratch + 2;
    scan_sums_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(d_in, n, tile_off, ntiles, ticket, d_total);
    scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(d_in, d_out, n, tile_off);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// ---- radix sort ------------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                           // keys per thread, blocked
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;        // 2048 keys per block
constexpr int RS_BITS = 8;
constexpr int RS_BINS = 1 << RS_BITS;

// histogram of the current digit per block, stored digit-major: hist[digit * nblocks + block]
static __global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned long long* __restrict__ keys, size_t n,
                                                             int shift, unsigned* __restrict__ hist, unsigned nblocks) {
    __shared__ unsigned s_h[RS_BINS];
    for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) s_h[i] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const size_t i = base + (size_t)k * RS_THREADS + threadIdx.x;  // striped: coalesced, order is irrelevant here
        if (i < n) atomicAdd(&s_h[(unsigned)(keys[i] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) hist[(size_t)i * nblocks + blockIdx.x] = s_h[i];
}

// stable scatter: element e of the block (blocked order: thread t owns e = t*ITEMS + k) goes to
// offs[digit][block] + (number of earlier elements of the block with the same digit)
static __global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned long long* __restrict__ keys_in,
                                                                const unsigned* __restrict__ vals_in,
                                                                unsigned long long* __restrict__ keys_out,
                                                                unsigned* __restrict__ vals_out, size_t n, int shift,
                                                                const unsigned* __restrict__ offs, unsigned nblocks) {
    constexpr int NWARP = RS_THREADS / 32;
    __shared__ unsigned s_cnt[NWARP][RS_BINS];   // per warp: digit counts of the current round, then exclusive bases
    __shared__ unsigned s_base[RS_BINS];         // running per-digit base inside the block (rounds so far)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) s_base[i] = offs[(size_t)i * nblocks + blockIdx.x];
    // rounds of RS_THREADS consecutive elements keep the block order = input order: round k covers elements
    // [k*RS_THREADS, (k+1)*RS_THREADS), thread t takes element k*RS_THREADS + t
#pragma unroll 1
    for (int k = 0; k < RS_ITEMS; ++k) {
        for (int i = threadIdx.x; i < NWARP * RS_BINS; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
        __syncthreads();
        const size_t i = base + (size_t)k * RS_THREADS + threadIdx.x;
        const bool in = i < n;
        unsigned long long key = 0;
        unsigned val = 0, digit = RS_BINS;  // RS_BINS = "no element"
        if (in) {
            key = keys_in[i];
            val = vals_in[i];
            digit = (unsigned)(key >> shift) & (RS_BINS - 1);
        }
        // rank inside the warp among equal digits (lanes in increasing order = input order)
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (in && rank_in_warp == 0) s_cnt[warp][digit] = __popc(peers);
        __syncthreads();
        // exclusive prefix over warps per digit, and advance the block base
        for (int d = threadIdx.x; d < RS_BINS; d += RS_THREADS) {
            unsigned run = s_base[d];
#pragma unroll
            for (int w = 0; w < NWARP; ++w) {
                const unsigned cnt = s_cnt[w][d];
                s_cnt[w][d] = run;
                run += cnt;
            }
            s_base[d] = run;
        }
        __syncthreads();
        if (in) {
            const unsigned pos = s_cnt[warp][digit] + rank_in_warp;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
    }
}

// scratch words needed by radix_sort_pairs for n elements
static inline size_t rs_scratch_words(size_t n) {
    const size_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    const size_t nh = nblocks * RS_BINS;
    return 2 * nh + scan_scratch_words(nh) + 8;
}

// Sorts (keys, vals) by the low `bits` bits of the keys, stable.  Ping-pongs between the two buffer pairs; returns
// (in *cur) which pair holds the result: 0 = (keys_a, vals_a), 1 = (keys_b, vals_b).  Scratch must be zeroed once
// at allocation (only the scan ticket needs it; it resets itself).
static int radix_sort_pairs(unsigned long long* keys_a, unsigned* vals_a, unsigned long long* keys_b, unsigned* vals_b,
                            size_t n, int bits, unsigned* d_scratch, int* cur, cudaStream_t s) {
    *cur = 0;
    if (n == 0) return SSDR_OK;
    const unsigned nblocks = (unsigned)((n + RS_TILE - 1) / RS_TILE);
    const size_t nh = (size_t)nblocks * RS_BINS;
    unsigned* scan_scr = d_scratch;  // scan ticket at a fixed position for every n (see scan_scratch_words)
    unsigned* hist = d_scratch + scan_scratch_words(nh);
    unsigned* offs = hist + nh;
    for (int shift = 0; shift < bits; shift += RS_BITS) {
        const unsigned long long* kin = *cur ? keys_b : keys_a;
        const unsigned* vin = *cur ? vals_b : vals_a;
        unsigned long long* kout = *cur ? keys_a : keys_b;
        unsigned* vout = *cur ? vals_a : vals_b;
        rs_hist_kernel<<<nblocks, RS_THREADS, 0, s>>>(kin, n, shift, hist, nblocks);
        SSDR_TRY(exclusive_scan_u32(hist, offs, nh, scan_scr, nullptr, s));
        rs_scatter_kernel<<<nblocks, RS_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, offs, nblocks);
        *cur ^= 1;
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

}  // namespace prim
}  // namespace ssdr
