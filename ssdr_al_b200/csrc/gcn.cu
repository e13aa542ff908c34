// gcn.cu -- the superpoint adjacency and the feature propagation that feed the FPS loop, on sm_100a.
//
// Replaces the dense numpy part of fps_adj_all (fps_gcn_cpu.py:95-117) and of GCN_FPS_sampling (:153-167):
//
//   A_ed, A_cd = 1e10 everywhere, per room block: centre distances / chamfer distances          (:60-101)
//   adj = exp(-(A_ed + A_cd));  adj += -1.0 * eye;  d_inv = power(rowsum(adj), -1), inf -> 0     (:104-111)
//   adj = adj @ diag(d_inv);  adj = adj + eye                                                    (:114-116)
//   optional: keep the gcn_top largest entries of every row                                      (:155-161)
//   V_0 = V;  V_{k+1} = adj @ V_k;  out = V_0 + V_1 + ... + V_g                                   (:163-167)
//
// What is reproduced exactly: the assembly (float64 centre distances in numpy's order, A_ed + A_cd), the diagonal
// arithmetic, the row sums (numpy's pairwise summation order, so d_inv sees the same float64 input), adj[i][j] * d_inv[j]
// (a product with a diagonal matrix adds exact zeros), the sequential sum of the V_k.  What is not bit-exact by
// construction: exp() and pow(x, -1) (libm vs CUDA: <= 1 ulp each) and the summation order of the float64 matrix
// products (BLAS leaves it unspecified) -- the tests state the tolerance (1e-12 relative on adj and on the features).
// Rows of an N x N float64 matrix stream once per product: the propagation is bound by HBM (8 N^2 bytes per product).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace ssdr {
namespace gcn {

enum { WS_ADJ = 0, WS_ADJ2 = 1, WS_VA = 2, WS_VB = 3, WS_SUM = 4, WS_DINV = 5, WS_BLK = 6, WS_CD = 7 };

struct Handle {
    size_t N = 0;
    double* adj = nullptr;  // (N, N) device
    int device = 0;
};

__global__ void fill_kernel(double* __restrict__ S, unsigned long long n, double v) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        S[i] = v;
}

// S[ref[j]][ref[k]] = sqrt(sum((c_k - c_j)^2)) + cd[j][k] for the superpoints j, k of one room (blockIdx.y)
__global__ void block_kernel(double* __restrict__ S, unsigned long long N, const long long* __restrict__ block_off,
                             const long long* __restrict__ cd_off, const long long* __restrict__ ref,
                             const double* __restrict__ centres, const double* __restrict__ cd) {
    const unsigned b = blockIdx.y;
    const long long o = block_off[b], n = block_off[b + 1] - o;
    const double* c = centres + 3 * o;
    const double* cdb = cd + cd_off[b];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n * n; t += (long long)gridDim.x * blockDim.x) {
        const long long j = t / n, k = t - j * n;
        // ssdr = centres - centres[j]; dist = sqrt(np.sum(ssdr * ssdr, axis=1)): three terms added left to right
        const double dx = __dsub_rn(c[3 * k], c[3 * j]), dy = __dsub_rn(c[3 * k + 1], c[3 * j + 1]),
                     dz = __dsub_rn(c[3 * k + 2], c[3 * j + 2]);
        const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        S[(unsigned long long)ref[o + j] * N + (unsigned long long)ref[o + k]] = __dadd_rn(d, cdb[t]);
    }
}

// adj = exp(-S) + (-1.0 * eye)
__global__ void exp_kernel(double* __restrict__ S, unsigned long long N) {
    const unsigned long long n = N * N;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = i / N;
        const double e = exp(-S[i]);
        S[i] = (i - r * N == r) ? __dadd_rn(e, -1.0) : e;
    }
}

// numpy's pairwise_sum of one contiguous float64 row (loops_utils.h.src): leaves of <= 128 values with eight strided
// accumulators, combined by recursive halving.  One warp per row: the lanes take the leaves in turn, lane 0 replays the
// recursion over the leaf sums.
constexpr int RS_WARPS = 4;
constexpr unsigned RS_MAX_LEAVES = 1024;  // leaves of one row: N <= 58k superpoints (the matrix alone is 27 GB there)

__device__ double leaf_sum(const double* a, unsigned n) {
    if (n < 8) {
        double r = 0.0;
        for (unsigned i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    unsigned i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

struct Frame {
    unsigned lo, n;
    int phase;
    double left;
};

// The recursion of pairwise_sum over [0, n): pass 0 (sums == nullptr) lists the leaves, pass 1 combines their sums.
__device__ double walk(unsigned n, unsigned* leaf_lo, unsigned* leaf_n, unsigned* n_leaves, const double* sums) {
    Frame st[40];
    int sp = 0;
    unsigned next = 0;
    double ret = 0.0;
    st[sp++] = Frame{0u, n, 0, 0.0};
    while (sp > 0) {
        Frame& f = st[sp - 1];
        if (f.n <= 128u) {
            if (sums) ret = sums[next];
            else if (next < RS_MAX_LEAVES) {
                leaf_lo[next] = f.lo;
                leaf_n[next] = f.n;
            }
            ++next;
            --sp;
            continue;
        }
        unsigned n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.phase == 0) {
            f.phase = 1;
            st[sp++] = Frame{f.lo, n2, 0, 0.0};
        } else if (f.phase == 1) {
            f.left = ret;
            f.phase = 2;
            st[sp++] = Frame{f.lo + n2, f.n - n2, 0, 0.0};
        } else {
            ret = __dadd_rn(f.left, ret);
            --sp;
        }
    }
    if (n_leaves) *n_leaves = next;
    return ret;
}

// leaves of a row of length N (the same for every row): computed once by one thread
__global__ void leaves_kernel(unsigned N, unsigned* __restrict__ leaf_lo, unsigned* __restrict__ leaf_n,
                              unsigned* __restrict__ n_leaves) {
    if (blockIdx.x == 0 && threadIdx.x == 0) walk(N, leaf_lo, leaf_n, n_leaves, nullptr);
}

// d_inv[i] = power(rowsum(adj[i]), -1) with inf -> 0
__global__ void __launch_bounds__(RS_WARPS * 32) rowsum_kernel(const double* __restrict__ adj, unsigned N,
                                                               const unsigned* __restrict__ leaf_lo,
                                                               const unsigned* __restrict__ leaf_n,
                                                               const unsigned* __restrict__ n_leaves_p,
                                                               double* __restrict__ d_inv) {
    __shared__ double s_sums[RS_WARPS][RS_MAX_LEAVES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned nl = *n_leaves_p;
    for (unsigned row = blockIdx.x * RS_WARPS + warp; row < N; row += gridDim.x * RS_WARPS) {
        const double* a = adj + (unsigned long long)row * N;
        for (unsigned q = lane; q < nl; q += 32) s_sums[warp][q] = leaf_sum(a + leaf_lo[q], leaf_n[q]);
        __syncwarp();
        if (lane == 0) {
            const double s = walk(N, nullptr, nullptr, nullptr, s_sums[warp]);
            const double inv = pow(s, -1.0);
            d_inv[row] = isinf(inv) ? 0.0 : inv;
        }
        __syncwarp();
    }
}

// adj[i][j] = adj[i][j] * d_inv[j] + (i == j)
__global__ void scale_kernel(double* __restrict__ adj, unsigned long long N, const double* __restrict__ d_inv) {
    const unsigned long long n = N * N;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = i / N, c = i - r * N;
        const double v = __dmul_rn(adj[i], d_inv[c]);
        adj[i] = c == r ? __dadd_rn(v, 1.0) : v;
    }
}

// ---- top-k mask of every row (np.argsort(adj, axis=1)[:, -k:]): entries are >= 0, so their bit patterns order like
// the values.  One CTA per row: eight 8-bit radix-select passes find the k-th largest value t; kept: everything above
// t, and of the entries equal to t the ones with the highest column index (what a stable ascending sort puts last).
constexpr int TK_THREADS = 256;
__global__ void __launch_bounds__(TK_THREADS) topk_mask_kernel(const double* __restrict__ adj, double* __restrict__ out,
                                                                unsigned N, unsigned k) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned s_need;
    for (unsigned row = blockIdx.x; row < N; row += gridDim.x) {
        const unsigned long long* a = reinterpret_cast<const unsigned long long*>(adj) + (unsigned long long)row * N;
        if (threadIdx.x == 0) {
            s_prefix = 0ull;
            s_need = k;
        }
        __syncthreads();
        for (int shift = 56; shift >= 0; shift -= 8) {
            s_hist[threadIdx.x] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
            for (unsigned j = threadIdx.x; j < N; j += TK_THREADS) {
                const unsigned long long v = a[j];
                if ((v & mask) == prefix) atomicAdd(&s_hist[(unsigned)(v >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned need = s_need;
                int d = 255;
                for (; d > 0; --d) {
                    if (s_hist[d] >= need) break;
                    need -= s_hist[d];
                }
                s_prefix = prefix | ((unsigned long long)d << shift);
                s_need = need;
            }
            __syncthreads();
        }
        const unsigned long long t = s_prefix;  // the k-th largest value; s_need of the entries equal to it are kept
        const unsigned need_eq = s_need;
        // entries equal to t, counted from the highest column down: a block-wide suffix count in chunks
        __shared__ unsigned s_cnt[TK_THREADS];
        __shared__ unsigned s_taken;
        if (threadIdx.x == 0) s_taken = 0;
        __syncthreads();
        double* o = out + (unsigned long long)row * N;
        for (long long base = (long long)N - 1; base >= 0; base -= TK_THREADS) {
            const long long j = base - threadIdx.x;  // thread 0 takes the highest column of the chunk
            const unsigned long long v = j >= 0 ? a[j] : 0ull;
            const bool eq = j >= 0 && v == t;
            s_cnt[threadIdx.x] = eq ? 1u : 0u;
            __syncthreads();
            // inclusive prefix over the chunk in thread order (Hillis-Steele; 256 entries)
            for (int off = 1; off < TK_THREADS; off <<= 1) {
                const unsigned add = threadIdx.x >= (unsigned)off ? s_cnt[threadIdx.x - off] : 0u;
                __syncthreads();
                s_cnt[threadIdx.x] += add;
                __syncthreads();
            }
            const unsigned before = s_taken + s_cnt[threadIdx.x] - (eq ? 1u : 0u);  // equal entries in higher columns
            if (j >= 0) {
                const bool keep = v > t || (eq && before < need_eq);
                o[j] = keep ? __longlong_as_double((long long)v) : 0.0;
            }
            __syncthreads();
            if (threadIdx.x == TK_THREADS - 1) s_taken += s_cnt[TK_THREADS - 1];
            __syncthreads();
        }
    }
}

// ---- out = adj @ V: (N, N) x (N, D) float64.  A CTA owns MM_ROWS rows; the k loop streams adj once (coalesced along
// the row), V chunks go through shared memory; every thread accumulates one row for a strip of columns with fma chains
// in ascending k.
constexpr int MM_ROWS = 32;
constexpr int MM_K = 32;
constexpr int MM_THREADS = 256;
constexpr int MM_DT = 32;  // columns of V per pass

__global__ void __launch_bounds__(MM_THREADS) matmul_kernel(const double* __restrict__ A, const double* __restrict__ V,
                                                            double* __restrict__ out, unsigned N, unsigned D) {
    __shared__ double s_a[MM_ROWS][MM_K + 1];
    __shared__ double s_v[MM_K][MM_DT];
    const unsigned r0 = blockIdx.x * MM_ROWS;
    const int tr = threadIdx.x / 8;          // 32 rows
    const int tc = (threadIdx.x % 8) * 4;    // 8 strips of 4 columns = 32 columns per pass
    for (unsigned d0 = 0; d0 < D; d0 += MM_DT) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (unsigned k0 = 0; k0 < N; k0 += MM_K) {
            for (int t = threadIdx.x; t < MM_ROWS * MM_K; t += MM_THREADS) {
                const int rr = t / MM_K, kk = t % MM_K;
                const unsigned r = r0 + rr, kx = k0 + kk;
                s_a[rr][kk] = (r < N && kx < N) ? A[(unsigned long long)r * N + kx] : 0.0;
            }
            for (int t = threadIdx.x; t < MM_K * MM_DT; t += MM_THREADS) {
                const int kk = t / MM_DT, dd = t % MM_DT;
                const unsigned kx = k0 + kk, dx = d0 + dd;
                s_v[kk][dd] = (kx < N && dx < D) ? V[(unsigned long long)kx * D + dx] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < MM_K; ++kk) {
                const double a = s_a[tr][kk];
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = fma(a, s_v[kk][tc + u], acc[u]);
            }
            __syncthreads();
        }
        const unsigned r = r0 + tr;
        if (r < N) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (d0 + tc + u < D) out[(unsigned long long)r * D + d0 + tc + u] = acc[u];
        }
    }
}

// ---- the same product on the FP64 tensor cores (mma.sync m8n8k4, DMMA): the fast path for even N and D with 16-byte
// aligned rows.  A CTA owns 32 rows x 32 columns, a warp 8 rows x 32 columns (four 8 x 8 accumulator tiles); the k loop
// moves 64-wide chunks of adj and V into a double-buffered shared-memory ring with 16-byte cp.async (zero fill beyond
// the edges), rows padded by four doubles so that the fragment loads of a half-warp hit 16 distinct bank pairs.
constexpr int DM_ROWS = 32, DM_COLS = 32, DM_K = 64, DM_THREADS = 128;
constexpr int DM_AS = DM_K + 4;     // row stride of the adj tile (doubles): 68 = 4 mod 16
constexpr int DM_VS = DM_COLS + 4;  // row stride of the V tile: 36 = 4 mod 16
constexpr size_t DM_SMEM = 2 * (size_t)(DM_ROWS * DM_AS + DM_K * DM_VS) * sizeof(double);

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(bytes));
}

__global__ void __launch_bounds__(DM_THREADS) matmul_dmma_kernel(const double* __restrict__ A, const double* __restrict__ V,
                                                                 double* __restrict__ out, unsigned N, unsigned D) {
    extern __shared__ __align__(16) unsigned char dm_smem[];
    double* s_a[2];
    double* s_v[2];
    s_a[0] = reinterpret_cast<double*>(dm_smem);
    s_a[1] = s_a[0] + DM_ROWS * DM_AS;
    s_v[0] = s_a[1] + DM_ROWS * DM_AS;
    s_v[1] = s_v[0] + DM_K * DM_VS;
    const unsigned r0 = blockIdx.x * DM_ROWS, d0 = blockIdx.y * DM_COLS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;  // fragment coordinates: A[fr][fk], B[fk][fr], C[fr][2 fk .. 2 fk + 1]
    double acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
    const unsigned nchunk = (N + DM_K - 1) / DM_K;
    auto issue = [&](unsigned ch, int buf) {
        const unsigned k0 = ch * DM_K;
        // adj tile: 32 rows x 32 pairs of doubles
        for (int t = tid; t < DM_ROWS * (DM_K / 2); t += DM_THREADS) {
            const int rr = t / (DM_K / 2), kp = t % (DM_K / 2);
            const unsigned r = r0 + rr, kx = k0 + 2 * kp;
            const bool ok = r < N && kx < N;
            cp_async16_zfill(s_a[buf] + rr * DM_AS + 2 * kp, A + (ok ? (unsigned long long)r * N + kx : 0ull), ok);
        }
        // V tile: 64 rows x 16 pairs
        for (int t = tid; t < DM_K * (DM_COLS / 2); t += DM_THREADS) {
            const int kk = t / (DM_COLS / 2), cp = t % (DM_COLS / 2);
            const unsigned kx = k0 + kk, dx = d0 + 2 * cp;
            const bool ok = kx < N && dx < D;
            cp_async16_zfill(s_v[buf] + kk * DM_VS + 2 * cp, V + (ok ? (unsigned long long)kx * D + dx : 0ull), ok);
        }
        asm volatile("cp.async.commit_group;");
    };
    issue(0, 0);
    for (unsigned ch = 0; ch < nchunk; ++ch) {
        const int buf = (int)(ch & 1u);
        if (ch + 1 < nchunk) {
            issue(ch + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const double* a = s_a[buf] + (warp * 8 + fr) * DM_AS + fk;
        const double* v = s_v[buf] + fk * DM_VS + fr;
#pragma unroll 4
        for (int kk = 0; kk < DM_K; kk += 4) {
            const double af = a[kk];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const double bf = v[kk * DM_VS + 8 * t];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(acc[t][0]), "+d"(acc[t][1])
                             : "d"(af), "d"(bf));
            }
        }
        __syncthreads();
    }
    const unsigned r = r0 + warp * 8 + fr;
    if (r < N) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const unsigned c = d0 + 8 * t + 2 * fk;
            if (c < D) out[(unsigned long long)r * D + c] = acc[t][0];
            if (c + 1 < D) out[(unsigned long long)r * D + c + 1] = acc[t][1];
        }
    }
}

__global__ void add_kernel(double* __restrict__ sum, const double* __restrict__ v, unsigned long long n) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        sum[i] = __dadd_rn(sum[i], v[i]);
}

static unsigned blocks_for(const Ctx* c, unsigned long long n, int threads) {
    const unsigned long long want = (n + threads - 1) / threads;
    const unsigned long long cap = (unsigned long long)c->sm_count * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

static int normalise(Ctx* c, cudaStream_t s, double* S, size_t N) {
    const unsigned long long nn = (unsigned long long)N * N;
    exp_kernel<<<blocks_for(c, nn, 256), 256, 0, s>>>(S, N);
    SSDR_TRY(c->ws[WS_DINV].reserve(N * sizeof(double) + (2 * RS_MAX_LEAVES + 4) * sizeof(unsigned)));
    double* d_inv = c->ws[WS_DINV].as<double>();
    unsigned* leaf_lo = reinterpret_cast<unsigned*>(d_inv + N);
    unsigned* leaf_n = leaf_lo + RS_MAX_LEAVES;
    unsigned* n_leaves = leaf_n + RS_MAX_LEAVES;
    leaves_kernel<<<1, 32, 0, s>>>((unsigned)N, leaf_lo, leaf_n, n_leaves);
    rowsum_kernel<<<(unsigned)c->sm_count * 2, RS_WARPS * 32, 0, s>>>(S, (unsigned)N, leaf_lo, leaf_n, n_leaves, d_inv);
    scale_kernel<<<blocks_for(c, nn, 256), 256, 0, s>>>(S, N, d_inv);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

}  // namespace gcn
}  // namespace ssdr

using namespace ssdr;

extern "C" {

int ssdr_gcn_adjacency_f64(size_t N, size_t n_blocks, const int64_t* block_off, const int64_t* ref, const double* centres,
                           const double* cd, void** handle) {
    using namespace gcn;
    SSDR_REQUIRE(handle, SSDR_ERR_INVALID, "NULL handle pointer");
    *handle = nullptr;
    SSDR_REQUIRE(N >= 1 && N < 0x7FFFFFFFull, SSDR_ERR_INVALID, "N=%zu out of range", N);
    SSDR_REQUIRE(N <= (size_t)RS_MAX_LEAVES * 57, SSDR_ERR_UNSUPPORTED, "N=%zu: more than %u superpoints", N,
                 RS_MAX_LEAVES * 57);
    SSDR_REQUIRE(n_blocks == 0 || (block_off && ref && centres && cd), SSDR_ERR_INVALID, "NULL pointer");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    cudaStream_t s = c->stream;
    size_t total = 0, cd_total = 0;
    std::vector<long long> cd_off(n_blocks + 1, 0);
    size_t max_n = 0;
    for (size_t b = 0; b < n_blocks; ++b) {
        SSDR_REQUIRE(block_off[b + 1] >= block_off[b], SSDR_ERR_INVALID, "block offsets must ascend");
        const size_t n = (size_t)(block_off[b + 1] - block_off[b]);
        cd_off[b + 1] = cd_off[b] + (long long)(n * n);
        max_n = n > max_n ? n : max_n;
    }
    if (n_blocks) {
        total = (size_t)block_off[n_blocks];
        cd_total = (size_t)cd_off[n_blocks];
        for (size_t i = 0; i < total; ++i)
            SSDR_REQUIRE(ref[i] >= 0 && (size_t)ref[i] < N, SSDR_ERR_INVALID, "ref[%zu]=%lld outside [0, %zu)", i,
                         (long long)ref[i], N);
    }
    double* S = nullptr;
    SSDR_CHECK_CUDA(cudaMalloc(&S, N * N * sizeof(double)));
    struct Guard {
        double* p;
        ~Guard() {
            if (p) cudaFree(p);
        }
    } guard{S};
    const unsigned long long nn = (unsigned long long)N * N;
    fill_kernel<<<blocks_for(c, nn, 256), 256, 0, s>>>(S, nn, 1e10 + 1e10);
    if (n_blocks && total) {
        const size_t meta_bytes = (2 * (n_blocks + 1) + total) * sizeof(long long);
        SSDR_TRY(c->ws[WS_BLK].reserve(meta_bytes + total * 3 * sizeof(double)));
        SSDR_TRY(c->ws[WS_CD].reserve(cd_total * sizeof(double) + 8));
        long long* d_off = c->ws[WS_BLK].as<long long>();
        long long* d_cdoff = d_off + (n_blocks + 1);
        long long* d_ref = d_cdoff + (n_blocks + 1);
        double* d_c = reinterpret_cast<double*>(d_ref + total);
        double* d_cd = c->ws[WS_CD].as<double>();
        SSDR_TRY(h2d(c, d_off, block_off, (n_blocks + 1) * sizeof(long long), s));
        SSDR_TRY(h2d(c, d_cdoff, cd_off.data(), (n_blocks + 1) * sizeof(long long), s));
        SSDR_TRY(h2d(c, d_ref, ref, total * sizeof(long long), s));
        SSDR_TRY(h2d(c, d_c, centres, total * 3 * sizeof(double), s));
        SSDR_TRY(h2d(c, d_cd, cd, cd_total * sizeof(double), s));
        SSDR_REQUIRE(n_blocks <= 65535, SSDR_ERR_UNSUPPORTED, "more than 65535 rooms");
        unsigned bx = (unsigned)((max_n * max_n + 255) / 256);
        bx = bx < 1 ? 1 : (bx > 1024 ? 1024 : bx);
        block_kernel<<<dim3(bx, (unsigned)n_blocks), 256, 0, s>>>(S, N, d_off, d_cdoff, d_ref, d_c, d_cd);
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    SSDR_TRY(normalise(c, s, S, N));
    SSDR_CHECK_CUDA(cudaStreamSynchronize(s));  // cd_off (host vector) and the caller's arrays are free again
    Handle* h = new Handle;
    h->N = N;
    h->adj = S;
    h->device = c->device;
    guard.p = nullptr;
    *handle = h;
    return SSDR_OK;
}

int ssdr_gcn_fetch(void* handle, double* adj_out) {
    using namespace gcn;
    SSDR_REQUIRE(handle && adj_out, SSDR_ERR_INVALID, "NULL pointer");
    Handle* h = static_cast<Handle*>(handle);
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    return d2h_sync(c, adj_out, h->adj, h->N * h->N * sizeof(double), c->stream);
}

int ssdr_gcn_free(void* handle) {
    using namespace gcn;
    if (!handle) return SSDR_OK;
    Handle* h = static_cast<Handle*>(handle);
    if (h->adj) cudaFree(h->adj);
    delete h;
    return SSDR_OK;
}

int ssdr_gcn_propagate_f64(void* handle, const double* adj_host, size_t N, const double* V, size_t D, int gcn_number,
                           int gcn_top, double* out) {
    using namespace gcn;
    SSDR_REQUIRE((handle || adj_host) && V && out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1 && N < 0x7FFFFFFFull && D >= 1 && D < 0x7FFFFFFFull, SSDR_ERR_INVALID, "bad shape (%zu, %zu)", N, D);
    SSDR_REQUIRE(gcn_number >= 0, SSDR_ERR_INVALID, "gcn_number must be >= 0");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    cudaStream_t s = c->stream;
    const double* adj = nullptr;
    if (handle) {
        Handle* h = static_cast<Handle*>(handle);
        SSDR_REQUIRE(h->N == N, SSDR_ERR_INVALID, "adjacency is %zu x %zu, features have %zu rows", h->N, h->N, N);
        adj = h->adj;
    } else {
        SSDR_TRY(c->ws[WS_ADJ].reserve(N * N * sizeof(double)));
        SSDR_TRY(h2d(c, c->ws[WS_ADJ].p, adj_host, N * N * sizeof(double), s));
        adj = c->ws[WS_ADJ].as<double>();
    }
    if (gcn_top > 0 && gcn_number > 0) {
        SSDR_REQUIRE((size_t)gcn_top <= N, SSDR_ERR_INVALID, "gcn_top=%d exceeds the %zu superpoints", gcn_top, N);
        SSDR_TRY(c->ws[WS_ADJ2].reserve(N * N * sizeof(double)));
        topk_mask_kernel<<<(unsigned)(N < (size_t)c->sm_count * 8 ? N : (size_t)c->sm_count * 8), TK_THREADS, 0, s>>>(
            adj, c->ws[WS_ADJ2].as<double>(), (unsigned)N, (unsigned)gcn_top);
        adj = c->ws[WS_ADJ2].as<double>();
    }
    const size_t vb = N * D * sizeof(double);
    SSDR_TRY(c->ws[WS_VA].reserve(vb));
    SSDR_TRY(c->ws[WS_VB].reserve(vb));
    SSDR_TRY(c->ws[WS_SUM].reserve(vb));
    double* va = c->ws[WS_VA].as<double>();
    double* vbuf = c->ws[WS_VB].as<double>();
    double* sum = c->ws[WS_SUM].as<double>();
    SSDR_TRY(h2d(c, va, V, vb, s));
    SSDR_CHECK_CUDA(cudaMemcpyAsync(sum, va, vb, cudaMemcpyDeviceToDevice, s));
    // tensor-core path: 16-byte cp.async needs even N and D (all workspaces and the adjacency are 256-byte aligned)
    static const bool dmma_on = [] {
        const char* e = getenv("SSDR_GCN_DMMA");  // SSDR_GCN_DMMA=0: the plain fma kernel (A/B runs)
        return !(e && e[0] == '0');
    }();
    const bool dmma = dmma_on && N % 2 == 0 && D % 2 == 0 && (reinterpret_cast<size_t>(adj) & 15) == 0;
    if (dmma)
        SSDR_CHECK_CUDA(cudaFuncSetAttribute(matmul_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DM_SMEM));
    for (int g = 0; g < gcn_number; ++g) {
        if (dmma)
            matmul_dmma_kernel<<<dim3((unsigned)((N + DM_ROWS - 1) / DM_ROWS), (unsigned)((D + DM_COLS - 1) / DM_COLS)),
                                 DM_THREADS, DM_SMEM, s>>>(adj, va, vbuf, (unsigned)N, (unsigned)D);
        else
            matmul_kernel<<<(unsigned)((N + MM_ROWS - 1) / MM_ROWS), MM_THREADS, 0, s>>>(adj, va, vbuf, (unsigned)N, (unsigned)D);
        add_kernel<<<blocks_for(c, (unsigned long long)N * D, 256), 256, 0, s>>>(sum, vbuf, (unsigned long long)N * D);
        double* t = va;
        va = vbuf;
        vbuf = t;
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    return d2h_sync(c, out, sum, vb, s);
}

}  // extern "C"
