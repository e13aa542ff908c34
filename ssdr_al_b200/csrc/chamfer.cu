// chamfer.cu -- chamfer adjacency between the S small clouds ("superpoints") of one room:
// fps_gcn_cpu.py:12-38 (create_cd / chamfer_distance), the step that feeds the FPS loop.
//
// Reference: every cloud is centred on its bbox centre (float64), a sklearn KDTree per cloud answers 1-NN queries, and
//   cd[c][i] = mean_{p in cloud i} min_{q in cloud c} |p - q|  +  mean_{q in cloud c} min_{p in cloud i} |q - p|
// with |.| = sqrt(((dx*dx + dy*dy) + dz*dz)) in float64 (sklearn's EuclideanDistance.rdist accumulates the squares in
// this order, the square root is taken once on the minimum) and np.mean = numpy's pairwise summation / n.
// Only DISTANCES leave the function, so no tie order is involved: brute force over all pairs gives the same minima,
// and the mean is reproduced with numpy's exact pairwise order -- the result is bit-identical to the reference unless
// the KD tree's own bound rounding hides a neighbour that is closer by less than an ulp (tests allow 1e-12 relative).
//
// Kernel: block (a, g) owns source cloud a and the target clouds b = g, g+G, ...; a target is streamed through shared
// memory in tiles, every thread keeps the running squared minima of a few source points in registers (float64, no FMA),
// the square roots land in a per-block scratch row, and the block sums that row in numpy's order:
// leaves of <= 128 values with eight strided accumulators, combined by the recursive halving of pairwise_sum.
#include "common.cuh"

namespace ssdr {
namespace chamfer {

enum { WS_PTS = 0, WS_OFF = 1, WS_DIR = 2, WS_SCRATCH = 3, WS_OUT = 4 };

constexpr int CT = 256;          // threads per block
constexpr int PA = 4;            // source points per thread and pass
constexpr int TB = 1024;         // target points per shared-memory tile
constexpr int MAX_LEAVES = 1024; // pairwise-sum leaves per source cloud (<= 128 values each, >= 57 once split)
// number of pairwise-sum leaves of n values is at most n / LEAF_MIN + 2 (a split leaf holds >= 57 values)
constexpr unsigned LEAF_MIN = 57;
__host__ __device__ inline size_t leaf_slots(size_t n) { return n / LEAF_MIN + 2; }

// numpy's pairwise_sum for n <= 128 (loops_utils.h.src: @TYPE@_pairwise_sum)
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

// Walks pairwise_sum's recursion over [0, n).  With sums == nullptr it only records the leaves (lo, len) number
// first .. first + MAX_LEAVES - 1 (a window: the tables live in shared memory, a large cloud has more leaves than that);
// otherwise it consumes the leaf sums in the same left-to-right order and returns the total.
__device__ double walk(unsigned n, unsigned* leaf_lo, unsigned* leaf_n, unsigned* n_leaves, const double* sums,
                       unsigned first = 0) {
    Frame st[40];
    int sp = 0;
    unsigned next = 0;
    double ret = 0.0;
    st[sp++] = Frame{0u, n, 0, 0.0};
    while (sp > 0) {
        Frame& f = st[sp - 1];
        if (f.n <= 128u) {
            if (sums) ret = sums[next];
            else if (leaf_lo && next >= first && next - first < (unsigned)MAX_LEAVES) {
                leaf_lo[next - first] = f.lo;
                leaf_n[next - first] = f.n;
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

// squared distances of PU register-resident source points to the nt target points of the tile, running minima
template <int PU>
__device__ __forceinline__ void scan_tile(const double* sx, const double* sy, const double* sz, unsigned nt,
                                          const double (&ax)[PA], const double (&ay)[PA], const double (&az)[PA],
                                          double (&best)[PA]) {
    for (unsigned j = 0; j < nt; ++j) {
        const double bx = sx[j], by = sy[j], bz = sz[j];
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            const double dx = __dsub_rn(ax[u], bx), dy = __dsub_rn(ay[u], by), dz = __dsub_rn(az[u], bz);
            const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            best[u] = d < best[u] ? d : best[u];
        }
    }
}

// shared memory of one block working on directed pairs
struct PairSmem {
    double sx[TB], sy[TB], sz[TB];
    unsigned lo[MAX_LEAVES], n[MAX_LEAVES];
    double sum[MAX_LEAVES];
    unsigned nleaf;
};

// mean over the points of cloud a of the distance to the nearest point of cloud b (whole block; every thread returns
// the value).  `mins` is the block's private scratch of na doubles; sm.lo / sm.n / sm.nleaf hold the leaf layout of na.
// `big_sums`: leaf_slots(na) doubles of global scratch, used when the cloud has more leaves than the shared table.
__device__ double directed_mean(const double* __restrict__ pts, unsigned long long a0, unsigned na,
                                unsigned long long b0, unsigned nb, double* __restrict__ mins, PairSmem& sm,
                                double* __restrict__ big_sums) {
    for (unsigned k0 = 0; k0 < na; k0 += CT * PA) {
        const int pu = (int)min((unsigned)PA, (na - k0 + CT - 1) / CT);
        double ax[PA], ay[PA], az[PA], best[PA];
#pragma unroll
        for (int u = 0; u < PA; ++u) {
            const unsigned k = k0 + (unsigned)u * CT + threadIdx.x;
            const unsigned kk = k < na ? k : na - 1;
            ax[u] = pts[3 * (a0 + kk)];
            ay[u] = pts[3 * (a0 + kk) + 1];
            az[u] = pts[3 * (a0 + kk) + 2];
            best[u] = INFINITY;
        }
        for (unsigned t0 = 0; t0 < nb; t0 += TB) {
            const unsigned nt = min((unsigned)TB, nb - t0);
            __syncthreads();
            for (unsigned j = threadIdx.x; j < nt; j += CT) {
                sm.sx[j] = pts[3 * (b0 + t0 + j)];
                sm.sy[j] = pts[3 * (b0 + t0 + j) + 1];
                sm.sz[j] = pts[3 * (b0 + t0 + j) + 2];
            }
            __syncthreads();
            // only as many register slots as this pass has source points for (block-uniform): small clouds would
            // otherwise spend most of the FP64 issue slots on duplicates of their last point
            switch (pu) {
                case 1: scan_tile<1>(sm.sx, sm.sy, sm.sz, nt, ax, ay, az, best); break;
                case 2: scan_tile<2>(sm.sx, sm.sy, sm.sz, nt, ax, ay, az, best); break;
                case 3: scan_tile<3>(sm.sx, sm.sy, sm.sz, nt, ax, ay, az, best); break;
                default: scan_tile<PA>(sm.sx, sm.sy, sm.sz, nt, ax, ay, az, best); break;
            }
        }
#pragma unroll
        for (int u = 0; u < PA; ++u) {
            const unsigned k = k0 + (unsigned)u * CT + threadIdx.x;
            if (k < na) mins[k] = __dsqrt_rn(best[u]);
        }
    }
    __syncthreads();  // mins complete (global writes of this block are visible to it after the barrier)
    const unsigned nleaf = sm.nleaf;
    const bool big = nleaf > (unsigned)MAX_LEAVES;  // the leaf table is then filled window by window
    double* sums = big ? big_sums : sm.sum;
    for (unsigned first = 0; first < nleaf; first += MAX_LEAVES) {
        if (big) {
            if (threadIdx.x == 0) walk(na, sm.lo, sm.n, nullptr, nullptr, first);
            __syncthreads();
        }
        const unsigned cnt = min((unsigned)MAX_LEAVES, nleaf - first);
        for (unsigned l = threadIdx.x; l < cnt; l += CT) sums[first + l] = leaf_sum(mins + sm.lo[l], sm.n[l]);
        __syncthreads();
    }
    __shared__ double s_mean;
    if (threadIdx.x == 0) s_mean = __ddiv_rn(walk(na, nullptr, nullptr, nullptr, sums), (double)na);
    __syncthreads();
    const double r = s_mean;
    __syncthreads();  // the next pair overwrites mins, sm.sum and s_mean
    return r;
}

__global__ void __launch_bounds__(CT) directed_kernel(const double* __restrict__ pts, const long long* __restrict__ off,
                                                      unsigned S, unsigned G, unsigned long long T,
                                                      double* __restrict__ scratch /* [G][T] mins, then [G][leaf area] */,
                                                      double* __restrict__ A /* [S][S]: A[a][b] = mean_a min_b */) {
    __shared__ PairSmem sm;
    const unsigned a = blockIdx.x, g = blockIdx.y;
    const unsigned long long a0 = (unsigned long long)off[a];
    const unsigned na = (unsigned)(off[a + 1] - off[a]);
    if (na == 0) return;
    double* mins = scratch + (size_t)g * T + a0;
    // leaf sums of a large cloud: block (a, g) owns leaf_slots(na) doubles at a0 / LEAF_MIN + 2a inside g's leaf area
    const size_t leaf_area = (size_t)(T / LEAF_MIN) + 2 * (size_t)S + 2;
    double* big_sums = scratch + (size_t)G * T + (size_t)g * leaf_area + (size_t)(a0 / LEAF_MIN) + 2 * (size_t)a;
    if (threadIdx.x == 0) walk(na, sm.lo, sm.n, &sm.nleaf, nullptr);  // the leaf layout depends on na only
    __syncthreads();
    for (unsigned b = g; b < S; b += G) {
        if (b == a) continue;
        const unsigned long long b0 = (unsigned long long)off[b];
        const unsigned nb = (unsigned)(off[b + 1] - off[b]);
        if (nb == 0) continue;
        const double m = directed_mean(pts, a0, na, b0, nb, mins, sm, big_sums);
        if (threadIdx.x == 0) A[(size_t)a * S + b] = m;
    }
}

// ---- farthest_superpoint_sample (sampler2.py:49-80): FPS over superpoints, distance = squared centroid distance +
// chamfer distance to the current pick; the chamfer ROW of the pick is computed when it is needed, like the reference
// does.  row_kernel: block i -> row[i] = mean_i min_c + mean_c min_i; step_kernel: one block folds the row into the
// running minimum (strict '<', sampler2.py:75-76) and takes the first arg-max (:78).  The pick travels between the
// kernels in device memory, so the whole loop is enqueued without a host round trip.
__global__ void __launch_bounds__(CT) row_kernel(const double* __restrict__ pts, const long long* __restrict__ off, unsigned S,
                                                 const int* __restrict__ picks, unsigned step, unsigned long long T,
                                                 unsigned nmax,
                                                 double* __restrict__ scratch /* [T] + [S][nmax] + [S][leaf_slots(nmax)] */,
                                                 double* __restrict__ row) {
    __shared__ PairSmem sm;
    const unsigned i = blockIdx.x;
    const unsigned c = (unsigned)picks[step];
    if (i == c) {
        if (threadIdx.x == 0) row[i] = 0.0;
        return;
    }
    const unsigned long long i0 = (unsigned long long)off[i], c0 = (unsigned long long)off[c];
    const unsigned ni = (unsigned)(off[i + 1] - off[i]), nc = (unsigned)(off[c + 1] - off[c]);
    if (threadIdx.x == 0) walk(ni, sm.lo, sm.n, &sm.nleaf, nullptr);
    __syncthreads();
    double* big_sums = scratch + T + (size_t)S * nmax + (size_t)i * leaf_slots(nmax);
    const double av1 = directed_mean(pts, i0, ni, c0, nc, scratch + i0, sm, big_sums);                      // :18, :20
    if (threadIdx.x == 0) walk(nc, sm.lo, sm.n, &sm.nleaf, nullptr);
    __syncthreads();
    const double av2 = directed_mean(pts, c0, nc, i0, ni, scratch + T + (size_t)i * nmax, sm, big_sums);    // :19, :21
    if (threadIdx.x == 0) row[i] = __dadd_rn(av1, av2);
}

// squared centroid distance in the centroids' OWN dtype (the caller's float32 ply coordinates stay float32 through
// `np.sum((cents - cur) ** 2, axis=-1)`, sampler2.py:69; only np.add with the float64 chamfer row widens the result)
__device__ __forceinline__ double centroid_sqdist(const double* c, unsigned i, double cx, double cy, double cz) {
    const double dx = __dsub_rn(c[3 * i], cx), dy = __dsub_rn(c[3 * i + 1], cy), dz = __dsub_rn(c[3 * i + 2], cz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));  // np.sum of 3: sequential
}
__device__ __forceinline__ double centroid_sqdist(const float* c, unsigned i, float cx, float cy, float cz) {
    const float dx = __fsub_rn(c[3 * i], cx), dy = __fsub_rn(c[3 * i + 1], cy), dz = __fsub_rn(c[3 * i + 2], cz);
    return (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

template <typename CT_>
__global__ void __launch_bounds__(1024) step_kernel(const CT_* __restrict__ cent, unsigned S, const double* __restrict__ row,
                                                    double* __restrict__ distance, int* __restrict__ picks, unsigned step) {
    __shared__ double s_d[32];
    __shared__ unsigned s_i[32];
    const unsigned c = (unsigned)picks[step];
    const CT_ cx = cent[3 * c], cy = cent[3 * c + 1], cz = cent[3 * c + 2];
    double bd = -INFINITY;
    unsigned bi = 0xFFFFFFFFu;
    for (unsigned i = threadIdx.x; i < S; i += blockDim.x) {
        const double e = centroid_sqdist(cent, i, cx, cy, cz);
        const double d = __dadd_rn(e, row[i]);
        double cur = distance[i];
        if (d < cur) {
            cur = d;
            distance[i] = d;
        }
        if (cur > bd) {  // ascending i inside a thread: strict '>' keeps the lowest index
            bd = cur;
            bi = i;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, m);
        const unsigned oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (od > bd || (od == bd && oi < bi)) {
            bd = od;
            bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_d[threadIdx.x >> 5] = bd;
        s_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned w = 1; w < blockDim.x / 32; ++w)
            if (s_d[w] > bd || (s_d[w] == bd && s_i[w] < bi)) {
                bd = s_d[w];
                bi = s_i[w];
            }
        picks[step + 1] = (int)bi;
    }
}

__global__ void fill_f64_kernel(double* p, unsigned n, double v) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void symmetrize_kernel(const double* __restrict__ A, unsigned S, double* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned long long)S * S) return;
    const unsigned r = (unsigned)(i / S), c = (unsigned)(i % S);
    // cd[c_idx][i] = mean over cloud i of the distance to cloud c_idx + mean over cloud c_idx of the distance to cloud i
    out[i] = r == c ? 0.0 : __dadd_rn(A[(size_t)c * S + r], A[(size_t)r * S + c]);
}

static int run_dev(Ctx* c, cudaStream_t s, const double* d_pts, const long long* d_off, const long long* h_off, size_t S,
                   double* d_out) {
    SSDR_REQUIRE(S >= 1, SSDR_ERR_INVALID, "no clouds");
    SSDR_REQUIRE(S <= 46340, SSDR_ERR_UNSUPPORTED, "more than 46340 clouds");
    const size_t T = (size_t)h_off[S];
    for (size_t i = 0; i < S; ++i) {
        SSDR_REQUIRE(h_off[i + 1] >= h_off[i], SSDR_ERR_INVALID, "offsets must not decrease");
        SSDR_REQUIRE((size_t)(h_off[i + 1] - h_off[i]) < 0xFFFFFFFFull, SSDR_ERR_UNSUPPORTED,
                     "cloud %zu has more than 2^32-2 points", i);
    }
    unsigned G = (unsigned)((4 * (size_t)c->sm_count + S - 1) / S);
    if (G < 1) G = 1;
    if (G > S) G = (unsigned)S;
    SSDR_TRY(c->ws[WS_DIR].reserve(S * S * sizeof(double)));
    const size_t leaf_area = T / LEAF_MIN + 2 * S + 2;  // see directed_kernel
    SSDR_TRY(c->ws[WS_SCRATCH].reserve((size_t)G * ((T ? T : 1) + leaf_area) * sizeof(double)));
    double* A = c->ws[WS_DIR].as<double>();
    SSDR_CHECK_CUDA(cudaMemsetAsync(A, 0, S * S * sizeof(double), s));
    if (T) {
        directed_kernel<<<dim3((unsigned)S, G), CT, 0, s>>>(d_pts, d_off, (unsigned)S, G, (unsigned long long)T,
                                                           c->ws[WS_SCRATCH].as<double>(), A);
        SSDR_CHECK_CUDA(cudaGetLastError());
    }
    symmetrize_kernel<<<(unsigned)((S * S + 255) / 256), 256, 0, s>>>(A, (unsigned)S, d_out);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

static int run_superpoint_fps(Ctx* c, cudaStream_t s, const double* d_pts, const long long* d_off, const long long* h_off,
                              size_t S, const void* d_cent, bool cent_f32, int trigger, size_t n_samples, int* d_picks) {
    SSDR_REQUIRE(S >= 1 && n_samples >= 1, SSDR_ERR_INVALID, "no clouds or no samples");
    SSDR_REQUIRE(trigger >= 0 && (size_t)trigger < S, SSDR_ERR_INVALID, "trigger_idx out of range");
    const size_t T = (size_t)h_off[S];
    size_t nmax = 0;
    for (size_t i = 0; i < S; ++i) {
        SSDR_REQUIRE(h_off[i + 1] > h_off[i], SSDR_ERR_INVALID, "empty or negative-size cloud %zu", i);
        const size_t n = (size_t)(h_off[i + 1] - h_off[i]);
        SSDR_REQUIRE(n < 0xFFFFFFFFull, SSDR_ERR_UNSUPPORTED, "cloud %zu has more than 2^32-2 points", i);
        nmax = n > nmax ? n : nmax;
    }
    const size_t scratch_doubles = T + S * nmax + S * leaf_slots(nmax);
    SSDR_REQUIRE(scratch_doubles * sizeof(double) <= ((size_t)64 << 30), SSDR_ERR_UNSUPPORTED,
                 "scratch for %zu clouds of up to %zu points exceeds 64 GB", S, nmax);
    SSDR_TRY(c->ws[WS_SCRATCH].reserve(scratch_doubles * sizeof(double)));
    SSDR_TRY(c->ws[WS_DIR].reserve(2 * S * sizeof(double)));
    double* row = c->ws[WS_DIR].as<double>();
    double* distance = row + S;
    fill_f64_kernel<<<(unsigned)((S + 255) / 256), 256, 0, s>>>(distance, (unsigned)S, 1e10);  // sampler2.py:64
    SSDR_CHECK_CUDA(cudaMemsetAsync(d_picks, 0, n_samples * sizeof(int), s));
    SSDR_CHECK_CUDA(cudaMemcpyAsync(d_picks, &trigger, sizeof(int), cudaMemcpyHostToDevice, s));
    for (size_t st = 0; st + 1 < n_samples; ++st) {
        row_kernel<<<(unsigned)S, CT, 0, s>>>(d_pts, d_off, (unsigned)S, d_picks, (unsigned)st, (unsigned long long)T,
                                             (unsigned)nmax, c->ws[WS_SCRATCH].as<double>(), row);
        if (cent_f32)
            step_kernel<float><<<1, 1024, 0, s>>>((const float*)d_cent, (unsigned)S, row, distance, d_picks, (unsigned)st);
        else
            step_kernel<double><<<1, 1024, 0, s>>>((const double*)d_cent, (unsigned)S, row, distance, d_picks, (unsigned)st);
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

}  // namespace chamfer
}  // namespace ssdr

using namespace ssdr;

extern "C" {

int ssdr_chamfer_matrix_f64(const double* points, const int64_t* offsets, size_t S, double* out) {
    SSDR_REQUIRE(points && offsets && out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(S >= 1 && offsets[0] == 0, SSDR_ERR_INVALID, "offsets must start at 0 and hold S + 1 entries");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    const size_t T = (size_t)offsets[S];
    SSDR_TRY(c->ws[chamfer::WS_PTS].reserve((T ? T : 1) * 3 * sizeof(double)));
    SSDR_TRY(c->ws[chamfer::WS_OFF].reserve((S + 1) * sizeof(long long)));
    SSDR_TRY(c->ws[chamfer::WS_OUT].reserve(S * S * sizeof(double)));
    SSDR_TRY(h2d(c, c->ws[chamfer::WS_PTS].p, points, T * 3 * sizeof(double), c->stream));
    SSDR_TRY(h2d(c, c->ws[chamfer::WS_OFF].p, offsets, (S + 1) * sizeof(long long), c->stream));
    SSDR_TRY(chamfer::run_dev(c, c->stream, c->ws[chamfer::WS_PTS].as<double>(), c->ws[chamfer::WS_OFF].as<long long>(),
                              reinterpret_cast<const long long*>(offsets), S, c->ws[chamfer::WS_OUT].as<double>()));
    return d2h_sync(c, out, c->ws[chamfer::WS_OUT].p, S * S * sizeof(double), c->stream);
}

int ssdr_superpoint_fps(const double* points, const int64_t* offsets, size_t S, const void* centroids, int centroid_dtype,
                        int32_t trigger_idx, size_t n_samples, int32_t* out) {
    SSDR_REQUIRE(points && offsets && centroids && out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(centroid_dtype == SSDR_F32 || centroid_dtype == SSDR_F64, SSDR_ERR_INVALID,
                 "centroid_dtype must be SSDR_F32 or SSDR_F64");
    const size_t csize = centroid_dtype == SSDR_F32 ? sizeof(float) : sizeof(double);
    SSDR_REQUIRE(S >= 1 && offsets[0] == 0, SSDR_ERR_INVALID, "offsets must start at 0 and hold S + 1 entries");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    const size_t T = (size_t)offsets[S];
    SSDR_TRY(c->ws[chamfer::WS_PTS].reserve((T ? T : 1) * 3 * sizeof(double)));
    SSDR_TRY(c->ws[chamfer::WS_OFF].reserve((S + 1) * sizeof(long long)));
    SSDR_TRY(c->ws[chamfer::WS_OUT].reserve(S * 3 * sizeof(double) + (n_samples + 1) * sizeof(int)));
    void* d_cent = c->ws[chamfer::WS_OUT].p;
    int* d_picks = reinterpret_cast<int*>(c->ws[chamfer::WS_OUT].as<double>() + S * 3);
    SSDR_TRY(h2d(c, c->ws[chamfer::WS_PTS].p, points, T * 3 * sizeof(double), c->stream));
    SSDR_TRY(h2d(c, c->ws[chamfer::WS_OFF].p, offsets, (S + 1) * sizeof(long long), c->stream));
    SSDR_TRY(h2d(c, d_cent, centroids, S * 3 * csize, c->stream));
    SSDR_TRY(chamfer::run_superpoint_fps(c, c->stream, c->ws[chamfer::WS_PTS].as<double>(),
                                         c->ws[chamfer::WS_OFF].as<long long>(),
                                         reinterpret_cast<const long long*>(offsets), S, d_cent,
                                         centroid_dtype == SSDR_F32, trigger_idx, n_samples, d_picks));
    return d2h_sync(c, out, d_picks, n_samples * sizeof(int), c->stream);
}
int ssdr_superpoint_fps_f64(const double* points, const int64_t* offsets, size_t S, const double* centroids,
                            int32_t trigger_idx, size_t n_samples, int32_t* out) {
    return ssdr_superpoint_fps(points, offsets, S, centroids, SSDR_F64, trigger_idx, n_samples, out);
}

int ssdr_chamfer_matrix_f64_dev(const double* d_points, const int64_t* d_offsets, const int64_t* h_offsets, size_t S,
                                double* d_out, void* stream) {
    SSDR_REQUIRE(d_points && d_offsets && h_offsets && d_out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(S >= 1 && h_offsets[0] == 0, SSDR_ERR_INVALID, "offsets must start at 0 and hold S + 1 entries");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(ctx_order(c, (cudaStream_t)stream));
    return chamfer::run_dev(c, (cudaStream_t)stream, d_points, reinterpret_cast<const long long*>(d_offsets),
                            reinterpret_cast<const long long*>(h_offsets), S, d_out);
}

}  // extern "C"
