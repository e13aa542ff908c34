// selection.cu -- farthest-feature sampling and k-center greedy on sm_100a.
//
// Replaces the per-pick numpy passes of farthest_features_sample (fps_gcn_cpu.py:137-146) and
// kCenterGreedy.update_distances/select_batch_ (kcenterGreedy.py:60-128).
//
// One PERSISTENT cooperative kernel runs every pick.  Every WARP owns a contiguous run of rows and streams it from
// HBM/L2 through its own cp.async multi-stage shared-memory ring (padded rows => conflict-free LDS; no block-wide
// barrier inside a pick; the ring keeps running across picks, so the next pick's first rows are already in flight
// while the grid agrees on the next centre).  The warp evaluates the distance of its rows to the current centre in
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
#include <math.h>

#include "common.cuh"

namespace ssdr {
int nccl_allreduce_max_u64(void* comm, unsigned long long* buf, size_t count, cudaStream_t stream);  // nccl_shim.cu
namespace sel {

constexpr int THREADS = 512;          // upper bound; the host launches p.nwarps*32 threads
constexpr int WARPS = THREADS / 32;
constexpr int MAX_LEAVES = 128;
constexpr int MAX_STACK = 10;

enum Mode { MODE_FPS = 0, MODE_KCENTER = 1 };

__host__ __device__ constexpr size_t align_up_dev(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct LeafPlan {
    int n_leaves;
    unsigned short start[MAX_LEAVES];
    unsigned short len[MAX_LEAVES];
    unsigned char merges[MAX_LEAVES];  // number of (left + right) combines after finishing leaf i
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
struct __align__(32) Mailbox {
    unsigned long long w[4];  // float: w[0] = {dist bits, tag}, w[1] = {~row, tag}; double: four chunks
};
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <typename T>
__device__ __forceinline__ void mailbox_post(Mailbox* mb, const Cand& c, unsigned tag) {
    const unsigned long long t = (unsigned long long)tag;
    st_relaxed_u64(&mb->w[0], ((c.hi >> 32) << 32) | t);
    st_relaxed_u64(&mb->w[1], ((c.hi & 0xFFFFFFFFull) << 32) | t);
    if (sizeof(T) == 8) {
        st_relaxed_u64(&mb->w[2], ((c.lo >> 32) << 32) | t);
        st_relaxed_u64(&mb->w[3], ((c.lo & 0xFFFFFFFFull) << 32) | t);
    }
}
// raw chunk words of one mailbox (all loads independent, so several mailboxes can be in flight per lane)
template <typename T>
__device__ __forceinline__ void mailbox_load(const Mailbox* mb, unsigned long long (&v)[4]) {
    v[0] = ld_relaxed_u64(&mb->w[0]);
    v[1] = ld_relaxed_u64(&mb->w[1]);
    if (sizeof(T) == 8) {
        v[2] = ld_relaxed_u64(&mb->w[2]);
        v[3] = ld_relaxed_u64(&mb->w[3]);
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

template <typename T, int MODE, int DT, int GT>  // DT: compile-time D (0 = generic), GT: row groups per unit (0 = runtime)
__global__ void __launch_bounds__(THREADS, 1) select_kernel(const Params<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
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

    // shared layout: centre row (generic path) | reduction scratch | per-warp stage rings
    T* s_center = reinterpret_cast<T*>(smem_raw);
    unsigned off = (unsigned)align_up_dev((size_t)(p.D + 8) * sizeof(T), 16);
    Cand* s_red = reinterpret_cast<Cand*>(smem_raw + off);
    off += (unsigned)align_up_dev(WARPS * sizeof(Cand), 16);
    const unsigned unit_elems = (unsigned)RW * (unsigned)stride;
    T* ring = reinterpret_cast<T*>(smem_raw + off) + (size_t)warp * S * unit_elems;
    __shared__ unsigned long long s_next_center;

    // contiguous run of units for this warp
    const unsigned long long nrows = p.row_end - p.row_begin;
    const unsigned long long units_total = (nrows + RW - 1) / RW;
    const unsigned long long TW = (unsigned long long)G * NWARP;
    const unsigned long long upw = (units_total + TW - 1) / TW;
    const unsigned long long gw = (unsigned long long)blockIdx.x * NWARP + warp;
    const unsigned long long u_begin = min(gw * upw, units_total);
    const unsigned long long u_end = min(u_begin + upw, units_total);
    const int n_it = (int)(u_end - u_begin);
    const unsigned long long w_row0 = p.row_begin + u_begin * (unsigned long long)RW;  // first row of this warp
    const unsigned long long w_rows = min((unsigned long long)n_it * RW, p.row_end - min(w_row0, p.row_end));
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
            if constexpr (DT > 0 && GT > 0) {
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
        cp_async_commit();
    };

    for (int q = 0; q < S - 1; ++q) issue();  // the ring never drains between picks
    int cur_stage = 0;

    Mailbox* boxes = reinterpret_cast<Mailbox*>(p.cand);
    for (int step = p.step_begin; step < p.step_end; ++step) {
        __syncthreads();  // warp 0 has published s_next_center; the generic centre row is no longer read
        unsigned long long c_row;
        if (step < p.n_forced) c_row = (unsigned long long)p.forced[step];
        else if (step == p.step_begin) {
            Cand w;
            w.hi = p.winners[2 * (step - 1)];
            w.lo = p.winners[2 * (step - 1) + 1];
            c_row = cand_row<T>(w);
        } else
            c_row = s_next_center;
        T creg[DT > 0 ? DT / 8 : 1];
        if (DT > 0) {
#pragma unroll
            for (int k = 0; k < (DT > 0 ? DT / 8 : 1); ++k) creg[k] = __ldcg(p.F + c_row * (unsigned long long)D + 8 * k + j);
        } else {
            for (int i = tid; i < D; i += NWARP * 32) s_center[i] = p.F[c_row * (unsigned long long)D + i];
            __syncthreads();
        }
        double xx_c = 0.0;
        if (MODE == MODE_KCENTER) xx_c = p.xx[c_row];

        Cand best;
        best.hi = 0;
        best.lo = 0;
        unsigned r_in = 0;
        for (int it = 0; it < n_it; ++it, r_in += RW) {
            issue();
            cp_async_wait_dyn(S - 1);
            __syncwarp();
            const int rows = (int)min((unsigned long long)RW, w_rows - r_in);
            const unsigned long long r0 = w_row0 + r_in;
            const T* tile = ring + (unsigned)cur_stage * unit_elems;
            if (++cur_stage == S) cur_stage = 0;
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
            if (lane < rows) {
                const unsigned long long gr = r0 + lane;
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
                best = cand_max<T>(best, make_cand<T>(m, gr));
            }
            __syncwarp();  // every lane is done with this stage before a later issue() overwrites it
        }

        // ---- CTA candidate -> mailbox; every CTA polls all mailboxes and reduces them redundantly
        best = warp_max<T>(best);
        if (lane == 0) s_red[warp] = best;
        __syncthreads();
        if (warp == 0) {
            Cand c = lane < NWARP ? s_red[lane] : Cand{0, 0};
            c = warp_max<T>(c);
            const unsigned tag = (unsigned)(step + 1);
            Mailbox* row_boxes = boxes + (size_t)(step & 1) * G;
            if (lane == 0) mailbox_post<T>(row_boxes + blockIdx.x, c, tag);
            // poll all mailboxes: every lane keeps its loads in flight together and retries only the missing ones
            Cand w{0, 0};
            constexpr int KM = 5;  // 5 x 32 = 160 >= 148 CTAs
            unsigned pending = 0;
#pragma unroll
            for (int k = 0; k < KM; ++k)
                if (lane + 32 * k < G) pending |= 1u << k;
            while (pending) {
                unsigned long long v[KM][4];
#pragma unroll
                for (int k = 0; k < KM; ++k)
                    if (pending & (1u << k)) mailbox_load<T>(row_boxes + lane + 32 * k, v[k]);
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
            }
            for (int bb = lane + 32 * KM; bb < G; bb += 32) {  // larger grids (not on B200): simple spin
                Cand o;
                unsigned long long v[4];
                do {
                    mailbox_load<T>(row_boxes + bb, v);
                } while (!mailbox_decode<T>(v, tag, &o));
                w = cand_max<T>(w, o);
            }
            __syncwarp();
            w = warp_max<T>(w);
            if (lane == 0) {
                s_next_center = cand_row<T>(w);
                if (blockIdx.x == 0) {
                    p.winners[2 * step] = w.hi;
                    p.winners[2 * step + 1] = w.lo;
                    if (p.picks) p.picks[step] = (long long)cand_row<T>(w);
                }
            }
        }
        // the __syncthreads at the top of the next step publishes s_next_center
    }
    cp_async_wait<0>();
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
};

typedef cudaError_t (*AttrFn)(int);

template <typename T, int MODE, int DT, int GT>
static int setup_variant(size_t smem, int threads, void** fn_out) {
    auto kern = select_kernel<T, MODE, DT, GT>;
    SSDR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    SSDR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
    SSDR_REQUIRE(nb >= 1, SSDR_ERR_CUDA, "selection kernel does not fit on an SM (smem %zu, %d threads)", smem, threads);
    *fn_out = (void*)kern;
    return SSDR_OK;
}

// Fill geometry (stride, unit rows, stages, warps, grid) for a (T, D) problem and pick the kernel variant.
template <typename T, int MODE>
static int configure(Ctx* c, Launch<T>& L, const T* dF, size_t N, size_t D) {
    Params<T>& p = L.p;
    p.F = dF;
    p.N = N;
    p.D = (int)D;
    const int vec_full = 16 / (int)sizeof(T);
    int vec = ((D % vec_full) == 0 && ((uintptr_t)dF % 16) == 0) ? vec_full : 1;
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
    const size_t fixed = align_up_dev((D + 8) * sizeof(T), 16) + align_up_dev(WARPS * sizeof(Cand), 16) + 64;
    auto need = [&](int gr, int st, int w) { return fixed + (size_t)w * st * 4 * gr * row_bytes; };
    // fixed variants: (D, groups per unit) compiled in; units of 2-4 KB, 3 stages, 16 warps
    const bool fixed_ok = vec == vec_full && (D == 32 || D == 64 || D == 128 || D == 256);
    int groups, nst = 3, nw = WARPS;
    if (fixed_ok) groups = D == 32 ? 4 : (D == 64 ? 2 : 1);
    else {
        groups = (int)(3072 / (4 * row_bytes));
        groups = groups < 1 ? 1 : (groups > 8 ? 8 : groups);
    }
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
    if (fixed_ok && D == 32) SSDR_TRY((setup_variant<T, MODE, 32, 4>(L.smem, threads, &L.fn)));
    else if (fixed_ok && D == 64) SSDR_TRY((setup_variant<T, MODE, 64, 2>(L.smem, threads, &L.fn)));
    else if (fixed_ok && D == 128) SSDR_TRY((setup_variant<T, MODE, 128, 1>(L.smem, threads, &L.fn)));
    else if (fixed_ok && D == 256) SSDR_TRY((setup_variant<T, MODE, 256, 1>(L.smem, threads, &L.fn)));
    else SSDR_TRY((setup_variant<T, MODE, 0, 0>(L.smem, threads, &L.fn)));
    L.grid = c->sm_count;
    return SSDR_OK;
}

template <typename T, int MODE>
static int launch_steps(Launch<T>& L, int step_begin, int step_end, cudaStream_t s) {
    L.p.step_begin = step_begin;
    L.p.step_end = step_end;
    // mailbox tags are step+1 and unique per launch range, but a previous call may have left equal tags behind
    SSDR_CHECK_CUDA(cudaMemsetAsync(L.p.cand, 0, (size_t)2 * L.grid * sizeof(Mailbox), s));
    void* args[] = {(void*)&L.p};
    SSDR_CHECK_CUDA(cudaLaunchCooperativeKernel(L.fn, dim3(L.grid), dim3(L.p.nwarps * 32), args, L.smem, s));
    return SSDR_OK;
}

// Common set-up: workspaces, initial min-distance, forced centres.  d_forced may alias ws.
template <typename T, int MODE>
static int prepare(Ctx* c, Launch<T>& L, const T* dF, size_t N, size_t D, size_t row_begin, size_t row_end,
                   const long long* d_forced, int n_forced, int n_steps, cudaStream_t s) {
    SSDR_REQUIRE(N >= 1 && N < 0xFFFFFFFFull, SSDR_ERR_INVALID, "N=%zu out of range", N);
    SSDR_REQUIRE(D >= 1 && D <= 8192, SSDR_ERR_UNSUPPORTED, "feature dimension D=%zu not in [1, 8192]", D);
    SSDR_TRY((configure<T, MODE>(c, L, dF, N, D)));
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

template <typename T>
static int fps_dev(Ctx* c, const T* dF, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* d_out,
                   cudaStream_t s) {
    SSDR_REQUIRE(dF && d_out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(n_samples >= 1 && n_samples < (1ull << 31), SSDR_ERR_INVALID, "n_samples=%zu out of range", n_samples);
    SSDR_REQUIRE(first >= 0 && (size_t)first < N, SSDR_ERR_INVALID, "first index %d outside [0, %zu)", first, N);
    const int n_steps = (int)n_samples - 1;
    Launch<T> L;
    SSDR_TRY(c->ws[WS_FORCED].reserve(sizeof(long long)));
    long long f64 = first;
    SSDR_CHECK_CUDA(cudaMemcpyAsync(c->ws[WS_FORCED].p, &f64, sizeof(f64), cudaMemcpyHostToDevice, s));
    SSDR_TRY((prepare<T, MODE_FPS>(c, L, dF, N, D, 0, N, c->ws[WS_FORCED].as<long long>(), 1, n_steps, s)));
    if (n_steps > 0) SSDR_TRY((launch_steps<T, MODE_FPS>(L, 0, n_steps, s)));
    fps_emit_kernel<<<(unsigned)((n_samples + 255) / 256), 256, 0, s>>>(L.p.picks, first, d_out, n_samples);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

template <typename T>
static int kcenter_dev(Ctx* c, const T* dX, size_t N, size_t D, const int64_t* d_sel, size_t n_sel, size_t n_pick,
                       int64_t* d_out, cudaStream_t s) {
    SSDR_REQUIRE(dX && (d_out || n_pick == 0) && (d_sel || n_sel == 0), SSDR_ERR_INVALID, "NULL pointer");
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
    SSDR_TRY((prepare<T, MODE_KCENTER>(c, L, dX, N, D, 0, N, forced, n_forced, n_steps, s)));
    if (n_steps > 0) SSDR_TRY((launch_steps<T, MODE_KCENTER>(L, 0, n_steps, s)));
    kcenter_emit_kernel<<<(unsigned)((n_pick + 255) / 256), 256, 0, s>>>(L.p.picks, (long long)n_sel,
                                                                        reinterpret_cast<long long*>(d_out), n_pick);
    SSDR_CHECK_CUDA(cudaGetLastError());
    return SSDR_OK;
}

// Row-sharded FPS (one process per GPU): every rank scans rows [row_begin,row_end) of its full copy of F; after each
// pick the packed (distance bits << 32 | ~row) winners are max-all-reduced (8 bytes) so all ranks continue from the
// same centre.  One cooperative launch per pick: the collective is host-enqueued, latency bound by design.
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

// host-pointer wrappers: stage in, run, copy picks out
template <typename T>
static int fps_host(const T* F, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* out) {
    SSDR_REQUIRE(F && out, SSDR_ERR_INVALID, "NULL pointer");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_TRY(c->ws[WS_F].reserve(N * D * sizeof(T)));
    SSDR_TRY(c->ws[WS_OUT].reserve(n_samples * sizeof(int32_t)));
    SSDR_TRY(h2d(c, c->ws[WS_F].p, F, N * D * sizeof(T), c->stream));
    SSDR_TRY(fps_dev<T>(c, c->ws[WS_F].as<T>(), N, D, first, n_samples, c->ws[WS_OUT].as<int32_t>(), c->stream));
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
    SSDR_TRY(kcenter_dev<T>(c, c->ws[WS_F].as<T>(), N, D, d_sel, n_sel, n_pick, d_out, c->stream));
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
    return sel::fps_dev<float>(c, F, N, D, first, n, out, (cudaStream_t)stream);
}
int ssdr_fps_f64_dev(const double* F, size_t N, size_t D, int32_t first, size_t n, int32_t* out, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    return sel::fps_dev<double>(c, F, N, D, first, n, out, (cudaStream_t)stream);
}
int ssdr_fps_f32_sharded(const float* d_F, size_t N, size_t D, size_t row_begin, size_t row_end, int32_t first,
                         size_t n_samples, int32_t* d_out, void* nccl_comm, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
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
    return sel::kcenter_dev<float>(c, X, N, D, sel_, n_sel, n_pick, out, (cudaStream_t)stream);
}
int ssdr_kcenter_f64_dev(const double* X, size_t N, size_t D, const int64_t* sel_, size_t n_sel, size_t n_pick,
                         int64_t* out, void* stream) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    return sel::kcenter_dev<double>(c, X, N, D, sel_, n_sel, n_pick, out, (cudaStream_t)stream);
}
}
