// kdtree.cuh -- nanoflann-identical KD-tree on the device, used to resolve KNN rows whose neighbour order depends on
// nanoflann's tree-visit order (equal / almost equal fp32 distances).
//
// build_kernel reproduces KDTreeSingleIndexAdaptor::buildIndex (utils/nearest_neighbors/nanoflann.hpp:1136-1147,
// divideTree :848-896, middleSplit_ :898-937, planeSplit :948-975) bit for bit -- same vind permutation, same
// divfeat/divlow/divhigh -- but level-synchronously and data-parallel:
//   * one CTA per batch item walks the tree one LEVEL at a time over flat position arrays;
//   * a node's tight bbox (computeMinMax) is a segmented atomic min/max over its positions;
//   * each of planeSplit's two Hoare sweeps is a prefix count: the k-th misplaced element from the left swaps with
//     the k-th misplaced element from the right (validated against the sequential code in
//     tools/proto_kdtree.py and tests/test_knn_gpu.py::test_device_tree_equals_oracle_tree).
// exact_query_kernel replays findNeighbors/searchLevel (:1163-1178, :1270-1328) and KNNResultSet::addPoint
// (:72-96) with one thread per flagged query and an explicit stack.
#pragma once
#include "common.cuh"

namespace ssdr {
namespace kdtree {

constexpr int BT = 1024;
constexpr int IPT = 4;
constexpr int LEAF = 10;
constexpr int MAX_DEPTH = 96;
constexpr int MAX_K = 64;

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

struct Tree {             // all pointers are per-item base pointers; item b uses offset b*N (positions) / b*cap (nodes)
    unsigned N, cap;      // points per item, node capacity per item (2N+2)
    unsigned* vind;       // [B*N]   position -> point index (nanoflann's vind)
    unsigned* node_of;    // [B*N]   position -> node currently owning it
    unsigned* lpos;       // [B*N]   scratch: k-th misplaced position from the left, per node at [l+k]
    unsigned* rpos;       // [B*N]   scratch: k-th misplaced position from the right
    unsigned* psat;       // [B*N]   inclusive prefix count of predicate-true positions
    unsigned* pfail;      // [B*N]   inclusive prefix count of predicate-false positions
    unsigned* nl;         // [B*cap] node range [nl, nr)
    unsigned* nr;
    float* lo;            // [B*cap*3] loose bbox (root bbox cut by the ancestors' planes) -- drives middleSplit_
    float* hi;
    unsigned* tlo;        // [B*cap*3] tight bbox, order-preserving uint encoding
    unsigned* thi;
    int* c1;              // [B*cap] children (-1 = leaf)
    int* c2;
    int* feat;            // [B*cap] split dimension
    float* cutval;        // [B*cap]
    unsigned* start;      // [B*cap] first position of the current Hoare sweep
    unsigned* totsat;     // [B*cap] predicate-true count of the current sweep
    unsigned* lim1;       // [B*cap]
    unsigned char* active;  // [B*cap]
    unsigned* n_nodes;    // [B]
    const unsigned char* item_needed;  // [B] build only where a flagged row lives
};

// block-wide inclusive scan of one u64 per thread (low word / high word carry two independent counters)
__device__ __forceinline__ unsigned long long block_scan_incl(unsigned long long v, unsigned long long* s_warp,
                                                              unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < (BT / 32) ? s_warp[lane] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        s_warp[lane] = w;  // inclusive warp totals
    }
    __syncthreads();
    const unsigned long long base = warp ? s_warp[warp - 1] : 0ull;
    *total = s_warp[BT / 32 - 1];
    v += base;
    __syncthreads();  // s_warp is reused by the next call
    return v;
}

__global__ void __launch_bounds__(BT, 1) build_kernel(const float* __restrict__ pts_all, Tree t) {
    const unsigned b = blockIdx.x;
    if (t.item_needed && !t.item_needed[b]) return;
    const unsigned N = t.N, cap = t.cap;
    const float* pts = pts_all + (size_t)b * N * 3;
    unsigned* vind = t.vind + (size_t)b * N;
    unsigned* node_of = t.node_of + (size_t)b * N;
    unsigned* lpos = t.lpos + (size_t)b * N;
    unsigned* rpos = t.rpos + (size_t)b * N;
    unsigned* psat = t.psat + (size_t)b * N;
    unsigned* pfail = t.pfail + (size_t)b * N;
    unsigned* nl = t.nl + (size_t)b * cap;
    unsigned* nr = t.nr + (size_t)b * cap;
    float* lo = t.lo + (size_t)b * cap * 3;
    float* hi = t.hi + (size_t)b * cap * 3;
    unsigned* tlo = t.tlo + (size_t)b * cap * 3;
    unsigned* thi = t.thi + (size_t)b * cap * 3;
    int* c1 = t.c1 + (size_t)b * cap;
    int* c2 = t.c2 + (size_t)b * cap;
    int* feat = t.feat + (size_t)b * cap;
    float* cutval = t.cutval + (size_t)b * cap;
    unsigned* start = t.start + (size_t)b * cap;
    unsigned* totsat = t.totsat + (size_t)b * cap;
    unsigned* lim1 = t.lim1 + (size_t)b * cap;
    unsigned char* active = t.active + (size_t)b * cap;

    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned s_level_begin, s_level_end, s_next, s_nactive;
    const unsigned tid = threadIdx.x;

    for (unsigned i = tid; i < N; i += BT) {
        vind[i] = i;
        node_of[i] = 0;
    }
    if (tid == 0) {
        nl[0] = 0;
        nr[0] = N;
        c1[0] = c2[0] = -1;
        s_level_begin = 0;
        s_level_end = 1;
    }
    __syncthreads();

    for (int level = 0;; ++level) {
        const unsigned lb = s_level_begin, le = s_level_end;
        // ---- P1: tight bbox (computeMinMax) of every node of this level
        for (unsigned n = lb + tid; n < le; n += BT) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                tlo[n * 3 + d] = 0xFFFFFFFFu;
                thi[n * 3 + d] = 0u;
            }
            active[n] = 0;
        }
        if (tid == 0) {
            s_next = 0;
            s_nactive = 0;
        }
        __syncthreads();
        for (unsigned base = 0; base < N; base += BT) {
            const unsigned i = base + tid;
            const bool in = i < N;
            unsigned n = in ? node_of[i] : 0xFFFFFFFFu;
            const bool mine = in && n >= lb;
            unsigned e[3] = {0, 0, 0};
            if (mine) {
                const unsigned p = vind[i];
#pragma unroll
                for (int d = 0; d < 3; ++d) e[d] = f2ord(__ldg(pts + 3 * (size_t)p + d));
            }
            if (!mine) n = 0xFFFFFFFFu;
            // warp aggregation when the whole warp sits in one node (the common case near the root)
            const unsigned n0 = __shfl_sync(0xffffffffu, n, 0);
            if (__all_sync(0xffffffffu, n == n0)) {
                if (n0 != 0xFFFFFFFFu) {
                    unsigned mn[3] = {e[0], e[1], e[2]}, mx[3] = {e[0], e[1], e[2]};
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int m = 16; m > 0; m >>= 1) {
                            mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
                            mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
                        }
                    if ((tid & 31) == 0) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            atomicMin(&tlo[n0 * 3 + d], mn[d]);
                            atomicMax(&thi[n0 * 3 + d], mx[d]);
                        }
                    }
                }
            } else if (mine) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    atomicMin(&tlo[n * 3 + d], e[d]);
                    atomicMax(&thi[n * 3 + d], e[d]);
                }
            }
        }
        __syncthreads();
        // ---- P2: middleSplit_ decisions (nanoflann.hpp:898-929)
        for (unsigned n = lb + tid; n < le; n += BT) {
            const unsigned l = nl[n], r = nr[n];
            if (level == 0) {  // root: loose bbox = data bbox (computeBoundingBox :1241-1263)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    lo[n * 3 + d] = ord2f(tlo[n * 3 + d]);
                    hi[n * 3 + d] = ord2f(thi[n * 3 + d]);
                }
            }
            if (r - l > (unsigned)LEAF) {
                const float EPS = 0.00001f;
                float span[3], max_span;
#pragma unroll
                for (int d = 0; d < 3; ++d) span[d] = __fsub_rn(hi[n * 3 + d], lo[n * 3 + d]);
                max_span = span[0];
                if (span[1] > max_span) max_span = span[1];
                if (span[2] > max_span) max_span = span[2];
                const float thr = __fmul_rn(__fsub_rn(1.0f, EPS), max_span);
                float max_spread = -1.f;
                int cf = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (span[d] > thr) {
                        const float spread = __fsub_rn(ord2f(thi[n * 3 + d]), ord2f(tlo[n * 3 + d]));
                        if (spread > max_spread) {
                            cf = d;
                            max_spread = spread;
                        }
                    }
                }
                const float split_val = __fdiv_rn(__fadd_rn(lo[n * 3 + cf], hi[n * 3 + cf]), 2.0f);
                const float mn = ord2f(tlo[n * 3 + cf]), mx = ord2f(thi[n * 3 + cf]);
                float cv;
                if (split_val < mn) cv = mn;
                else if (split_val > mx) cv = mx;
                else cv = split_val;
                feat[n] = cf;
                cutval[n] = cv;
                start[n] = l;
                active[n] = 1;
                atomicAdd(&s_nactive, 1u);
            }
        }
        __syncthreads();
        if (s_nactive == 0) break;

        // ---- planeSplit: two Hoare sweeps as prefix counts (nanoflann.hpp:948-975)
        for (int sweep = 0; sweep < 2; ++sweep) {
            unsigned long long carry = 0;
            for (unsigned base = 0; base < N; base += BT * IPT) {
                const unsigned i0 = base + tid * IPT;
                unsigned long long f[IPT];
                unsigned long long local = 0;
#pragma unroll
                for (int k = 0; k < IPT; ++k) {
                    const unsigned i = i0 + k;
                    unsigned long long fl = 0;
                    if (i < N) {
                        const unsigned n = node_of[i];
                        if (n >= lb && active[n] && i >= start[n]) {
                            const float v = __ldg(pts + 3 * (size_t)vind[i] + feat[n]);
                            const float cv = cutval[n];
                            const bool sat = sweep == 0 ? (v < cv) : (v <= cv);
                            fl = sat ? 1ull : (1ull << 32);
                        }
                    }
                    local += fl;
                    f[k] = local;  // inclusive inside the thread
                }
                unsigned long long tot;
                const unsigned long long incl = block_scan_incl(local, s_warp, &tot);
                const unsigned long long excl = incl - local + carry;
#pragma unroll
                for (int k = 0; k < IPT; ++k) {
                    const unsigned i = i0 + k;
                    if (i < N) {
                        const unsigned long long v = excl + f[k];
                        psat[i] = (unsigned)(v & 0xFFFFFFFFull);
                        pfail[i] = (unsigned)(v >> 32);
                    }
                }
                carry += tot;
            }
            __syncthreads();
            for (unsigned n = lb + tid; n < le; n += BT) {
                if (active[n]) {
                    const unsigned s0 = start[n], r = nr[n];
                    unsigned ts = 0;
                    if (r > s0) ts = psat[r - 1] - (s0 > 0 ? psat[s0 - 1] : 0u);
                    totsat[n] = ts;
                }
            }
            __syncthreads();
            for (unsigned base = 0; base < N; base += BT) {
                const unsigned i = base + tid;
                if (i < N) {
                    const unsigned n = node_of[i];
                    if (n >= lb && active[n]) {
                        const unsigned s0 = start[n];
                        if (i >= s0) {
                            const unsigned ts = totsat[n], lim = s0 + ts, l = nl[n];
                            const unsigned bs = s0 > 0 ? psat[s0 - 1] : 0u, bf = s0 > 0 ? pfail[s0 - 1] : 0u;
                            const unsigned cs = psat[i] - bs, cfl = pfail[i] - bf;  // inclusive counts from s0
                            const bool sat = (i == 0 ? psat[0] : psat[i] - psat[i - 1]) != 0;
                            if (!sat) {
                                if (i < lim) lpos[l + (cfl - 1)] = i;
                            } else {
                                if (i >= lim) rpos[l + (ts - cs)] = i;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            for (unsigned base = 0; base < N; base += BT) {
                const unsigned i = base + tid;
                if (i < N) {
                    const unsigned n = node_of[i];
                    if (n >= lb && active[n]) {
                        const unsigned s0 = start[n], ts = totsat[n], lim = s0 + ts, l = nl[n];
                        unsigned m = 0;  // misplaced pairs = predicate-false positions in [s0, lim)
                        if (lim > s0) m = pfail[lim - 1] - (s0 > 0 ? pfail[s0 - 1] : 0u);
                        const unsigned k = i - l;
                        if (k < m) {
                            const unsigned a = lpos[l + k], bb = rpos[l + k];
                            const unsigned va = vind[a], vb = vind[bb];
                            vind[a] = vb;
                            vind[bb] = va;
                        }
                    }
                }
            }
            __syncthreads();
            for (unsigned n = lb + tid; n < le; n += BT) {
                if (active[n]) {
                    const unsigned lim = start[n] + totsat[n];
                    if (sweep == 0) {
                        lim1[n] = lim;
                        start[n] = lim;
                    } else {
                        // ---- split index (middleSplit_ :934-936) and children (divideTree :877-892)
                        const unsigned l = nl[n], r = nr[n], count = r - l;
                        const unsigned l1 = lim1[n] - l, l2 = lim - l;
                        unsigned idx;
                        if (l1 > count / 2) idx = l1;
                        else if (l2 < count / 2) idx = l2;
                        else idx = count / 2;
                        const unsigned off = atomicAdd(&s_next, 2u);
                        const unsigned a = le + off, bb = a + 1;
                        const int cf = feat[n];
                        const float cv = cutval[n];
                        nl[a] = l;
                        nr[a] = l + idx;
                        nl[bb] = l + idx;
                        nr[bb] = r;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const float lv = lo[n * 3 + d], hv = hi[n * 3 + d];
                            lo[a * 3 + d] = lv;
                            hi[a * 3 + d] = (d == cf) ? cv : hv;
                            lo[bb * 3 + d] = (d == cf) ? cv : lv;
                            hi[bb * 3 + d] = hv;
                        }
                        c1[a] = c2[a] = c1[bb] = c2[bb] = -1;
                        c1[n] = (int)a;
                        c2[n] = (int)bb;
                    }
                }
            }
            __syncthreads();
        }
        // ---- positions move to their child node
        for (unsigned base = 0; base < N; base += BT) {
            const unsigned i = base + tid;
            if (i < N) {
                const unsigned n = node_of[i];
                if (n >= lb && active[n]) {
                    const unsigned a = (unsigned)c1[n];
                    node_of[i] = i < nr[a] ? a : (unsigned)c2[n];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_level_begin = le;
            s_level_end = le + s_next;
        }
        __syncthreads();
    }
    if (tid == 0) t.n_nodes[b] = s_level_end;
}

__global__ void mark_items_kernel(const unsigned* __restrict__ flag_list, unsigned n_flag, unsigned Q,
                                  unsigned char* __restrict__ item_needed) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_flag) item_needed[flag_list[i] / Q] = 1;
}

// ---- exact replay of nanoflann's search for the flagged rows -------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(64) exact_query_kernel(const float* __restrict__ pts_all,
                                                         const float* __restrict__ q_all, Tree t, unsigned Q, int K,
                                                         const unsigned* __restrict__ flag_list, unsigned n_flag,
                                                         OutT* __restrict__ out, unsigned* __restrict__ overflow) {
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_flag) return;
    const unsigned row = flag_list[f];
    const unsigned b = row / Q;
    const unsigned N = t.N, cap = t.cap;
    const float* pts = pts_all + (size_t)b * N * 3;
    const unsigned* vind = t.vind + (size_t)b * N;
    const unsigned* nl = t.nl + (size_t)b * cap;
    const unsigned* nr = t.nr + (size_t)b * cap;
    const unsigned* tlo = t.tlo + (size_t)b * cap * 3;
    const unsigned* thi = t.thi + (size_t)b * cap * 3;
    const int* c1 = t.c1 + (size_t)b * cap;
    const int* c2 = t.c2 + (size_t)b * cap;
    const int* feat = t.feat + (size_t)b * cap;
    const float q[3] = {q_all[3 * (size_t)row], q_all[3 * (size_t)row + 1], q_all[3 * (size_t)row + 2]};

    float rd[MAX_K];
    unsigned ri[MAX_K];
    int count = 0;
    rd[K - 1] = 3.402823466e+38f;  // KNNResultSet::init (:47-53)

    // computeInitialDistances (:977-995) against the root bbox
    float d0[3] = {0.f, 0.f, 0.f};
    float distsq = 0.f;
    for (int d = 0; d < 3; ++d) {
        const float blo = ord2f(tlo[d]), bhi = ord2f(thi[d]);
        if (q[d] < blo) {
            const float df = __fsub_rn(q[d], blo);
            d0[d] = __fmul_rn(df, df);
            distsq = __fadd_rn(distsq, d0[d]);
        }
        if (q[d] > bhi) {
            const float df = __fsub_rn(q[d], bhi);
            d0[d] = __fmul_rn(df, df);
            distsq = __fadd_rn(distsq, d0[d]);
        }
    }
    int st_node[MAX_DEPTH];
    float st_min[MAX_DEPTH], st_d[MAX_DEPTH][3];
    int sp = 0;
    st_node[0] = 0;
    st_min[0] = distsq;
    st_d[0][0] = d0[0];
    st_d[0][1] = d0[1];
    st_d[0][2] = d0[2];
    sp = 1;
    bool first = true;
    while (sp > 0) {
        --sp;
        int node = st_node[sp];
        const float mind = st_min[sp];
        float dd[3] = {st_d[sp][0], st_d[sp][1], st_d[sp][2]};
        if (!first && !(mind <= rd[K - 1])) continue;  // mindistsq*epsError <= worstDist()  (:1319)
        first = false;
        while (c1[node] >= 0) {
            const int idx = feat[node];
            const float val = q[idx];
            const int a = c1[node], bb = c2[node];
            const float divlow = ord2f(thi[a * 3 + idx]);
            const float divhigh = ord2f(tlo[bb * 3 + idx]);
            const float diff1 = __fsub_rn(val, divlow), diff2 = __fsub_rn(val, divhigh);
            int best, other;
            float cut;
            if (__fadd_rn(diff1, diff2) < 0.f) {
                best = a;
                other = bb;
                cut = __fmul_rn(diff2, diff2);
            } else {
                best = bb;
                other = a;
                cut = __fmul_rn(diff1, diff1);
            }
            if (sp < MAX_DEPTH) {
                st_node[sp] = other;
                st_min[sp] = __fsub_rn(__fadd_rn(mind, cut), dd[idx]);
                st_d[sp][0] = idx == 0 ? cut : dd[0];
                st_d[sp][1] = idx == 1 ? cut : dd[1];
                st_d[sp][2] = idx == 2 ? cut : dd[2];
                ++sp;
            } else {
                *overflow = 1u;  // tree deeper than MAX_DEPTH: reported to the host, never silently wrong
            }
            node = best;
        }
        const float worst = rd[K - 1];  // snapshot once per leaf (:1277)
        for (unsigned i = nl[node]; i < nr[node]; ++i) {
            const unsigned index = vind[i];
            float dist = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float df = __fsub_rn(q[d], __ldg(pts + 3 * (size_t)index + d));
                dist = __fadd_rn(dist, __fmul_rn(df, df));
            }
            if (dist < worst) {  // KNNResultSet::addPoint (:72-96), strict '>' shifting
                int j;
                for (j = count; j > 0; --j) {
                    if (rd[j - 1] > dist) {
                        if (j < K) {
                            rd[j] = rd[j - 1];
                            ri[j] = ri[j - 1];
                        }
                    } else
                        break;
                }
                if (j < K) {
                    rd[j] = dist;
                    ri[j] = index;
                }
                if (count < K) ++count;
            }
        }
    }
    OutT* o = out + (size_t)row * K;
    for (int j = 0; j < count; ++j) o[j] = (OutT)ri[j];
}

enum { TW_BASE = 16 };  // workspace slots TW_BASE.. are owned by this header

// Carve the tree arrays out of three workspace slabs.
static int alloc_tree(Ctx* c, size_t B, size_t N, Tree* out) {
    const size_t cap = 2 * N + 2;
    const size_t posb = B * N * sizeof(unsigned), nodeb = B * cap * sizeof(unsigned);
    SSDR_TRY(c->ws[TW_BASE + 0].reserve(6 * posb));
    SSDR_TRY(c->ws[TW_BASE + 1].reserve(nodeb * 21 + B * cap + 64));
    SSDR_TRY(c->ws[TW_BASE + 2].reserve(B * sizeof(unsigned) + B + 72));
    Tree t;
    t.N = (unsigned)N;
    t.cap = (unsigned)cap;
    unsigned* pb = c->ws[TW_BASE + 0].as<unsigned>();
    t.vind = pb;
    t.node_of = pb + B * N;
    t.lpos = pb + 2 * B * N;
    t.rpos = pb + 3 * B * N;
    t.psat = pb + 4 * B * N;
    t.pfail = pb + 5 * B * N;
    unsigned* nb = c->ws[TW_BASE + 1].as<unsigned>();
    const size_t u = B * cap;
    t.nl = nb;
    t.nr = nb + u;
    t.lo = reinterpret_cast<float*>(nb + 2 * u);
    t.hi = reinterpret_cast<float*>(nb + 5 * u);
    t.tlo = nb + 8 * u;
    t.thi = nb + 11 * u;
    t.c1 = reinterpret_cast<int*>(nb + 14 * u);
    t.c2 = reinterpret_cast<int*>(nb + 15 * u);
    t.feat = reinterpret_cast<int*>(nb + 16 * u);
    t.cutval = reinterpret_cast<float*>(nb + 17 * u);
    t.start = nb + 18 * u;
    t.totsat = nb + 19 * u;
    t.lim1 = nb + 20 * u;
    t.active = reinterpret_cast<unsigned char*>(nb + 21 * u);
    t.n_nodes = c->ws[TW_BASE + 2].as<unsigned>();
    t.item_needed = nullptr;
    *out = t;
    return SSDR_OK;
}

// Build the trees of the items that own flagged rows, then overwrite those rows with nanoflann's exact answer.
template <typename OutT>
static int resolve_flagged(Ctx* c, cudaStream_t s, const float* d_pts, size_t B, size_t N, const float* d_q, size_t Q,
                           size_t K, OutT* d_out, const unsigned* flag_list, unsigned n_flag,
                           unsigned long long* builds) {
    SSDR_REQUIRE(K <= (size_t)MAX_K, SSDR_ERR_UNSUPPORTED, "K=%zu > %d in the exact tie path", K, MAX_K);
    Tree t;
    SSDR_TRY(alloc_tree(c, B, N, &t));
    unsigned char* needed = reinterpret_cast<unsigned char*>(t.n_nodes + B);
    t.item_needed = needed;
    unsigned* overflow = reinterpret_cast<unsigned*>(needed + ((B + 3) / 4) * 4);
    SSDR_CHECK_CUDA(cudaMemsetAsync(needed, 0, ((B + 3) / 4) * 4 + 4, s));
    mark_items_kernel<<<(n_flag + 255) / 256, 256, 0, s>>>(flag_list, n_flag, (unsigned)Q, needed);
    build_kernel<<<(unsigned)B, BT, 0, s>>>(d_pts, t);
    exact_query_kernel<OutT><<<(n_flag + 63) / 64, 64, 0, s>>>(d_pts, d_q, t, (unsigned)Q, (int)K, flag_list, n_flag,
                                                              d_out, overflow);
    SSDR_CHECK_CUDA(cudaGetLastError());
    unsigned h_over = 0;
    SSDR_TRY(d2h_sync(c, &h_over, overflow, sizeof(unsigned), s));
    SSDR_REQUIRE(h_over == 0, SSDR_ERR_UNSUPPORTED, "KD-tree deeper than %d levels in the exact tie path", MAX_DEPTH);
    if (builds) *builds = B;  // upper bound; items without flagged rows exit immediately
    return SSDR_OK;
}

}  // namespace kdtree
}  // namespace ssdr
