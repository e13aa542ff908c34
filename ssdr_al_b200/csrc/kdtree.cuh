// kdtree.cuh -- nanoflann-identical KD-tree on the device, used to resolve KNN rows whose neighbour order depends on
// nanoflann's tree-visit order (equal / almost equal fp32 distances).
//
// build_kernel reproduces KDTreeSingleIndexAdaptor::buildIndex (utils/nearest_neighbors/nanoflann.hpp:1136-1147,
// divideTree :848-896, middleSplit_ :898-937, planeSplit :948-975) bit for bit -- same vind permutation, same
// divfeat/divlow/divhigh -- as ONE persistent cooperative launch for all batch items.  The working array is
// position-ordered float4 (x, y, z, point index), i.e. nanoflann's vind with the coordinates carried along, so
// every pass is a coalesced stream and the later search reads a leaf with one load per point:
//   * TOP: nodes with more than MED_MAX points are split level by level from global memory, one CTA per node
//     (block-wide prefix counts), a grid barrier between levels;
//   * SUBTREES: every node of <= MED_MAX points is handed to one CTA that loads it into shared memory ONCE and grows
//     its whole subtree there -- CTA-wide splits for nodes > SMALL_MAX, one warp per node below (ballot/popc prefix
//     counts) -- with block barriers only, then writes the permuted points back;
//   * computeMinMax is a warp/block min-max reduction; each of planeSplit's two Hoare sweeps is a prefix count: the
//     k-th misplaced element from the left swaps with the k-th misplaced element from the right (SURVEY.md A.5;
//     checked against the sequential code in tools/proto_kdtree.py and
//     tests/test_knn_gpu.py::test_device_tree_equals_sequential_tree).
// exact_query_kernel replays findNeighbors/searchLevel (:1163-1178, :1270-1328) and KNNResultSet::addPoint
// (:72-96) with one thread per flagged query, an explicit stack and one 32-byte node record per visit.
#pragma once
#include "common.cuh"

namespace ssdr {
namespace kdtree {

constexpr int BT = 1024;
constexpr int NW = BT / 32;
constexpr int IPT = 4;          // subtree CTA-wide scan: MED_MAX == BT * IPT
constexpr int IPT_BIG = 4;      // top-level scan chunk per thread (8 and loop unrolling measured slower: spills)
constexpr int LEAF = 10;
constexpr int SMALL_MAX = 1024;   // split by one warp (32 slots per lane)
constexpr int MED_MAX = 4096;     // whole subtree grown in the shared memory of one CTA
constexpr int LIST_MED = 4;       // > SMALL_MAX nodes inside one subtree level (<= MED_MAX / (SMALL_MAX+1))
constexpr int LIST_SMALL = 384;   // split-able nodes inside one subtree level (<= MED_MAX / (LEAF+1))
constexpr int MAX_LEVELS = 512;
constexpr int MAX_GROUP_LEVELS = 64;  // levels that may use several CTAs per node (node size halves per level)
constexpr int MAX_DEPTH = 96;
constexpr int MAX_K = 64;

struct __align__(16) NodeRec {
    int c1, c2;            // children (per-item node ids), -1 = leaf          nanoflann.hpp:853-856
    int feat;              // split dimension                                   :877
    unsigned pad;
    float divlow, divhigh; // tight max of left child / min of right child      :886-887
    unsigned l, r;         // position range [l, r) in pp
};

struct __align__(16) Entry {  // a node waiting to be split inside a subtree (positions relative to the subtree)
    unsigned gid;
    unsigned short l, r;
    float lo[3], hi[3];        // loose bbox (drives middleSplit_)
};

// dynamic shared memory of build_kernel (subtree phase)
constexpr size_t SM_PP = 0;
constexpr size_t SM_PSAT = SM_PP + (size_t)MED_MAX * 16;
constexpr size_t SM_PFAIL = SM_PSAT + (size_t)MED_MAX * 2;
constexpr size_t SM_LPOS = SM_PFAIL + (size_t)MED_MAX * 2;
constexpr size_t SM_RPOS = SM_LPOS + (size_t)MED_MAX;
constexpr size_t SM_WSCR = SM_RPOS + (size_t)MED_MAX;
constexpr size_t SM_LISTS = SM_WSCR + (size_t)NW * SMALL_MAX * 2;  // per warp: sL, sR (u16 x SMALL_MAX/2 each)
constexpr size_t SM_TOTAL = SM_LISTS + 2 * (size_t)(LIST_MED + LIST_SMALL) * sizeof(Entry);

struct Tree {
    unsigned N, cap, B;    // points per item, node capacity per item, items
    unsigned lcap;         // capacity of one top-level work list / of the subtree list
    float4* pp;            // [B*N]   position-ordered points: (x, y, z, index) -- nanoflann's vind + coordinates
    unsigned* lpos;        // [B*N]   top-level scratch: k-th misplaced position from the left, at [l+k]
    unsigned* rpos;        // [B*N]   ... from the right
    unsigned* psat;        // [B*N]   inclusive prefix count of predicate-true positions (from the sweep start)
    unsigned* pfail;       // [B*N]   inclusive prefix count of predicate-false positions
    NodeRec* nodes;        // [B*cap]
    float* nlo;            // [B*cap*3] loose bbox of top-level nodes and subtree roots
    float* nhi;
    float* root_lo;        // [B*3] tight root bbox (computeBoundingBox :1241-1263)
    float* root_hi;
    unsigned* node_count;  // [B]
    unsigned* list;        // [2 parities][lcap] top-level nodes (global ids: item*cap + node)
    unsigned* sublist;     // [lcap] subtree roots
    unsigned* list_cnt;    // [MAX_LEVELS+1] top-level counts; [MAX_LEVELS+1] = subtree count
    unsigned* barrier;     // grid barrier counter
    unsigned* gbar;        // [MAX_GROUP_LEVELS*G] group barrier counters (one per level and node slot), zeroed per build
    unsigned* gred;        // [MAX_GROUP_LEVELS*G*8] group reductions: ~min xyz, max xyz, max(left cut coord), ~min(right)
    unsigned* gpart;       // [G*2*G*2] per node slot, per sweep, per member CTA: (true count, false count)
    unsigned* root_red;    // [B*8] root bbox reduction (order-preserving atomics, identity 0) + ticket, zeroed per build
    const unsigned* n_flag; // device count of flagged rows (nullptr = build unconditionally)
    unsigned* error;       // bit 0: node capacity, bit 1: level cap, bit 2: DFS stack, bit 3: list capacity
    const unsigned char* item_needed;  // [B] build only where a flagged row lives (nullptr = all)
    unsigned long long* tstamps;       // [16] optional %globaltimer marks (diagnostics; nullptr = off)
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
__device__ __forceinline__ void mark(const unsigned long long* dummy, unsigned long long* ts, int slot) {
    (void)dummy;
    if (ts && blockIdx.x == 0 && threadIdx.x == 0) ts[slot] = gtimer();
}
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_rel_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& phase) {
    __syncthreads();
    ++phase;
    if (threadIdx.x == 0) {
        __threadfence();
        red_rel_add(counter, 1u);
        const unsigned target = phase * gridDim.x;
        while (ld_acq(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// block-wide inclusive scan of one u64 per thread (low / high word carry two independent counters)
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
        unsigned long long w = lane < NW ? s_warp[lane] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned long long base = warp ? s_warp[warp - 1] : 0ull;
    *total = s_warp[NW - 1];
    v += base;
    __syncthreads();
    return v;
}

// middleSplit_ (:898-929): split dimension and cut value from the loose bbox and the node's tight min/max
__device__ __forceinline__ void decide_split(const float lo[3], const float hi[3], const float mn[3], const float mx[3],
                                             int* cf_out, float* cv_out) {
    const float EPS = 0.00001f;
    float span[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) span[d] = __fsub_rn(hi[d], lo[d]);
    float max_span = span[0];
    if (span[1] > max_span) max_span = span[1];
    if (span[2] > max_span) max_span = span[2];
    const float thr = __fmul_rn(__fsub_rn(1.0f, EPS), max_span);
    float max_spread = -1.f;
    int cf = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (span[d] > thr) {
            const float spread = __fsub_rn(mx[d], mn[d]);
            if (spread > max_spread) {
                cf = d;
                max_spread = spread;
            }
        }
    }
    const float lo_c = cf == 0 ? lo[0] : (cf == 1 ? lo[1] : lo[2]);
    const float hi_c = cf == 0 ? hi[0] : (cf == 1 ? hi[1] : hi[2]);
    const float mn_c = cf == 0 ? mn[0] : (cf == 1 ? mn[1] : mn[2]);
    const float mx_c = cf == 0 ? mx[0] : (cf == 1 ? mx[1] : mx[2]);
    const float split_val = __fdiv_rn(__fadd_rn(lo_c, hi_c), 2.0f);
    float cv;
    if (split_val < mn_c) cv = mn_c;
    else if (split_val > mx_c) cv = mx_c;
    else cv = split_val;
    *cf_out = cf;
    *cv_out = cv;
}

__device__ __forceinline__ float comp(const float4& v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }

__device__ __forceinline__ void store_split(const Tree& t, unsigned g, unsigned ga, unsigned b, unsigned l, unsigned r,
                                            unsigned idx, int cf, float divlow, float divhigh) {
    // ga = global id of the left child (right child = ga + 1); per-item ids go into the parent's record
    NodeRec ra, rb, me;
    ra.c1 = ra.c2 = rb.c1 = rb.c2 = -1;
    ra.feat = rb.feat = 0;
    ra.pad = rb.pad = me.pad = 0;
    ra.divlow = ra.divhigh = rb.divlow = rb.divhigh = 0.f;
    ra.l = l;
    ra.r = l + idx;
    rb.l = l + idx;
    rb.r = r;
    t.nodes[ga] = ra;
    t.nodes[ga + 1] = rb;
    me.c1 = (int)(ga - b * t.cap);
    me.c2 = me.c1 + 1;
    me.feat = cf;
    me.divlow = divlow;
    me.divhigh = divhigh;
    me.l = l;
    me.r = r;
    t.nodes[g] = me;
}

// divideTree's bookkeeping for one TOP-LEVEL split (:877-892): children records, loose bboxes, work lists
__device__ __forceinline__ void emit_children_top(const Tree& t, unsigned g, unsigned b, unsigned l, unsigned r,
                                                  unsigned idx, int cf, float cv, float divlow, float divhigh,
                                                  const float lo[3], const float hi[3], int level) {
    const unsigned a = atomicAdd(&t.node_count[b], 2u);
    if (a + 1 >= t.cap) {
        atomicOr(t.error, 1u);
        return;
    }
    const unsigned ga = b * t.cap + a, gb = ga + 1;
    store_split(t, g, ga, b, l, r, idx, cf, divlow, divhigh);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        t.nlo[(size_t)ga * 3 + d] = lo[d];
        t.nhi[(size_t)ga * 3 + d] = (d == cf) ? cv : hi[d];
        t.nlo[(size_t)gb * 3 + d] = (d == cf) ? cv : lo[d];
        t.nhi[(size_t)gb * 3 + d] = hi[d];
    }
    if (level + 1 >= MAX_LEVELS) {
        atomicOr(t.error, 2u);
        return;
    }
    unsigned* next = t.list + (size_t)((level + 1) & 1) * t.lcap;
    const unsigned cnt[2] = {idx, (r - l) - idx};
    const unsigned gid[2] = {ga, gb};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (cnt[k] > (unsigned)MED_MAX) {
            const unsigned pos = atomicAdd(&t.list_cnt[level + 1], 1u);
            if (pos < t.lcap) next[pos] = gid[k];
            else atomicOr(t.error, 8u);
        } else if (cnt[k] > (unsigned)LEAF) {
            const unsigned pos = atomicAdd(&t.list_cnt[MAX_LEVELS + 1], 1u);
            if (pos < t.lcap) t.sublist[pos] = gid[k];
            else atomicOr(t.error, 8u);
        }
    }
}

// ---- TOP: one CTA splits one node of > MED_MAX points over global memory ----------------------------------------
__device__ __forceinline__ void split_big(const Tree& t, unsigned g, int level, unsigned long long* s_warp,
                                          float* s_red, float* s_bc) {
    const unsigned tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned b = g / t.cap;
    float4* pp = t.pp + (size_t)b * t.N;
    unsigned* lpos = t.lpos + (size_t)b * t.N;
    unsigned* rpos = t.rpos + (size_t)b * t.N;
    unsigned* psat = t.psat + (size_t)b * t.N;
    unsigned* pfail = t.pfail + (size_t)b * t.N;
    __syncthreads();  // smem broadcast slots may still be read by a previous node's stragglers
    const unsigned l = __ldcg(&t.nodes[g].l), r = __ldcg(&t.nodes[g].r);
    const unsigned count = r - l;
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = __ldcg(&t.nlo[(size_t)g * 3 + d]);
        hi[d] = __ldcg(&t.nhi[(size_t)g * 3 + d]);
    }
    // computeMinMax (:827-836)
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = l + tid; i < r; i += BT) {
        const float4 v = __ldcg(pp + i);
        mn[0] = fminf(mn[0], v.x);
        mx[0] = fmaxf(mx[0], v.x);
        mn[1] = fminf(mn[1], v.y);
        mx[1] = fmaxf(mx[1], v.y);
        mn[2] = fminf(mn[2], v.z);
        mx[2] = fmaxf(mx[2], v.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
        }
        if (lane == 0) {
            s_red[d * NW + warp] = mn[d];
            s_red[(3 + d) * NW + warp] = mx[d];
        }
    }
    __syncthreads();
    if (tid == 0) {
        float amn[3], amx[3];
        for (int d = 0; d < 3; ++d) {
            amn[d] = s_red[d * NW];
            amx[d] = s_red[(3 + d) * NW];
            for (int w = 1; w < NW; ++w) {
                amn[d] = fminf(amn[d], s_red[d * NW + w]);
                amx[d] = fmaxf(amx[d], s_red[(3 + d) * NW + w]);
            }
        }
        int cf;
        float cv;
        decide_split(lo, hi, amn, amx, &cf, &cv);
        s_bc[0] = __int_as_float(cf);
        s_bc[1] = cv;
    }
    __syncthreads();
    const int cf = __float_as_int(s_bc[0]);
    const float cv = s_bc[1];

    // planeSplit (:948-975)
    unsigned start = l, lim1 = l, lim2 = l;
    for (int sweep = 0; sweep < 2; ++sweep) {
        unsigned long long carry = 0;
        for (unsigned base = start; base < r; base += BT * IPT_BIG) {
            const unsigned i0 = base + tid * IPT_BIG;
            unsigned long long f[IPT_BIG];
            float vals[IPT_BIG];
#pragma unroll
            for (int k = 0; k < IPT_BIG; ++k) vals[k] = (i0 + k < r) ? comp(__ldcg(pp + i0 + k), cf) : 0.f;
            unsigned long long local = 0;
#pragma unroll
            for (int k = 0; k < IPT_BIG; ++k) {
                const unsigned i = i0 + k;
                unsigned long long fl = 0;
                if (i < r) {
                    const float v = vals[k];
                    const bool sat = sweep == 0 ? (v < cv) : (v <= cv);
                    fl = sat ? 1ull : (1ull << 32);
                }
                local += fl;
                f[k] = local;
            }
            unsigned long long tot;
            const unsigned long long incl = block_scan_incl(local, s_warp, &tot);
            const unsigned long long excl = incl - local + carry;
#pragma unroll
            for (int k = 0; k < IPT_BIG; ++k) {
                const unsigned i = i0 + k;
                if (i < r) {
                    const unsigned long long v = excl + f[k];
                    psat[i] = (unsigned)(v & 0xFFFFFFFFull);
                    pfail[i] = (unsigned)(v >> 32);
                }
            }
            carry += tot;
        }
        __syncthreads();
        const unsigned tot_sat = r > start ? psat[r - 1] : 0u;
        const unsigned lim = start + tot_sat;
        const unsigned m = lim > start ? pfail[lim - 1] : 0u;  // misplaced pairs
        for (unsigned i = start + tid; i < r; i += BT) {
            const unsigned cs = psat[i], cfl = pfail[i];
            const bool sat = (i == start ? cs : cs - psat[i - 1]) != 0;
            if (!sat) {
                if (i < lim) lpos[l + (cfl - 1)] = i;
            } else {
                if (i >= lim) rpos[l + (tot_sat - cs)] = i;
            }
        }
        __syncthreads();
        for (unsigned k = tid; k < m; k += BT) {
            const unsigned a = lpos[l + k], c = rpos[l + k];
            const float4 va = __ldcg(pp + a), vc = __ldcg(pp + c);
            pp[a] = vc;
            pp[c] = va;
        }
        __syncthreads();
        if (sweep == 0) {
            lim1 = lim;
            start = lim;
        } else {
            lim2 = lim;
        }
    }
    const unsigned l1 = lim1 - l, l2 = lim2 - l;
    unsigned idx;  // :934-936
    if (l1 > count / 2) idx = l1;
    else if (l2 < count / 2) idx = l2;
    else idx = count / 2;
    float dlow = -INFINITY, dhigh = INFINITY;
    for (unsigned i = l + tid; i < r; i += BT) {
        const float v = comp(__ldcg(pp + i), cf);
        if (i - l < idx) dlow = fmaxf(dlow, v);
        else dhigh = fminf(dhigh, v);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        dlow = fmaxf(dlow, __shfl_xor_sync(0xffffffffu, dlow, m));
        dhigh = fminf(dhigh, __shfl_xor_sync(0xffffffffu, dhigh, m));
    }
    if (lane == 0) {
        s_red[warp] = dlow;
        s_red[NW + warp] = dhigh;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NW; ++w) {
            dlow = fmaxf(dlow, s_red[w]);
            dhigh = fminf(dhigh, s_red[NW + w]);
        }
        emit_children_top(t, g, b, l, r, idx, cf, cv, dlow, dhigh, lo, hi, level);
    }
}

// ---- TOP, group mode: when a level has fewer big nodes than CTAs, k = gridDim / nodes CTAs share one node -----------
// Each member owns a contiguous slice of the node's positions; cross-CTA prefix counts go through per-member totals,
// and the members meet at a group barrier (a counter private to this node and level) between the phases.
__device__ __forceinline__ void group_sync(unsigned* counter, unsigned k, unsigned& phase) {
    __syncthreads();
    ++phase;
    if (threadIdx.x == 0) {
        // release/acquire on the counter order every member's writes: the release is cumulative over what the
        // bar.sync above made visible to this thread, the acquire is published to the CTA by the bar.sync below
        red_rel_add(counter, 1u);
        const unsigned target = phase * k;
        while (ld_acq(counter) < target) {
        }
    }
    __syncthreads();
}
__device__ __forceinline__ unsigned f2ord_u(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f_u(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__device__ __forceinline__ void split_big_group(const Tree& t, unsigned g, int level, unsigned slot, unsigned k,
                                                unsigned s, unsigned long long* s_warp, float* s_red) {
    const unsigned tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned G = gridDim.x;
    const unsigned b = g / t.cap;
    float4* pp = t.pp + (size_t)b * t.N;
    unsigned* lpos = t.lpos + (size_t)b * t.N;
    unsigned* rpos = t.rpos + (size_t)b * t.N;
    unsigned* psat = t.psat + (size_t)b * t.N;
    unsigned* pfail = t.pfail + (size_t)b * t.N;
    unsigned* bar = t.gbar + (size_t)level * G + slot;
    unsigned* red = t.gred + ((size_t)level * G + slot) * 8;
    unsigned phase = 0;
    const unsigned l = __ldcg(&t.nodes[g].l), r = __ldcg(&t.nodes[g].r);
    const unsigned count = r - l;
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = __ldcg(&t.nlo[(size_t)g * 3 + d]);
        hi[d] = __ldcg(&t.nhi[(size_t)g * 3 + d]);
    }
    // ---- M: computeMinMax over the member's slice, combined with order-preserving atomics (identity 0)
    {
        const unsigned len = (count + k - 1) / k;
        const unsigned a = min(l + s * len, r), e = min(a + len, r);
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (unsigned i = a + tid; i < e; i += BT) {
            const float4 v = __ldcg(pp + i);
            mn[0] = fminf(mn[0], v.x);
            mx[0] = fmaxf(mx[0], v.x);
            mn[1] = fminf(mn[1], v.y);
            mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z);
            mx[2] = fmaxf(mx[2], v.z);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
                mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
            }
            if (lane == 0) {
                s_red[d * NW + warp] = mn[d];
                s_red[(3 + d) * NW + warp] = mx[d];
            }
        }
        __syncthreads();
        if (tid < 6) {  // one atomic per CTA and quantity; +-inf (empty slice) never reaches the atomics
            float v = s_red[tid * NW];
            for (int w = 1; w < NW; ++w) v = tid < 3 ? fminf(v, s_red[tid * NW + w]) : fmaxf(v, s_red[tid * NW + w]);
            if (tid < 3 && v != INFINITY) atomicMax(&red[tid], ~f2ord_u(v));
            if (tid >= 3 && v != -INFINITY) atomicMax(&red[tid], f2ord_u(v));
        }
    }
    group_sync(bar, k, phase);
    float tmn[3], tmx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        tmn[d] = ord2f_u(~__ldcg(&red[d]));
        tmx[d] = ord2f_u(__ldcg(&red[3 + d]));
    }
    int cf;
    float cv;
    decide_split(lo, hi, tmn, tmx, &cf, &cv);

    // ---- planeSplit (:948-975), two sweeps
    unsigned start = l, lim1 = l, lim2 = l;
    for (int sweep = 0; sweep < 2; ++sweep) {
        unsigned* part = t.gpart + (((size_t)slot * 2 + sweep) * G) * 2;
        const unsigned cnt = r - start;
        const unsigned len = (cnt + k - 1) / k;
        const unsigned a = min(start + s * len, r), e = min(a + len, r);
        // A1: local inclusive prefix counts over [a, e)
        unsigned long long carry = 0;
        for (unsigned base = a; base < e; base += BT * IPT_BIG) {
            const unsigned i0 = base + tid * IPT_BIG;
            unsigned long long f[IPT_BIG];
            unsigned long long local = 0;
#pragma unroll
            for (int q = 0; q < IPT_BIG; ++q) {
                const unsigned i = i0 + q;
                unsigned long long fl = 0;
                if (i < e) {
                    const float v = comp(__ldcg(pp + i), cf);
                    const bool sat = sweep == 0 ? (v < cv) : (v <= cv);
                    fl = sat ? 1ull : (1ull << 32);
                }
                local += fl;
                f[q] = local;
            }
            unsigned long long tot;
            const unsigned long long incl = block_scan_incl(local, s_warp, &tot);
            const unsigned long long excl = incl - local + carry;
#pragma unroll
            for (int q = 0; q < IPT_BIG; ++q) {
                const unsigned i = i0 + q;
                if (i < e) {
                    const unsigned long long v = excl + f[q];
                    psat[i] = (unsigned)(v & 0xFFFFFFFFull);
                    pfail[i] = (unsigned)(v >> 32);
                }
            }
            carry += tot;
        }
        if (tid == 0) {
            part[2 * s] = (unsigned)(carry & 0xFFFFFFFFull);
            part[2 * s + 1] = (unsigned)(carry >> 32);
        }
        group_sync(bar, k, phase);
        // A2: global ranks = member base + local counts (warp 0 sums the member totals, smem broadcasts them)
        if (warp == 0) {
            unsigned bs = 0, bf = 0, ts = 0, mm = 0;
            const unsigned lim_guess_len = len;
            for (unsigned q = lane; q < k; q += 32) {
                const unsigned ps = __ldcg(&part[2 * q]), pf = __ldcg(&part[2 * q + 1]);
                if (q < s) {
                    bs += ps;
                    bf += pf;
                }
                ts += ps;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                bs += __shfl_xor_sync(0xffffffffu, bs, o);
                bf += __shfl_xor_sync(0xffffffffu, bf, o);
                ts += __shfl_xor_sync(0xffffffffu, ts, o);
            }
            const unsigned lim_w = start + ts;
            if (lim_w > start) {  // misplaced pairs = predicate-false positions in [start, lim)
                const unsigned sq = (lim_w - 1 - start) / lim_guess_len;  // member that owns position lim-1
                for (unsigned q = lane; q < sq; q += 32) mm += __ldcg(&part[2 * q + 1]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mm += __shfl_xor_sync(0xffffffffu, mm, o);
                mm += __ldcg(&pfail[lim_w - 1]);
            }
            if (lane == 0) {
                s_red[0] = __uint_as_float(bs);
                s_red[1] = __uint_as_float(bf);
                s_red[2] = __uint_as_float(ts);
                s_red[3] = __uint_as_float(mm);
            }
        }
        __syncthreads();
        const unsigned base_sat = __float_as_uint(s_red[0]), base_fail = __float_as_uint(s_red[1]);
        const unsigned tot_sat = __float_as_uint(s_red[2]), m = __float_as_uint(s_red[3]);
        const unsigned lim = start + tot_sat;
        for (unsigned i = a + tid; i < e; i += BT) {
            const unsigned ls = psat[i], lf = pfail[i];
            const bool sat = (i == a ? ls : ls - psat[i - 1]) != 0;
            if (!sat) {
                if (i < lim) lpos[l + (base_fail + lf - 1)] = i;
            } else {
                if (i >= lim) rpos[l + (tot_sat - (base_sat + ls))] = i;
            }
        }
        group_sync(bar, k, phase);
        // A3: swaps, pairs dealt round-robin to the members
        for (unsigned q = s * BT + tid; q < m; q += k * BT) {
            const unsigned pa = __ldcg(&lpos[l + q]), pc = __ldcg(&rpos[l + q]);
            const float4 va = __ldcg(pp + pa), vc = __ldcg(pp + pc);
            pp[pa] = vc;
            pp[pc] = va;
        }
        group_sync(bar, k, phase);
        if (sweep == 0) {
            lim1 = lim;
            start = lim;
        } else {
            lim2 = lim;
        }
    }
    const unsigned l1 = lim1 - l, l2 = lim2 - l;
    unsigned idx;  // :934-936
    if (l1 > count / 2) idx = l1;
    else if (l2 < count / 2) idx = l2;
    else idx = count / 2;
    // ---- F: divlow / divhigh over the member's slice
    {
        const unsigned len = (count + k - 1) / k;
        const unsigned a = min(l + s * len, r), e = min(a + len, r);
        float dlow = -INFINITY, dhigh = INFINITY;
        for (unsigned i = a + tid; i < e; i += BT) {
            const float v = comp(__ldcg(pp + i), cf);
            if (i - l < idx) dlow = fmaxf(dlow, v);
            else dhigh = fminf(dhigh, v);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            dlow = fmaxf(dlow, __shfl_xor_sync(0xffffffffu, dlow, m));
            dhigh = fminf(dhigh, __shfl_xor_sync(0xffffffffu, dhigh, m));
        }
        __syncthreads();  // s_red still holds the A2 broadcast of the last sweep
        if (lane == 0) {
            s_red[warp] = dlow;
            s_red[NW + warp] = dhigh;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NW; ++w) {
                dlow = fmaxf(dlow, s_red[w]);
                dhigh = fminf(dhigh, s_red[NW + w]);
            }
            if (dlow != -INFINITY) atomicMax(&red[6], f2ord_u(dlow));
            if (dhigh != INFINITY) atomicMax(&red[7], ~f2ord_u(dhigh));
        }
    }
    group_sync(bar, k, phase);
    if (s == 0 && tid == 0) {
        const float dlow = ord2f_u(__ldcg(&red[6])), dhigh = ord2f_u(~__ldcg(&red[7]));
        emit_children_top(t, g, b, l, r, idx, cf, cv, dlow, dhigh, lo, hi, level);
    }
}

// ---- SUBTREES: everything below lives in the shared memory of one CTA ------------------------------------------
struct SubCtx {
    float4* spp;            // [count] the subtree's points, position-ordered
    unsigned short* psat;   // CTA-wide split scratch
    unsigned short* pfail;
    unsigned short* lpos;
    unsigned short* rpos;
    Entry* next;            // next level's lists: [0, LIST_MED) medium, [LIST_MED, ..) small
    unsigned* cnt_next;     // [2] medium, small
    unsigned* nalloc;       // node ids handed out inside the reserved block
    unsigned base_gid;      // global id of the first reserved node
    unsigned nreserved;
    unsigned b, l0;         // item, absolute position of the subtree's first point
};

__device__ __forceinline__ void emit_children_sub(const Tree& t, const SubCtx& sc, const Entry& e, unsigned idx, int cf,
                                                  float cv, float divlow, float divhigh) {
    const unsigned a = atomicAdd(sc.nalloc, 2u);
    if (a + 1 >= sc.nreserved) {
        atomicOr(t.error, 1u);
        return;
    }
    const unsigned ga = sc.base_gid + a;
    store_split(t, e.gid, ga, sc.b, sc.l0 + e.l, sc.l0 + e.r, idx, cf, divlow, divhigh);
    const unsigned cnt[2] = {idx, (unsigned)(e.r - e.l) - idx};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (cnt[k] <= (unsigned)LEAF) continue;
        Entry ce;
        ce.gid = ga + k;
        ce.l = (unsigned short)(k == 0 ? e.l : e.l + idx);
        ce.r = (unsigned short)(k == 0 ? e.l + idx : e.r);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            ce.lo[d] = (k == 1 && d == cf) ? cv : e.lo[d];
            ce.hi[d] = (k == 0 && d == cf) ? cv : e.hi[d];
        }
        if (cnt[k] > (unsigned)SMALL_MAX) {
            const unsigned pos = atomicAdd(&sc.cnt_next[0], 1u);
            if (pos < (unsigned)LIST_MED) sc.next[pos] = ce;
            else atomicOr(t.error, 8u);
        } else {
            const unsigned pos = atomicAdd(&sc.cnt_next[1], 1u);
            if (pos < (unsigned)LIST_SMALL) sc.next[LIST_MED + pos] = ce;
            else atomicOr(t.error, 8u);
        }
    }
}

// one warp splits one node of <= SMALL_MAX points in place
__device__ __forceinline__ void split_small_sm(const Tree& t, const SubCtx& sc, const Entry& e, unsigned short* wscr) {
    const int lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    float4* sp = sc.spp + e.l;
    unsigned short* sL = wscr;
    unsigned short* sR = wscr + SMALL_MAX / 2;
    const unsigned count = (unsigned)(e.r - e.l);
    const int nslot = (int)((count + 31) / 32);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int s = 0; s < nslot; ++s) {
        const unsigned p = (unsigned)s * 32 + lane;
        if (p < count) {
            const float4 v = sp[p];
            mn[0] = fminf(mn[0], v.x);
            mx[0] = fmaxf(mx[0], v.x);
            mn[1] = fminf(mn[1], v.y);
            mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z);
            mx[2] = fmaxf(mx[2], v.z);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
        }
    int cf;
    float cv;
    decide_split(e.lo, e.hi, mn, mx, &cf, &cv);
    const float* sval = reinterpret_cast<const float*>(sp) + cf;  // component cf of point p = sval[4*p]
    unsigned start = 0, lim1 = 0, lim2 = 0;
    for (int sweep = 0; sweep < 2; ++sweep) {
        unsigned tot = 0;
        for (int s = 0; s < nslot; ++s) {
            const unsigned p = (unsigned)s * 32 + lane;
            const bool in = p < count && p >= start;
            const float v = in ? sval[4 * p] : 0.f;
            const bool sat = in && (sweep == 0 ? (v < cv) : (v <= cv));
            tot += __popc(__ballot_sync(0xffffffffu, sat));
        }
        const unsigned lim = start + tot;
        unsigned sb = 0, fb = 0, m = 0;
        for (int s = 0; s < nslot; ++s) {
            const unsigned p = (unsigned)s * 32 + lane;
            const bool in = p < count && p >= start;
            const float v = in ? sval[4 * p] : 0.f;
            const bool sat = in && (sweep == 0 ? (v < cv) : (v <= cv));
            const bool fail = in && !sat;
            const unsigned bs = __ballot_sync(0xffffffffu, sat), bf = __ballot_sync(0xffffffffu, fail);
            const bool left_misplaced = fail && p < lim;
            if (left_misplaced) sL[fb + __popc(bf & ltmask)] = (unsigned short)p;
            if (sat && p >= lim) sR[tot - (sb + __popc(bs & ltmask) + 1)] = (unsigned short)p;
            m += __popc(__ballot_sync(0xffffffffu, left_misplaced));
            sb += __popc(bs);
            fb += __popc(bf);
        }
        __syncwarp();
        for (unsigned k = lane; k < m; k += 32) {
            const unsigned a = sL[k], c = sR[k];
            const float4 va = sp[a], vc = sp[c];
            sp[a] = vc;
            sp[c] = va;
        }
        __syncwarp();
        if (sweep == 0) {
            lim1 = lim;
            start = lim;
        } else {
            lim2 = lim;
        }
    }
    unsigned idx;  // :934-936
    if (lim1 > count / 2) idx = lim1;
    else if (lim2 < count / 2) idx = lim2;
    else idx = count / 2;
    float dlow = -INFINITY, dhigh = INFINITY;
    for (int s = 0; s < nslot; ++s) {
        const unsigned p = (unsigned)s * 32 + lane;
        if (p < count) {
            const float v = sval[4 * p];
            if (p < idx) dlow = fmaxf(dlow, v);
            else dhigh = fminf(dhigh, v);
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        dlow = fmaxf(dlow, __shfl_xor_sync(0xffffffffu, dlow, m));
        dhigh = fminf(dhigh, __shfl_xor_sync(0xffffffffu, dhigh, m));
    }
    if (lane == 0) emit_children_sub(t, sc, e, idx, cf, cv, dlow, dhigh);
    __syncwarp();
}

// the whole CTA splits one node of SMALL_MAX < count <= MED_MAX points in place (count <= BT*IPT: one scan per sweep)
__device__ __forceinline__ void split_med_sm(const Tree& t, const SubCtx& sc, const Entry& e, unsigned long long* s_warp,
                                             float* s_red, float* s_bc) {
    const unsigned tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    float4* sp = sc.spp + e.l;
    const unsigned count = (unsigned)(e.r - e.l);
    __syncthreads();
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = tid; i < count; i += BT) {
        const float4 v = sp[i];
        mn[0] = fminf(mn[0], v.x);
        mx[0] = fmaxf(mx[0], v.x);
        mn[1] = fminf(mn[1], v.y);
        mx[1] = fmaxf(mx[1], v.y);
        mn[2] = fminf(mn[2], v.z);
        mx[2] = fmaxf(mx[2], v.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
        }
        if (lane == 0) {
            s_red[d * NW + warp] = mn[d];
            s_red[(3 + d) * NW + warp] = mx[d];
        }
    }
    __syncthreads();
    if (tid == 0) {
        float amn[3], amx[3];
        for (int d = 0; d < 3; ++d) {
            amn[d] = s_red[d * NW];
            amx[d] = s_red[(3 + d) * NW];
            for (int w = 1; w < NW; ++w) {
                amn[d] = fminf(amn[d], s_red[d * NW + w]);
                amx[d] = fmaxf(amx[d], s_red[(3 + d) * NW + w]);
            }
        }
        int cf;
        float cv;
        decide_split(e.lo, e.hi, amn, amx, &cf, &cv);
        s_bc[0] = __int_as_float(cf);
        s_bc[1] = cv;
    }
    __syncthreads();
    const int cf = __float_as_int(s_bc[0]);
    const float cv = s_bc[1];
    const float* sval = reinterpret_cast<const float*>(sp) + cf;

    unsigned start = 0, lim1 = 0, lim2 = 0;
    for (int sweep = 0; sweep < 2; ++sweep) {
        const unsigned i0 = start + tid * IPT;
        unsigned long long f[IPT];
        unsigned long long local = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const unsigned i = i0 + k;
            unsigned long long fl = 0;
            if (i < count) {
                const float v = sval[4 * i];
                const bool sat = sweep == 0 ? (v < cv) : (v <= cv);
                fl = sat ? 1ull : (1ull << 32);
            }
            local += fl;
            f[k] = local;
        }
        unsigned long long tot;
        const unsigned long long incl = block_scan_incl(local, s_warp, &tot);
        const unsigned long long excl = incl - local;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const unsigned i = i0 + k;
            if (i < count) {
                const unsigned long long v = excl + f[k];
                sc.psat[i] = (unsigned short)(v & 0xFFFFull);
                sc.pfail[i] = (unsigned short)((v >> 32) & 0xFFFFull);
            }
        }
        __syncthreads();
        const unsigned tot_sat = (unsigned)(tot & 0xFFFFFFFFull);
        const unsigned lim = start + tot_sat;
        const unsigned m = lim > start ? sc.pfail[lim - 1] : 0u;
        for (unsigned i = start + tid; i < count; i += BT) {
            const unsigned cs = sc.psat[i], cfl = sc.pfail[i];
            const bool sat = (i == start ? cs : cs - sc.psat[i - 1]) != 0;
            if (!sat) {
                if (i < lim) sc.lpos[cfl - 1] = (unsigned short)i;
            } else {
                if (i >= lim) sc.rpos[tot_sat - cs] = (unsigned short)i;
            }
        }
        __syncthreads();
        for (unsigned k = tid; k < m; k += BT) {
            const unsigned a = sc.lpos[k], c = sc.rpos[k];
            const float4 va = sp[a], vc = sp[c];
            sp[a] = vc;
            sp[c] = va;
        }
        __syncthreads();
        if (sweep == 0) {
            lim1 = lim;
            start = lim;
        } else {
            lim2 = lim;
        }
    }
    unsigned idx;
    if (lim1 > count / 2) idx = lim1;
    else if (lim2 < count / 2) idx = lim2;
    else idx = count / 2;
    float dlow = -INFINITY, dhigh = INFINITY;
    for (unsigned i = tid; i < count; i += BT) {
        const float v = sval[4 * i];
        if (i < idx) dlow = fmaxf(dlow, v);
        else dhigh = fminf(dhigh, v);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        dlow = fmaxf(dlow, __shfl_xor_sync(0xffffffffu, dlow, m));
        dhigh = fminf(dhigh, __shfl_xor_sync(0xffffffffu, dhigh, m));
    }
    if (lane == 0) {
        s_red[warp] = dlow;
        s_red[NW + warp] = dhigh;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NW; ++w) {
            dlow = fmaxf(dlow, s_red[w]);
            dhigh = fminf(dhigh, s_red[NW + w]);
        }
        emit_children_sub(t, sc, e, idx, cf, cv, dlow, dhigh);
    }
    __syncthreads();
}

__device__ __forceinline__ void build_subtree(const Tree& t, unsigned groot, unsigned char* dsm,
                                              unsigned long long* s_warp, float* s_red, float* s_bc, unsigned* s_ctl) {
    // s_ctl: [0..1] counts of list 0 (medium, small), [2..3] counts of list 1, [4] nalloc, [5] base node id
    const unsigned tid = threadIdx.x;
    const int warp = tid >> 5;
    float4* spp = reinterpret_cast<float4*>(dsm + SM_PP);
    Entry* lists = reinterpret_cast<Entry*>(dsm + SM_LISTS);
    __syncthreads();  // the previous subtree of this CTA is completely done with shared memory
    const unsigned b = groot / t.cap;
    const unsigned l0 = __ldcg(&t.nodes[groot].l), r0 = __ldcg(&t.nodes[groot].r);
    const unsigned count = r0 - l0;
    float4* pp = t.pp + (size_t)b * t.N + l0;
    for (unsigned i = tid; i < count; i += BT) spp[i] = __ldcg(pp + i);
    if (tid == 0) {
        const unsigned base = atomicAdd(&t.node_count[b], 2u * count);
        if (base + 2u * count > t.cap) atomicOr(t.error, 1u);
        s_ctl[0] = count > (unsigned)SMALL_MAX ? 1u : 0u;
        s_ctl[1] = count > (unsigned)SMALL_MAX ? 0u : 1u;
        s_ctl[2] = s_ctl[3] = 0u;
        s_ctl[4] = 0u;
        s_ctl[5] = base;
        Entry e;
        e.gid = groot;
        e.l = 0;
        e.r = (unsigned short)count;
        for (int d = 0; d < 3; ++d) {
            e.lo[d] = __ldcg(&t.nlo[(size_t)groot * 3 + d]);
            e.hi[d] = __ldcg(&t.nhi[(size_t)groot * 3 + d]);
        }
        lists[count > (unsigned)SMALL_MAX ? 0 : LIST_MED] = e;
    }
    __syncthreads();
    SubCtx sc;
    sc.spp = spp;
    sc.psat = reinterpret_cast<unsigned short*>(dsm + SM_PSAT);
    sc.pfail = reinterpret_cast<unsigned short*>(dsm + SM_PFAIL);
    sc.lpos = reinterpret_cast<unsigned short*>(dsm + SM_LPOS);
    sc.rpos = reinterpret_cast<unsigned short*>(dsm + SM_RPOS);
    sc.nalloc = &s_ctl[4];
    sc.base_gid = b * t.cap + s_ctl[5];
    sc.nreserved = 2u * count;
    sc.b = b;
    sc.l0 = l0;
    const bool ok = s_ctl[5] + 2u * count <= t.cap;
    for (int lvl = 0; ok; ++lvl) {
        const int cur = lvl & 1;
        Entry* cl = lists + (size_t)cur * (LIST_MED + LIST_SMALL);
        sc.next = lists + (size_t)(cur ^ 1) * (LIST_MED + LIST_SMALL);
        sc.cnt_next = &s_ctl[(cur ^ 1) * 2];
        const unsigned nmed = min(s_ctl[cur * 2], (unsigned)LIST_MED);
        const unsigned nsmall = min(s_ctl[cur * 2 + 1], (unsigned)LIST_SMALL);
        if (nmed == 0 && nsmall == 0) break;
        for (unsigned m = 0; m < nmed; ++m) split_med_sm(t, sc, cl[m], s_warp, s_red, s_bc);
        __syncthreads();
        for (unsigned w = warp; w < nsmall; w += NW)
            split_small_sm(t, sc, cl[LIST_MED + w], reinterpret_cast<unsigned short*>(dsm + SM_WSCR) + (size_t)warp * SMALL_MAX);
        __syncthreads();
        if (tid == 0) s_ctl[cur * 2] = s_ctl[cur * 2 + 1] = 0u;
        __syncthreads();
    }
    for (unsigned i = tid; i < count; i += BT) pp[i] = spp[i];
}

__global__ void __launch_bounds__(BT, 1) build_kernel(const float* __restrict__ pts_all, const Tree t) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ unsigned long long s_warp[32];
    __shared__ float s_red[6 * NW];
    __shared__ float s_bc[4];
    __shared__ unsigned s_ctl[8];
    const unsigned tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned phase = 0;
    mark(nullptr, t.tstamps, 0);

    // the whole tie path is enqueued without knowing whether any row was flagged: nothing to do in that case
    if (t.n_flag && __ldcg(t.n_flag) == 0) return;

    // ---- roots: pp = (point, identity index) (init_vind :1232-1238), data bbox (computeBoundingBox :1241-1263).
    // kroot CTAs share one item; the last of them to finish (ticket) writes the root record and queues it.
    {
        const unsigned kroot = max(1u, gridDim.x / t.B);
        for (unsigned w0 = blockIdx.x; w0 < t.B * kroot; w0 += gridDim.x) {
            const unsigned b = w0 / kroot, sl = w0 % kroot;
            if (t.item_needed && !t.item_needed[b]) continue;
            const float* pts = pts_all + (size_t)b * t.N * 3;
            float4* pp = t.pp + (size_t)b * t.N;
            const unsigned len = (t.N + kroot - 1) / kroot;
            const unsigned a = min(sl * len, t.N), e = min(a + len, t.N);
            float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (unsigned i = a + tid; i < e; i += BT) {
                float4 v;
                v.x = __ldg(pts + 3 * (size_t)i);
                v.y = __ldg(pts + 3 * (size_t)i + 1);
                v.z = __ldg(pts + 3 * (size_t)i + 2);
                v.w = __uint_as_float(i);
                pp[i] = v;
                mn[0] = fminf(mn[0], v.x);
                mx[0] = fmaxf(mx[0], v.x);
                mn[1] = fminf(mn[1], v.y);
                mx[1] = fmaxf(mx[1], v.y);
                mn[2] = fminf(mn[2], v.z);
                mx[2] = fmaxf(mx[2], v.z);
            }
            __syncthreads();
#pragma unroll
            for (int d = 0; d < 3; ++d) {
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) {
                    mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], m));
                    mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], m));
                }
                if (lane == 0) {
                    s_red[d * NW + warp] = mn[d];
                    s_red[(3 + d) * NW + warp] = mx[d];
                }
            }
            __syncthreads();
            unsigned* rr = t.root_red + (size_t)b * 8;  // [0..2] ~ord(min), [3..5] ord(max), [6] ticket
            if (tid < 6) {
                float v = s_red[tid * NW];
                for (int w = 1; w < NW; ++w) v = tid < 3 ? fminf(v, s_red[tid * NW + w]) : fmaxf(v, s_red[tid * NW + w]);
                if (tid < 3 && v != INFINITY) atomicMax(&rr[tid], ~f2ord_u(v));
                if (tid >= 3 && v != -INFINITY) atomicMax(&rr[tid], f2ord_u(v));
            }
            __threadfence();
            __syncthreads();
            if (tid == 0 && atomicAdd(&rr[6], 1u) == kroot - 1) {
                __threadfence();
                const unsigned g = b * t.cap;
                for (int d = 0; d < 3; ++d) {
                    const float lo_ = ord2f_u(~__ldcg(&rr[d])), hi_ = ord2f_u(__ldcg(&rr[3 + d]));
                    t.nlo[(size_t)g * 3 + d] = lo_;
                    t.nhi[(size_t)g * 3 + d] = hi_;
                    t.root_lo[b * 3 + d] = lo_;
                    t.root_hi[b * 3 + d] = hi_;
                }
                NodeRec root;
                root.c1 = root.c2 = -1;
                root.feat = 0;
                root.pad = 0;
                root.divlow = root.divhigh = 0.f;
                root.l = 0;
                root.r = t.N;
                t.nodes[g] = root;
                t.node_count[b] = 1;
                if (t.N > (unsigned)MED_MAX) {
                    const unsigned pos = atomicAdd(&t.list_cnt[0], 1u);
                    t.list[pos] = g;
                } else if (t.N > (unsigned)LEAF) {
                    const unsigned pos = atomicAdd(&t.list_cnt[MAX_LEVELS + 1], 1u);
                    t.sublist[pos] = g;
                }
            }
            __syncthreads();
        }
    }
    grid_sync(t.barrier, phase);
    mark(nullptr, t.tstamps, 1);

    // ---- TOP levels
    for (int level = 0; level < MAX_LEVELS; ++level) {
        const unsigned nbig = __ldcg(&t.list_cnt[level]);
        if (nbig == 0) break;
        const unsigned* cur = t.list + (size_t)(level & 1) * t.lcap;
        const unsigned kgrp = gridDim.x / nbig;  // CTAs per node when the level has fewer nodes than CTAs
        if (kgrp >= 2 && level < MAX_GROUP_LEVELS) {
            const unsigned slot = blockIdx.x / kgrp;
            if (slot < nbig) split_big_group(t, __ldcg(cur + slot), level, slot, kgrp, blockIdx.x % kgrp, s_warp, s_red);
        } else {
            for (unsigned i = blockIdx.x; i < nbig; i += gridDim.x)
                split_big(t, __ldcg(cur + i), level, s_warp, s_red, s_bc);
        }
        grid_sync(t.barrier, phase);
        if (level < 8) mark(nullptr, t.tstamps, 2 + level);
    }
    mark(nullptr, t.tstamps, 10);
    // ---- SUBTREES (independent; no further grid barrier)
    const unsigned nsub = min(__ldcg(&t.list_cnt[MAX_LEVELS + 1]), t.lcap);
    for (unsigned i = blockIdx.x; i < nsub; i += gridDim.x)
        build_subtree(t, __ldcg(t.sublist + i), dyn_smem, s_warp, s_red, s_bc, s_ctl);
    mark(nullptr, t.tstamps, 11);  // CTA 0's own end
    if (t.tstamps && threadIdx.x == 0) atomicMax(&t.tstamps[12], gtimer());  // last CTA's end
}

__global__ void mark_items_kernel(const unsigned* __restrict__ flag_list, const unsigned* __restrict__ n_flag_ptr,
                                  unsigned Q, unsigned char* __restrict__ item_needed) {
    const unsigned n_flag = *n_flag_ptr;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_flag; i += gridDim.x * blockDim.x)
        item_needed[flag_list[i] / Q] = 1;
}

__device__ __forceinline__ NodeRec load_node(const NodeRec* p) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    NodeRec n;
    n.c1 = (int)a.x;
    n.c2 = (int)a.y;
    n.feat = (int)a.z;
    n.pad = a.w;
    n.divlow = __uint_as_float(c.x);
    n.divhigh = __uint_as_float(c.y);
    n.l = c.z;
    n.r = c.w;
    return n;
}

// ---- exact replay of nanoflann's search for the flagged rows -------------------------------------------------
// per_warp != 0: one query per 32-thread CTA, walked by lane 0 alone -- a pointer-chasing DFS gains nothing from SIMT
// and loses to divergence, so with few flagged rows every query gets its own warp scheduler slot; per_warp == 0
// (duplicate-heavy clouds, most rows flagged): one query per thread.
template <typename OutT>
__global__ void __launch_bounds__(32) exact_query_kernel(const float* __restrict__ q_all, const Tree t, unsigned Q,
                                                         int K, const unsigned* __restrict__ flag_list,
                                                         const unsigned* __restrict__ n_flag_ptr,
                                                         OutT* __restrict__ out) {
    // persistent grid over the flagged rows: with few rows each query gets a warp to itself (walked by lane 0),
    // with many rows (duplicate-heavy clouds) every thread takes rows
    const unsigned n_flag = *n_flag_ptr;
    const bool per_warp = n_flag <= gridDim.x;
    if (per_warp && threadIdx.x != 0) return;
    const unsigned f0 = per_warp ? blockIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned fstride = per_warp ? gridDim.x : gridDim.x * blockDim.x;
  for (unsigned f = f0; f < n_flag; f += fstride) {
    const unsigned row = flag_list[f];
    const unsigned b = row / Q;
    const float4* pp = t.pp + (size_t)b * t.N;
    const NodeRec* nodes = t.nodes + (size_t)b * t.cap;
    const float q[3] = {q_all[3 * (size_t)row], q_all[3 * (size_t)row + 1], q_all[3 * (size_t)row + 2]};

    float rd[MAX_K];
    unsigned ri[MAX_K];
    int count = 0;
    rd[K - 1] = 3.402823466e+38f;  // KNNResultSet::init (:47-53)

    // computeInitialDistances (:977-995) against the root bbox
    float d0[3] = {0.f, 0.f, 0.f};
    float distsq = 0.f;
    for (int d = 0; d < 3; ++d) {
        const float blo = t.root_lo[b * 3 + d], bhi = t.root_hi[b * 3 + d];
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
    st_node[0] = 0;
    st_min[0] = distsq;
    st_d[0][0] = d0[0];
    st_d[0][1] = d0[1];
    st_d[0][2] = d0[2];
    int sp = 1;
    bool first = true;
    while (sp > 0) {
        --sp;
        const int node = st_node[sp];
        const float mind = st_min[sp];
        float dd[3] = {st_d[sp][0], st_d[sp][1], st_d[sp][2]};
        if (!first && !(mind <= rd[K - 1])) continue;  // mindistsq*epsError <= worstDist()  (:1319)
        first = false;
        NodeRec nd = load_node(nodes + node);
        while (nd.c1 >= 0) {
            const int idx = nd.feat;
            const float val = idx == 0 ? q[0] : (idx == 1 ? q[1] : q[2]);
            const float ddi = idx == 0 ? dd[0] : (idx == 1 ? dd[1] : dd[2]);
            const float diff1 = __fsub_rn(val, nd.divlow), diff2 = __fsub_rn(val, nd.divhigh);
            int best, other;
            float cut;
            if (__fadd_rn(diff1, diff2) < 0.f) {
                best = nd.c1;
                other = nd.c2;
                cut = __fmul_rn(diff2, diff2);
            } else {
                best = nd.c2;
                other = nd.c1;
                cut = __fmul_rn(diff1, diff1);
            }
            if (sp < MAX_DEPTH) {
                st_node[sp] = other;
                st_min[sp] = __fsub_rn(__fadd_rn(mind, cut), ddi);
                st_d[sp][0] = idx == 0 ? cut : dd[0];
                st_d[sp][1] = idx == 1 ? cut : dd[1];
                st_d[sp][2] = idx == 2 ? cut : dd[2];
                ++sp;
            } else {
                atomicOr(t.error, 4u);  // deeper than MAX_DEPTH: reported to the host, never silently wrong
            }
            nd = load_node(nodes + best);
        }
        // leaf (:1277-1287): fetch its <= LEAF points with independent loads, then insert in vind order
        const int n = (int)(nd.r - nd.l);
        float4 buf[LEAF];
#pragma unroll
        for (int k = 0; k < LEAF; ++k)
            if (k < n) buf[k] = __ldg(pp + nd.l + k);
        const float worst = rd[K - 1];  // snapshot once per leaf (:1277)
#pragma unroll
        for (int k = 0; k < LEAF; ++k) {
            if (k < n) {
                const float dx = __fsub_rn(q[0], buf[k].x), dy = __fsub_rn(q[1], buf[k].y), dz = __fsub_rn(q[2], buf[k].z);
                const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (dist < worst) {  // KNNResultSet::addPoint (:72-96), strict '>' shifting
                    const unsigned index = __float_as_uint(buf[k].w);
                    int jj;
                    for (jj = count; jj > 0; --jj) {
                        if (rd[jj - 1] > dist) {
                            if (jj < K) {
                                rd[jj] = rd[jj - 1];
                                ri[jj] = ri[jj - 1];
                            }
                        } else
                            break;
                    }
                    if (jj < K) {
                        rd[jj] = dist;
                        ri[jj] = index;
                    }
                    if (count < K) ++count;
                }
            }
        }
    }
    OutT* o = out + (size_t)row * K;
    for (int jj = 0; jj < count; ++jj) o[jj] = (OutT)ri[jj];
  }
}

enum { TW_BASE = 16 };  // workspace slots TW_BASE.. are owned by this header

// Carve the tree arrays out of workspace slabs and zero the small control block.
static int alloc_tree(Ctx* c, cudaStream_t s, size_t B, size_t N, Tree* out) {
    const size_t cap = 3 * N + 64;
    const size_t lcap = B * (N / (LEAF + 1) + 2) + 16;
    SSDR_TRY(c->ws[TW_BASE + 0].reserve(B * N * (sizeof(float4) + 4 * sizeof(unsigned))));
    SSDR_TRY(c->ws[TW_BASE + 1].reserve(B * cap * (sizeof(NodeRec) + 6 * sizeof(float))));
    SSDR_TRY(c->ws[TW_BASE + 2].reserve(3 * lcap * sizeof(unsigned)));
    const size_t ctl_words = 6 * B + B + (size_t)(MAX_LEVELS + 2) + 8 + (B + 3) / 4 + 4 + 8 * B;
    SSDR_TRY(c->ws[TW_BASE + 3].reserve(ctl_words * sizeof(unsigned)));
    const size_t G = (size_t)c->sm_count;
    const size_t grp_words = (size_t)MAX_GROUP_LEVELS * G * 9 + G * 2 * G * 2;
    SSDR_TRY(c->ws[TW_BASE + 4].reserve(grp_words * sizeof(unsigned)));
    Tree t;
    t.N = (unsigned)N;
    t.cap = (unsigned)cap;
    t.B = (unsigned)B;
    t.lcap = (unsigned)lcap;
    t.pp = c->ws[TW_BASE + 0].as<float4>();
    unsigned* pb = reinterpret_cast<unsigned*>(t.pp + B * N);
    t.lpos = pb;
    t.rpos = pb + B * N;
    t.psat = pb + 2 * B * N;
    t.pfail = pb + 3 * B * N;
    t.nodes = c->ws[TW_BASE + 1].as<NodeRec>();
    t.nlo = reinterpret_cast<float*>(t.nodes + B * cap);
    t.nhi = t.nlo + B * cap * 3;
    t.list = c->ws[TW_BASE + 2].as<unsigned>();
    t.sublist = t.list + 2 * lcap;
    unsigned* ctl = c->ws[TW_BASE + 3].as<unsigned>();
    SSDR_CHECK_CUDA(cudaMemsetAsync(ctl, 0, ctl_words * sizeof(unsigned), s));
    t.root_lo = reinterpret_cast<float*>(ctl);
    t.root_hi = reinterpret_cast<float*>(ctl + 3 * B);
    t.node_count = ctl + 6 * B;
    t.list_cnt = ctl + 7 * B;
    t.barrier = t.list_cnt + (size_t)(MAX_LEVELS + 2);
    t.error = t.barrier + 4;
    t.gbar = c->ws[TW_BASE + 4].as<unsigned>();
    t.gred = t.gbar + (size_t)MAX_GROUP_LEVELS * G;
    t.gpart = t.gred + (size_t)MAX_GROUP_LEVELS * G * 8;
    SSDR_CHECK_CUDA(cudaMemsetAsync(t.gbar, 0, (size_t)MAX_GROUP_LEVELS * G * 9 * sizeof(unsigned), s));
    t.root_red = t.error + 4 + (B + 3) / 4 + 1;
    t.n_flag = nullptr;
    t.item_needed = nullptr;
    t.tstamps = nullptr;
    *out = t;
    return SSDR_OK;
}
static unsigned char* tree_needed_flags(const Tree& t) { return reinterpret_cast<unsigned char*>(t.error + 4); }

static int launch_build(Ctx* c, cudaStream_t s, const float* d_pts, const Tree& t) {
    const size_t smem = SM_TOTAL;
    SSDR_CHECK_CUDA(cudaFuncSetAttribute(build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    SSDR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, build_kernel, BT, smem));
    SSDR_REQUIRE(nb >= 1, SSDR_ERR_CUDA, "tree build kernel does not fit on an SM");
    const float* pts = d_pts;
    Tree tt = t;
    void* args[] = {(void*)&pts, (void*)&tt};
    SSDR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)build_kernel, dim3(c->sm_count), dim3(BT), args, smem, s));
    return SSDR_OK;
}

static int check_tree_error(Ctx* c, cudaStream_t s, const Tree& t) {
    unsigned h_err = 0;
    SSDR_TRY(d2h_sync(c, &h_err, t.error, sizeof(unsigned), s));
    SSDR_REQUIRE(h_err == 0, SSDR_ERR_UNSUPPORTED,
                 "exact tie path gave up (flags %u: 1 node capacity, 2 more than %d tree levels, 4 search stack deeper "
                 "than %d, 8 work list capacity)", h_err, MAX_LEVELS, MAX_DEPTH);
    return SSDR_OK;
}

// Enqueue the whole tie path behind the main kernel WITHOUT a host round trip: mark the items that own flagged rows,
// build their trees, overwrite the flagged rows with nanoflann's exact answer.  Every kernel reads the flagged-row
// count on the device and returns at once when it is zero.  The caller checks t_out->error after its final sync.
template <typename OutT>
static int enqueue_tie_path(Ctx* c, cudaStream_t s, const float* d_pts, size_t B, size_t N, const float* d_q, size_t Q,
                            size_t K, OutT* d_out, const unsigned* flag_list, const unsigned* d_flag_count,
                            Tree* t_out, cudaEvent_t ev_mid = nullptr) {
    SSDR_REQUIRE(K <= (size_t)MAX_K, SSDR_ERR_UNSUPPORTED, "K=%zu > %d in the exact tie path", K, MAX_K);
    Tree t;
    SSDR_TRY(alloc_tree(c, s, B, N, &t));
    unsigned char* needed = tree_needed_flags(t);
    t.item_needed = needed;
    t.n_flag = d_flag_count;
    mark_items_kernel<<<64, 256, 0, s>>>(flag_list, d_flag_count, (unsigned)Q, needed);
    SSDR_TRY(launch_build(c, s, d_pts, t));
    if (ev_mid) SSDR_CHECK_CUDA(cudaEventRecord(ev_mid, s));
    exact_query_kernel<OutT><<<(unsigned)c->sm_count * 16, 32, 0, s>>>(d_q, t, (unsigned)Q, (int)K, flag_list,
                                                                       d_flag_count, d_out);
    SSDR_CHECK_CUDA(cudaGetLastError());
    *t_out = t;
    return SSDR_OK;
}

static int tree_error_to_status(unsigned h_err) {
    SSDR_REQUIRE(h_err == 0, SSDR_ERR_UNSUPPORTED,
                 "exact tie path gave up (flags %u: 1 node capacity, 2 more than %d tree levels, 4 search stack deeper "
                 "than %d, 8 work list capacity)", h_err, MAX_LEVELS, MAX_DEPTH);
    return SSDR_OK;
}

}  // namespace kdtree
}  // namespace ssdr
