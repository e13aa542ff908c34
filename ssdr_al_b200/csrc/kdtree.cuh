// kdtree.cuh -- nanoflann-identical KD-tree on the device, used to resolve KNN rows whose neighbour order depends on
// nanoflann's tree-visit order (equal / almost equal fp32 distances).
//
// build_kernel reproduces KDTreeSingleIndexAdaptor::buildIndex (utils/nearest_neighbors/nanoflann.hpp:1136-1147,
// divideTree :848-896, middleSplit_ :898-937, planeSplit :948-975) bit for bit -- same vind permutation, same
// divfeat/divlow/divhigh -- as ONE persistent cooperative launch for all batch items.  The working array is
// position-ordered float4 (x, y, z, point index), i.e. nanoflann's vind with the coordinates carried along, so
// every pass is a coalesced stream and the later search reads a leaf with one load per point:
//   * TOP: nodes with more than MED_MAX points are split level by level from global memory -- one CTA per node, or a
//     group of CTAs per node (private group barrier, per-member partial counts) while a level has fewer nodes than
//     SMs -- with a grid barrier between levels; a node's tight bbox (computeMinMax) comes from its parent's last pass;
//   * SUBTREES: every node of <= MED_MAX points is handed to one CTA that loads it into shared memory ONCE and grows
//     its whole subtree there -- groups of 2..32 warps (named barriers) for nodes > PER_WARP points, and 1..8 nodes
//     per warp below that (aligned lane slices, masked ballot/popc prefix counts) -- then writes the permuted
//     points back;
//   * each of planeSplit's two Hoare sweeps is a prefix count: the k-th misplaced element from the left swaps with
//     the k-th misplaced element from the right (checked against the sequential code in
//     tools/proto_kdtree.py and tests/test_knn_gpu.py::test_device_tree_equals_sequential_tree); the second sweep
//     only moves points equal to the cut value and is skipped when the node has none;
//   * the last trees stay valid in their workspace: a later call with the same support cloud (checked on the device)
//     reuses them instead of building again.
// exact_query_kernel replays findNeighbors/searchLevel (:1163-1178, :1270-1328) and KNNResultSet::addPoint
// (:72-96) with one thread per flagged query, an explicit stack and one 32-byte node record per visit.
#pragma once
#include "common.cuh"

namespace ssdr {
namespace kdtree {

constexpr int BT = 1024;
constexpr int NW = BT / 32;
constexpr int IPT_BIG = 4;      // top-level scan chunk per thread (8 and loop unrolling measured slower: spills)
constexpr int LEAF = 10;
constexpr int MED_MAX = 4096;     // whole subtree grown in the shared memory of one CTA (<= 8192: 32 warps x PER_WARP)
constexpr int PER_WARP = 256;     // points per warp when a group of warps splits a node (IPT_SUB rows of 32)
constexpr int IPT_SUB = PER_WARP / 32;
constexpr int LIST_BIG = MED_MAX / (PER_WARP + 1) + 1;   // nodes of > PER_WARP points inside one subtree level
constexpr int LIST_SMALL = MED_MAX / (LEAF + 1) + 1;     // split-able nodes inside one subtree level
constexpr int MAX_ROUNDS = 8;     // rounds of warp groups scheduled at once
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
constexpr size_t SM_LPOS = SM_PP + (size_t)MED_MAX * 16;   // u16 [MED_MAX]: k-th misplaced position from the left
constexpr size_t SM_RPOS = SM_LPOS + (size_t)MED_MAX * 2;  // u16 [MED_MAX]: ... from the right
constexpr size_t SM_LISTS = SM_RPOS + (size_t)MED_MAX * 2;
constexpr size_t SM_CLS = SM_LISTS + 2 * (size_t)(LIST_BIG + LIST_SMALL) * sizeof(Entry);  // u16 [2][4][LIST_SMALL]
constexpr size_t SM_TOTAL = SM_CLS + 2 * 4 * (size_t)LIST_SMALL * 2;

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
    float* tmn;            // [B*cap*3] tight min / max of the node's points (computeMinMax), handed down by the parent's
    float* tmx;            //           last pass so that a split starts without a pass of its own
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
    unsigned char* item_flags;         // [B] storage of the item_needed flags inside the control block
    unsigned* built;                   // control word: 1 once the trees of this workspace have been built
    int skip_if_built;                 // asynchronous reuse: a later call on the SAME cloud builds only if nobody did
};

__device__ __forceinline__ unsigned f2ord_u(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f_u(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// both children's tight bboxes, accumulated point by point (identity = +-inf)
struct ChildBox {
    float mn[2][3], mx[2][3];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                mn[c][d] = INFINITY;
                mx[c][d] = -INFINITY;
            }
    }
    __device__ __forceinline__ void add(const float4& v, bool right) {
        mn[0][0] = fminf(mn[0][0], right ? INFINITY : v.x);
        mn[0][1] = fminf(mn[0][1], right ? INFINITY : v.y);
        mn[0][2] = fminf(mn[0][2], right ? INFINITY : v.z);
        mx[0][0] = fmaxf(mx[0][0], right ? -INFINITY : v.x);
        mx[0][1] = fmaxf(mx[0][1], right ? -INFINITY : v.y);
        mx[0][2] = fmaxf(mx[0][2], right ? -INFINITY : v.z);
        mn[1][0] = fminf(mn[1][0], right ? v.x : INFINITY);
        mn[1][1] = fminf(mn[1][1], right ? v.y : INFINITY);
        mn[1][2] = fminf(mn[1][2], right ? v.z : INFINITY);
        mx[1][0] = fmaxf(mx[1][0], right ? v.x : -INFINITY);
        mx[1][1] = fmaxf(mx[1][1], right ? v.y : -INFINITY);
        mx[1][2] = fmaxf(mx[1][2], right ? v.z : -INFINITY);
    }
    // reduce over the lanes named in `mask` (redux.sync on order-preserving encodings); every lane of the mask gets it
    __device__ __forceinline__ void reduce(unsigned mask) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                mn[c][d] = ord2f_u(__reduce_min_sync(mask, f2ord_u(mn[c][d])));
                mx[c][d] = ord2f_u(__reduce_max_sync(mask, f2ord_u(mx[c][d])));
            }
    }
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
                                                  unsigned idx, int cf, float cv, const float cmn[2][3],
                                                  const float cmx[2][3], const float lo[3], const float hi[3],
                                                  int level) {
    const float divlow = cf == 0 ? cmx[0][0] : (cf == 1 ? cmx[0][1] : cmx[0][2]);   // :886-887
    const float divhigh = cf == 0 ? cmn[1][0] : (cf == 1 ? cmn[1][1] : cmn[1][2]);
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
        t.tmn[(size_t)ga * 3 + d] = cmn[0][d];
        t.tmx[(size_t)ga * 3 + d] = cmx[0][d];
        t.tmn[(size_t)gb * 3 + d] = cmn[1][d];
        t.tmx[(size_t)gb * 3 + d] = cmx[1][d];
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
                                          float* s_red) {
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
    // computeMinMax (:827-836) was done by the parent's last pass (roots: the data bbox)
    float amn[3], amx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        amn[d] = __ldcg(&t.tmn[(size_t)g * 3 + d]);
        amx[d] = __ldcg(&t.tmx[(size_t)g * 3 + d]);
    }
    int cf;
    float cv;
    decide_split(lo, hi, amn, amx, &cf, &cv);

    // planeSplit (:948-975).  The second sweep ("<= cutval" over what the first left on the right) only moves points
    // that EQUAL cutval: when the node has none (the usual case for an unclamped mid-plane) it is a no-op and skipped.
    unsigned start = l, lim1 = l, lim2 = l;
    bool has_eq = false;
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
                    has_eq |= v == cv;
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
        const bool any_eq = __syncthreads_or(has_eq) != 0;
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
            if (!any_eq) {
                lim2 = lim;
                break;
            }
        } else {
            lim2 = lim;
        }
    }
    const unsigned l1 = lim1 - l, l2 = lim2 - l;
    unsigned idx;  // :934-936
    if (l1 > count / 2) idx = l1;
    else if (l2 < count / 2) idx = l2;
    else idx = count / 2;
    // tight bboxes of both children in one pass (their computeMinMax; divlow/divhigh are two of the twelve values)
    ChildBox cb;
    cb.init();
    for (unsigned i = l + tid; i < r; i += BT) cb.add(__ldcg(pp + i), i - l >= idx);
    cb.reduce(0xffffffffu);
    unsigned* sr = reinterpret_cast<unsigned*>(s_red);
    __syncthreads();  // s_red may still be read by the sweeps' stragglers
    if (tid < 12) sr[tid] = (tid % 6 < 3) ? 0xFFFFFFFFu : 0u;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                atomicMin(&sr[c2 * 6 + d], f2ord_u(cb.mn[c2][d]));
                atomicMax(&sr[c2 * 6 + 3 + d], f2ord_u(cb.mx[c2][d]));
            }
    }
    __syncthreads();
    if (tid == 0) {
        for (int c2 = 0; c2 < 2; ++c2)
            for (int d = 0; d < 3; ++d) {
                cb.mn[c2][d] = ord2f_u(sr[c2 * 6 + d]);
                cb.mx[c2][d] = ord2f_u(sr[c2 * 6 + 3 + d]);
            }
        emit_children_top(t, g, b, l, r, idx, cf, cv, cb.mn, cb.mx, lo, hi, level);
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
    unsigned* red = t.gred + ((size_t)level * G + slot) * 12;
    unsigned phase = 0;
    const unsigned l = __ldcg(&t.nodes[g].l), r = __ldcg(&t.nodes[g].r);
    const unsigned count = r - l;
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = __ldcg(&t.nlo[(size_t)g * 3 + d]);
        hi[d] = __ldcg(&t.nhi[(size_t)g * 3 + d]);
    }
    // computeMinMax was done by the parent's last pass (roots: the data bbox)
    float tmn[3], tmx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        tmn[d] = __ldcg(&t.tmn[(size_t)g * 3 + d]);
        tmx[d] = __ldcg(&t.tmx[(size_t)g * 3 + d]);
    }
    int cf;
    float cv;
    decide_split(lo, hi, tmn, tmx, &cf, &cv);

    // ---- planeSplit (:948-975), two sweeps; the second one is skipped when no point of the node equals cutval
    // (every member learns that from bit 31 of the members' false counts)
    unsigned start = l, lim1 = l, lim2 = l;
    bool has_eq = false;
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
                    has_eq |= v == cv;
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
        const bool cta_eq = __syncthreads_or(has_eq) != 0;
        if (tid == 0) {
            part[2 * s] = (unsigned)(carry & 0xFFFFFFFFull);
            part[2 * s + 1] = (unsigned)(carry >> 32) | (cta_eq ? 0x80000000u : 0u);
        }
        group_sync(bar, k, phase);
        // A2: global ranks = member base + local counts (warp 0 sums the member totals, smem broadcasts them)
        if (warp == 0) {
            unsigned bs = 0, bf = 0, ts = 0, mm = 0, eqf = 0;
            const unsigned lim_guess_len = len;
            for (unsigned q = lane; q < k; q += 32) {
                const unsigned ps = __ldcg(&part[2 * q]), pfw = __ldcg(&part[2 * q + 1]);
                const unsigned pf = pfw & 0x7FFFFFFFu;
                eqf |= pfw >> 31;
                if (q < s) {
                    bs += ps;
                    bf += pf;
                }
                ts += ps;
            }
            eqf = __any_sync(0xffffffffu, eqf != 0) ? 1u : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                bs += __shfl_xor_sync(0xffffffffu, bs, o);
                bf += __shfl_xor_sync(0xffffffffu, bf, o);
                ts += __shfl_xor_sync(0xffffffffu, ts, o);
            }
            const unsigned lim_w = start + ts;
            if (lim_w > start) {  // misplaced pairs = predicate-false positions in [start, lim)
                const unsigned sq = (lim_w - 1 - start) / lim_guess_len;  // member that owns position lim-1
                for (unsigned q = lane; q < sq; q += 32) mm += __ldcg(&part[2 * q + 1]) & 0x7FFFFFFFu;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mm += __shfl_xor_sync(0xffffffffu, mm, o);
                mm += __ldcg(&pfail[lim_w - 1]);
            }
            if (lane == 0) {
                s_red[0] = __uint_as_float(bs);
                s_red[1] = __uint_as_float(bf);
                s_red[2] = __uint_as_float(ts);
                s_red[3] = __uint_as_float(mm);
                s_red[4] = __uint_as_float(eqf);
            }
        }
        __syncthreads();
        const unsigned base_sat = __float_as_uint(s_red[0]), base_fail = __float_as_uint(s_red[1]);
        const unsigned tot_sat = __float_as_uint(s_red[2]), m = __float_as_uint(s_red[3]);
        const bool any_eq = __float_as_uint(s_red[4]) != 0u;
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
            if (!any_eq) {  // same decision in every member: the barrier phases stay aligned
                lim2 = lim;
                break;
            }
        } else {
            lim2 = lim;
        }
    }
    const unsigned l1 = lim1 - l, l2 = lim2 - l;
    unsigned idx;  // :934-936
    if (l1 > count / 2) idx = l1;
    else if (l2 < count / 2) idx = l2;
    else idx = count / 2;
    // ---- F: both children's tight bboxes over the member's slice (identity 0: [c*6+d] = ~ord(min), [c*6+3+d] = ord(max))
    {
        const unsigned len = (count + k - 1) / k;
        const unsigned a = min(l + s * len, r), e = min(a + len, r);
        ChildBox cb;
        cb.init();
        for (unsigned i = a + tid; i < e; i += BT) cb.add(__ldcg(pp + i), i - l >= idx);
        cb.reduce(0xffffffffu);
        unsigned* sr = reinterpret_cast<unsigned*>(s_red);
        __syncthreads();  // s_red still holds the A2 broadcast of the last sweep
        if (tid < 12) sr[tid] = 0u;
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (cb.mn[c2][d] != INFINITY) atomicMax(&sr[c2 * 6 + d], ~f2ord_u(cb.mn[c2][d]));
                    if (cb.mx[c2][d] != -INFINITY) atomicMax(&sr[c2 * 6 + 3 + d], f2ord_u(cb.mx[c2][d]));
                }
        }
        __syncthreads();
        if (tid < 12 && sr[tid] != 0u) atomicMax(&red[tid], sr[tid]);
    }
    group_sync(bar, k, phase);
    if (s == 0 && tid == 0) {
        float cmn[2][3], cmx[2][3];
        for (int c2 = 0; c2 < 2; ++c2)
            for (int d = 0; d < 3; ++d) {
                cmn[c2][d] = ord2f_u(~__ldcg(&red[c2 * 6 + d]));
                cmx[c2][d] = ord2f_u(__ldcg(&red[c2 * 6 + 3 + d]));
            }
        emit_children_top(t, g, b, l, r, idx, cf, cv, cmn, cmx, lo, hi, level);
    }
}

// ---- SUBTREES: everything below lives in the shared memory of one CTA ------------------------------------------
struct SubCtx {
    float4* spp;            // [count] the subtree's points, position-ordered
    unsigned short* lpos;   // [count] misplaced positions of the running sweep, at [node.l + k]
    unsigned short* rpos;
    Entry* next;            // next level's lists: [0, LIST_BIG) big, [LIST_BIG, ..) small
    unsigned short* cls_next;  // [4][LIST_SMALL] small entries by size class (slots into next + LIST_BIG)
    unsigned* cnt_next;     // [6] big, small, then the four size classes (<= 256, 128, 64, 32 points)
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
        if (cnt[k] > (unsigned)PER_WARP) {
            const unsigned pos = atomicAdd(&sc.cnt_next[0], 1u);
            if (pos < (unsigned)LIST_BIG) sc.next[pos] = ce;
            else atomicOr(t.error, 8u);
        } else {
            const unsigned pos = atomicAdd(&sc.cnt_next[1], 1u);
            if (pos < (unsigned)LIST_SMALL) {
                sc.next[LIST_BIG + pos] = ce;
                const int c = cnt[k] > 128u ? 0 : (cnt[k] > 64u ? 1 : (cnt[k] > 32u ? 2 : 3));
                sc.cls_next[c * LIST_SMALL + atomicAdd(&sc.cnt_next[2 + c], 1u)] = (unsigned short)pos;
            } else {
                atomicOr(t.error, 8u);
            }
        }
    }
}

// named barrier over the g warps of a group (g == 1: the warp itself); ids 1..15, 0 stays __syncthreads'
__device__ __forceinline__ void group_bar(int g, int id) {
    if (g == 1) __syncwarp();
    else if (g == NW) __syncthreads();  // the whole CTA works on one node: nobody else can be at barrier 0
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(g * 32) : "memory");
}

// warps that split a node of `count` points together: a power of two, PER_WARP points per warp
__device__ __forceinline__ int group_warps(unsigned count) {
    const unsigned need = (count + PER_WARP - 1) / PER_WARP;
    int g = 1;
    while ((unsigned)g < need) g <<= 1;
    return g;
}

// One group of g warps (warps [off, off+g) of the CTA) splits one node in place: computeMinMax + middleSplit_
// (:898-929), planeSplit's two Hoare sweeps (:942-975) as prefix counts, the children's tight cut bounds (:886-887).
// Every warp owns a contiguous chunk of 32*ipt positions, lanes interleaved inside it, so a position's rank among
// the predicate-true / -false positions is (warps before) + (rows before in the warp) + (lanes before in the row):
// two ballots per row and one packed count per warp through shared memory.
struct GroupScratch {
    float* red;        // [8 * NW]  per warp: bbox partials (6), cut-bound partials (2)
    unsigned* tot;     // [NW]      per warp: packed (true count | false count << 16) of the sweep
    unsigned* m;       // [NW]      per group (at its first warp): misplaced pairs of the sweep
    unsigned* eq;      // [NW]      per warp: some point of its chunk equals the cut value
};

__device__ __forceinline__ void split_node_sm(const Tree& t, const SubCtx& sc, const Entry& e, int g, int off,
                                              const GroupScratch& gs) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gw = warp - off;
    const unsigned G = (unsigned)g * 32u, gt = (unsigned)gw * 32u + (unsigned)lane;
    const int bid = 1 + (off >> 1);
    const unsigned ltmask = (1u << lane) - 1u;
    float4* sp = sc.spp + e.l;
    const unsigned count = (unsigned)(e.r - e.l);

    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = gt; i < count; i += G) {
        const float4 v = sp[i];
        mn[0] = fminf(mn[0], v.x);
        mx[0] = fmaxf(mx[0], v.x);
        mn[1] = fminf(mn[1], v.y);
        mx[1] = fmaxf(mx[1], v.y);
        mn[2] = fminf(mn[2], v.z);
        mx[2] = fmaxf(mx[2], v.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(FULL, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL, mx[d], m));
        }
    if (g > 1) {
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                gs.red[d * NW + warp] = mn[d];
                gs.red[(3 + d) * NW + warp] = mx[d];
            }
        }
        group_bar(g, bid);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = lane < g ? gs.red[d * NW + off + lane] : INFINITY;
            mx[d] = lane < g ? gs.red[(3 + d) * NW + off + lane] : -INFINITY;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                mn[d] = fminf(mn[d], __shfl_xor_sync(FULL, mn[d], m));
                mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL, mx[d], m));
            }
        }
    }
    int cf;
    float cv;
    decide_split(e.lo, e.hi, mn, mx, &cf, &cv);  // every thread, same inputs
    const float* sval = reinterpret_cast<const float*>(sp) + cf;  // component cf of point p = sval[4*p]

    const unsigned ipt = (count + G - 1) / G;  // rows per warp, <= IPT_SUB
    const unsigned W = ipt * 32u;
    unsigned start = 0, lim1 = 0, lim2 = 0;
    for (int sweep = 0; sweep < 2; ++sweep) {
        const unsigned base = start + (unsigned)gw * W;
        unsigned bs[IPT_SUB], bf[IPT_SUB];
        unsigned wsat = 0, wfail = 0;
        bool eq = false;
#pragma unroll
        for (int k = 0; k < IPT_SUB; ++k) {
            bs[k] = bf[k] = 0u;
            if ((unsigned)k < ipt) {
                const unsigned p = base + (unsigned)k * 32u + (unsigned)lane;
                const bool in = p < count;
                const float v = in ? sval[4 * p] : 0.f;
                const bool sat = in && (sweep == 0 ? (v < cv) : (v <= cv));
                eq |= in && v == cv;
                bs[k] = __ballot_sync(FULL, sat);
                bf[k] = __ballot_sync(FULL, in && !sat);
                wsat += __popc(bs[k]);
                wfail += __popc(bf[k]);
            }
        }
        bool any_eq = __any_sync(FULL, eq);
        unsigned sat_before = 0, fail_before = 0, tot_sat = wsat;
        if (g > 1) {
            if (lane == 0) {
                gs.tot[warp] = wsat | (wfail << 16);
                gs.eq[warp] = any_eq ? 1u : 0u;
            }
            group_bar(g, bid);
            any_eq = __any_sync(FULL, lane < g && gs.eq[off + lane] != 0u);
            const unsigned x = lane < g ? gs.tot[off + lane] : 0u;
            unsigned incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned n = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += n;
            }
            const unsigned mine = __shfl_sync(FULL, incl - x, gw);
            sat_before = mine & 0xFFFFu;
            fail_before = mine >> 16;
            tot_sat = __shfl_sync(FULL, incl, 31) & 0xFFFFu;
        }
        const unsigned lim = start + tot_sat;
        unsigned sb = sat_before, fb = fail_before;
#pragma unroll
        for (int k = 0; k < IPT_SUB; ++k) {
            if ((unsigned)k < ipt) {
                const unsigned p = base + (unsigned)k * 32u + (unsigned)lane;
                const bool sat = (bs[k] >> lane) & 1u, fail = (bf[k] >> lane) & 1u;
                const unsigned cs = sb + __popc(bs[k] & ltmask) + (sat ? 1u : 0u);   // inclusive ranks
                const unsigned cfl = fb + __popc(bf[k] & ltmask) + (fail ? 1u : 0u);
                if (fail && p < lim) sc.lpos[e.l + cfl - 1] = (unsigned short)p;
                if (sat && p >= lim) sc.rpos[e.l + tot_sat - cs] = (unsigned short)p;
                if (p + 1 == lim) gs.m[off] = cfl;  // false positions left of lim == pairs to swap
                sb += __popc(bs[k]);
                fb += __popc(bf[k]);
            }
        }
        group_bar(g, bid);
        const unsigned m = lim > start ? gs.m[off] : 0u;
        for (unsigned k = gt; k < m; k += G) {
            const unsigned a = sc.lpos[e.l + k], c = sc.rpos[e.l + k];
            const float4 va = sp[a], vc = sp[c];
            sp[a] = vc;
            sp[c] = va;
        }
        group_bar(g, bid);
        if (sweep == 0) {
            lim1 = lim;
            start = lim;
            if (!any_eq) {  // no point equals cutval: the "<=" sweep cannot move anything (group-uniform decision)
                lim2 = lim;
                break;
            }
        } else {
            lim2 = lim;
        }
    }
    unsigned idx;  // :934-936
    if (lim1 > count / 2) idx = lim1;
    else if (lim2 < count / 2) idx = lim2;
    else idx = count / 2;
    float dlow = -INFINITY, dhigh = INFINITY;
    for (unsigned i = gt; i < count; i += G) {
        const float v = sval[4 * i];
        if (i < idx) dlow = fmaxf(dlow, v);
        else dhigh = fminf(dhigh, v);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        dlow = fmaxf(dlow, __shfl_xor_sync(FULL, dlow, m));
        dhigh = fminf(dhigh, __shfl_xor_sync(FULL, dhigh, m));
    }
    if (g > 1) {
        if (lane == 0) {
            gs.red[6 * NW + warp] = dlow;
            gs.red[7 * NW + warp] = dhigh;
        }
        group_bar(g, bid);
        if (gw == 0) {
            dlow = lane < g ? gs.red[6 * NW + off + lane] : -INFINITY;
            dhigh = lane < g ? gs.red[7 * NW + off + lane] : INFINITY;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                dlow = fmaxf(dlow, __shfl_xor_sync(FULL, dlow, m));
                dhigh = fminf(dhigh, __shfl_xor_sync(FULL, dhigh, m));
            }
        }
    }
    if (gt == 0) emit_children_sub(t, sc, e, idx, cf, cv, dlow, dhigh);
    // the group's scratch slots are free for the warps' next nodes only once everybody has read them
    group_bar(g, bid);
}

// Nodes of at most 8*L points, 32/L of them per warp: L lanes (an aligned slice of the warp) split one node, the
// same way a warp group does (rows of L positions, ranks from masked ballots), with no shared scratch at all.  The
// deep levels of a subtree hold hundreds of such nodes; one warp per node would walk them eight rounds per level.
template <int L>
__device__ __forceinline__ void split_small_nodes(const Tree& t, const SubCtx& sc, const Entry* cl,
                                                  const unsigned short* ids, unsigned first, unsigned n_class) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / L, sl = lane % L;
    const bool valid = first + (unsigned)sub < n_class;
    const Entry& e = cl[LIST_BIG + ids[valid ? first + sub : first]];
    const unsigned count = valid ? (unsigned)(e.r - e.l) : 0u;  // idle slices see an empty node and write nothing
    const unsigned gmask = L == 32 ? FULL : (((1u << L) - 1u) << (sub * L));
    const unsigned ltmask = gmask & ((1u << lane) - 1u);
    float4* sp = sc.spp + e.l;

    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = sl; i < count; i += L) {
        const float4 v = sp[i];
        mn[0] = fminf(mn[0], v.x);
        mx[0] = fmaxf(mx[0], v.x);
        mn[1] = fminf(mn[1], v.y);
        mx[1] = fmaxf(mx[1], v.y);
        mn[2] = fminf(mn[2], v.z);
        mx[2] = fmaxf(mx[2], v.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int m = L / 2; m > 0; m >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(FULL, mn[d], m));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL, mx[d], m));
        }
    int cf;
    float cv;
    decide_split(e.lo, e.hi, mn, mx, &cf, &cv);
    const float* sval = reinterpret_cast<const float*>(sp) + cf;

    unsigned start = 0, lim1 = 0, lim2 = 0;
    for (int sweep = 0; sweep < 2; ++sweep) {
        unsigned bs[IPT_SUB], bf[IPT_SUB];
        unsigned tot_sat = 0;
        // rows any slice of the warp still needs (warp-uniform, so the ballots below stay convergent)
        const unsigned rows = __reduce_max_sync(FULL, (count - start + L - 1) / L);
        bool eq = false;
#pragma unroll
        for (int k = 0; k < IPT_SUB; ++k) {
            bs[k] = bf[k] = 0u;
            if ((unsigned)k >= rows) continue;
            const unsigned p = start + (unsigned)(k * L + sl);
            const bool in = p < count;
            const float v = in ? sval[4 * p] : 0.f;
            const bool sat = in && (sweep == 0 ? (v < cv) : (v <= cv));
            eq |= in && v == cv;
            bs[k] = __ballot_sync(FULL, sat) & gmask;
            bf[k] = __ballot_sync(FULL, in && !sat) & gmask;
            tot_sat += __popc(bs[k]);
        }
        const unsigned lim = start + tot_sat;
        unsigned sb = 0, fb = 0, m = 0;
#pragma unroll
        for (int k = 0; k < IPT_SUB; ++k) {
            if ((unsigned)k >= rows) continue;
            const unsigned p = start + (unsigned)(k * L + sl);
            const bool sat = (bs[k] >> lane) & 1u, fail = (bf[k] >> lane) & 1u;
            const unsigned cs = sb + __popc(bs[k] & ltmask) + (sat ? 1u : 0u);
            const unsigned cfl = fb + __popc(bf[k] & ltmask) + (fail ? 1u : 0u);
            const bool left_mis = fail && p < lim;
            if (left_mis) sc.lpos[e.l + cfl - 1] = (unsigned short)p;
            if (sat && p >= lim) sc.rpos[e.l + tot_sat - cs] = (unsigned short)p;
            m += __popc(__ballot_sync(FULL, left_mis) & gmask);
            sb += __popc(bs[k]);
            fb += __popc(bf[k]);
        }
        __syncwarp();
        for (unsigned k = sl; k < m; k += L) {
            const unsigned a = sc.lpos[e.l + k], c = sc.rpos[e.l + k];
            const float4 va = sp[a], vc = sp[c];
            sp[a] = vc;
            sp[c] = va;
        }
        __syncwarp();
        if (sweep == 0) {
            lim1 = lim;
            start = lim;
            if (!__any_sync(FULL, eq)) {  // no node of this warp holds a point equal to its cutval: "<=" sweep is a no-op
                lim2 = lim;
                break;
            }
        } else {
            lim2 = lim;
        }
    }
    unsigned idx;  // :934-936
    if (lim1 > count / 2) idx = lim1;
    else if (lim2 < count / 2) idx = lim2;
    else idx = count / 2;
    float dlow = -INFINITY, dhigh = INFINITY;
    for (unsigned i = sl; i < count; i += L) {
        const float v = sval[4 * i];
        if (i < idx) dlow = fmaxf(dlow, v);
        else dhigh = fminf(dhigh, v);
    }
#pragma unroll
    for (int m = L / 2; m > 0; m >>= 1) {
        dlow = fmaxf(dlow, __shfl_xor_sync(FULL, dlow, m));
        dhigh = fminf(dhigh, __shfl_xor_sync(FULL, dhigh, m));
    }
    if (valid && sl == 0) emit_children_sub(t, sc, e, idx, cf, cv, dlow, dhigh);
    __syncwarp();
}

__device__ __forceinline__ void build_subtree(const Tree& t, unsigned groot, unsigned char* dsm, float* s_red8,
                                              unsigned* s_tot, unsigned* s_m, unsigned* s_eq, unsigned* s_ctl,
                                              unsigned char* s_slot /* [MAX_ROUNDS*NW] */,
                                              unsigned char* s_off /* [LIST_BIG] */) {
    // s_ctl: [0..5] counts of list 0 (big, small, 4 size classes), [6..11] of list 1, [12] nalloc, [13] base node id,
    //        [14] big nodes scheduled so far, [15] rounds of the current schedule
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
        const bool big = count > (unsigned)PER_WARP;
        const int c0 = count > 128u ? 0 : (count > 64u ? 1 : (count > 32u ? 2 : 3));
        for (int i = 0; i < 13; ++i) s_ctl[i] = 0u;
        s_ctl[0] = big ? 1u : 0u;
        s_ctl[1] = big ? 0u : 1u;
        if (!big) {
            s_ctl[2 + c0] = 1u;
            reinterpret_cast<unsigned short*>(dsm + SM_CLS)[c0 * LIST_SMALL] = 0;
        }
        s_ctl[13] = base;
        Entry e;
        e.gid = groot;
        e.l = 0;
        e.r = (unsigned short)count;
        for (int d = 0; d < 3; ++d) {
            e.lo[d] = __ldcg(&t.nlo[(size_t)groot * 3 + d]);
            e.hi[d] = __ldcg(&t.nhi[(size_t)groot * 3 + d]);
        }
        lists[count > (unsigned)PER_WARP ? 0 : LIST_BIG] = e;
    }
    __syncthreads();
    SubCtx sc;
    sc.spp = spp;
    sc.lpos = reinterpret_cast<unsigned short*>(dsm + SM_LPOS);
    sc.rpos = reinterpret_cast<unsigned short*>(dsm + SM_RPOS);
    sc.nalloc = &s_ctl[12];
    sc.base_gid = b * t.cap + s_ctl[13];
    sc.nreserved = 2u * count;
    sc.b = b;
    sc.l0 = l0;
    GroupScratch gs;
    gs.red = s_red8;
    gs.tot = s_tot;
    gs.m = s_m;
    gs.eq = s_eq;
    const bool ok = s_ctl[13] + 2u * count <= t.cap;
    unsigned short* cls = reinterpret_cast<unsigned short*>(dsm + SM_CLS);  // [2][4][LIST_SMALL]
    for (int lvl = 0; ok; ++lvl) {
        const int cur = lvl & 1;
        Entry* cl = lists + (size_t)cur * (LIST_BIG + LIST_SMALL);
        sc.next = lists + (size_t)(cur ^ 1) * (LIST_BIG + LIST_SMALL);
        sc.cnt_next = &s_ctl[(cur ^ 1) * 6];
        sc.cls_next = cls + (size_t)(cur ^ 1) * 4 * LIST_SMALL;
        const unsigned short* ccl = cls + (size_t)cur * 4 * LIST_SMALL;
        const unsigned nbig = min(s_ctl[cur * 6], (unsigned)LIST_BIG);
        const unsigned nsmall = min(s_ctl[cur * 6 + 1], (unsigned)LIST_SMALL);
        if (nbig == 0 && nsmall == 0) break;
        // nodes of more than PER_WARP points: groups of warps, packed into rounds by thread 0
        for (unsigned done = 0; done < nbig;) {
            if (tid == 0) {
                unsigned* w32 = reinterpret_cast<unsigned*>(s_slot);
                for (int i = 0; i < MAX_ROUNDS * NW / 4; ++i) w32[i] = 0xFFFFFFFFu;
                unsigned r = 0, at = 0, j = done;
                for (; j < nbig; ++j) {
                    const unsigned g = (unsigned)group_warps((unsigned)(cl[j].r - cl[j].l));
                    if (at + g > (unsigned)NW || (g == 2u && at == (unsigned)NW - 2u)) {  // (barrier ids stop at 15)
                        ++r;
                        at = 0;
                    }
                    if (r == (unsigned)MAX_ROUNDS) break;
                    s_off[j] = (unsigned char)at;
                    for (unsigned w = at; w < at + g; ++w) s_slot[r * NW + w] = (unsigned char)j;
                    at += g;
                }
                s_ctl[14] = j;
                s_ctl[15] = r < (unsigned)MAX_ROUNDS ? r + 1u : (unsigned)MAX_ROUNDS;
            }
            __syncthreads();
            const unsigned upto = s_ctl[14], rounds = s_ctl[15];
            for (unsigned r = 0; r < rounds; ++r) {
                // a warp may sit in differently shaped groups in consecutive rounds, and a barrier id must never be
                // used with two thread counts at once: rounds are separated CTA-wide (levels rarely need two)
                if (r) __syncthreads();
                const unsigned j = s_slot[r * NW + warp];
                if (j != 0xFFu) split_node_sm(t, sc, cl[j], group_warps((unsigned)(cl[j].r - cl[j].l)), (int)s_off[j], gs);
            }
            done = upto;
            if (done < nbig) __syncthreads();  // the schedule tables are rewritten
        }
        // nodes of at most PER_WARP points: 1, 2, 4 or 8 of them per warp by size class
        if (nsmall) {
            const unsigned n0 = min(s_ctl[cur * 6 + 2], nsmall), n1 = min(s_ctl[cur * 6 + 3], nsmall);
            const unsigned n2 = min(s_ctl[cur * 6 + 4], nsmall), n3 = min(s_ctl[cur * 6 + 5], nsmall);
            const unsigned u0 = n0, u1 = u0 + (n1 + 1) / 2, u2 = u1 + (n2 + 3) / 4, u3 = u2 + (n3 + 7) / 8;
            for (unsigned u = warp; u < u3; u += NW) {
                if (u < u0) split_small_nodes<32>(t, sc, cl, ccl, u, n0);
                else if (u < u1) split_small_nodes<16>(t, sc, cl, ccl + LIST_SMALL, (u - u0) * 2, n1);
                else if (u < u2) split_small_nodes<8>(t, sc, cl, ccl + 2 * LIST_SMALL, (u - u1) * 4, n2);
                else split_small_nodes<4>(t, sc, cl, ccl + 3 * LIST_SMALL, (u - u2) * 8, n3);
            }
        }
        __syncthreads();
        if (tid < 6) s_ctl[cur * 6 + tid] = 0u;
        if (t.tstamps && t.N <= (unsigned)MED_MAX && lvl < 8) mark(nullptr, t.tstamps, 2 + lvl);  // diagnostics
        __syncthreads();
    }
    for (unsigned i = tid; i < count; i += BT) pp[i] = spp[i];
}

__global__ void __launch_bounds__(BT, 1) build_kernel(const float* __restrict__ pts_all, const Tree t) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ unsigned long long s_warp[32];
    __shared__ float s_red[12 * NW];
    __shared__ unsigned s_ctl[16];
    __shared__ unsigned s_tot[NW], s_m[NW], s_eq[NW];
    __shared__ __align__(4) unsigned char s_slot[MAX_ROUNDS * NW];
    __shared__ unsigned char s_off[LIST_BIG];
    const unsigned tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned phase = 0;
    mark(nullptr, t.tstamps, 0);

    // the whole tie path is enqueued without knowing whether any row was flagged: nothing to do in that case
    if (t.n_flag && __ldcg(t.n_flag) == 0) return;
    // a call that shares its support cloud with the previous one (the pyramid's 1-NN up-sampling and the k=16 query of
    // the next level) finds the trees in place; every CTA reads the word before CTA 0 sets it behind the first barrier
    const bool already = t.skip_if_built && __ldcg(t.built) != 0;
    if (already) return;

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
                    t.tmn[(size_t)g * 3 + d] = lo_;  // the root's loose bbox IS the data min/max
                    t.tmx[(size_t)g * 3 + d] = hi_;
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
    if (blockIdx.x == 0 && tid == 0) *t.built = 1u;

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
                split_big(t, __ldcg(cur + i), level, s_warp, s_red);
        }
        grid_sync(t.barrier, phase);
        if (level < 8) mark(nullptr, t.tstamps, 2 + level);
    }
    mark(nullptr, t.tstamps, 10);
    // ---- SUBTREES (independent; no further grid barrier)
    const unsigned nsub = min(__ldcg(&t.list_cnt[MAX_LEVELS + 1]), t.lcap);
    for (unsigned i = blockIdx.x; i < nsub; i += gridDim.x)
        build_subtree(t, __ldcg(t.sublist + i), dyn_smem, s_red, s_tot, s_m, s_eq, s_ctl, s_slot, s_off);
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
// KNNResultSet::addPoint (:72-96) on a register-resident list: strict '>' shifting == the new entry lands behind
// entries of equal distance.  Requires d < rd[KC-1].  Branch-free: 2 FMNMX + 1 FSETP + 2 SEL per slot.
template <int KC>
__device__ __forceinline__ void rs_insert(float (&rd)[KC], unsigned (&ri)[KC], float d, unsigned id) {
    bool p_prev = true;
#pragma unroll
    for (int j = KC - 1; j > 0; --j) {
        const bool p = d < rd[j - 1];
        const float nd = fmaxf(rd[j - 1], fminf(rd[j], d));
        ri[j] = p ? ri[j - 1] : (p_prev ? id : ri[j]);
        rd[j] = nd;
        p_prev = p;
    }
    rd[0] = fminf(rd[0], d);
    ri[0] = p_prev ? id : ri[0];
}

template <typename OutT, int KC>
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

    // result set in registers (KC >= K slots; slots behind K-1 only ever receive entries pushed out of the top K)
    float rd[KC];
    unsigned ri[KC];
#pragma unroll
    for (int j = 0; j < KC; ++j) {
        rd[j] = 3.402823466e+38f;
        ri[j] = 0u;
    }
    int count = 0;
    float worst = 3.402823466e+38f;  // == rd[K-1]; KNNResultSet::init (:47-53)

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
        if (!first && !(mind <= worst)) continue;  // mindistsq*epsError <= worstDist()  (:1319)
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
        const float wsnap = worst;  // snapshot once per leaf (:1277)
#pragma unroll
        for (int k = 0; k < LEAF; ++k) {
            if (k < n) {
                const float dx = __fsub_rn(q[0], buf[k].x), dy = __fsub_rn(q[1], buf[k].y), dz = __fsub_rn(q[2], buf[k].z);
                const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (dist < wsnap) {  // addPoint is called; it stores nothing once the set is full and dist >= worst
                    if (dist < worst) {
                        rs_insert<KC>(rd, ri, dist, __float_as_uint(buf[k].w));
                        float w = rd[0];
#pragma unroll
                        for (int j = 1; j < KC; ++j) w = (j == K - 1) ? rd[j] : w;
                        worst = w;
                    }
                    if (count < K) ++count;
                }
            }
        }
    }
    OutT* o = out + (size_t)row * K;
#pragma unroll
    for (int jj = 0; jj < KC; ++jj)
        if (jj < count) o[jj] = (OutT)ri[jj];
  }
}

enum { TW_BASE = 16 };  // workspace slots TW_BASE.. are owned by this header

// Carve the tree arrays out of workspace slabs and zero the small control block.
static void invalidate_tree_cache(const Ctx* c);

static int alloc_tree(Ctx* c, cudaStream_t s, size_t B, size_t N, Tree* out) {
    invalidate_tree_cache(c);  // the workspaces carved below are the ones a kept tree lives in
    const size_t cap = 3 * N + 64;
    const size_t lcap = B * (N / (LEAF + 1) + 2) + 16;
    SSDR_TRY(c->ws[TW_BASE + 0].reserve(B * N * (sizeof(float4) + 4 * sizeof(unsigned))));
    SSDR_TRY(c->ws[TW_BASE + 1].reserve(B * cap * (sizeof(NodeRec) + 12 * sizeof(float))));
    SSDR_TRY(c->ws[TW_BASE + 2].reserve(3 * lcap * sizeof(unsigned)));
    const size_t ctl_words = 6 * B + B + (size_t)(MAX_LEVELS + 2) + 8 + (B + 3) / 4 + 4 + 8 * B + 4;
    SSDR_TRY(c->ws[TW_BASE + 3].reserve(ctl_words * sizeof(unsigned)));
    const size_t G = (size_t)c->sm_count;
    const size_t grp_words = (size_t)MAX_GROUP_LEVELS * G * 13 + G * 2 * G * 2;
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
    t.tmn = t.nhi + B * cap * 3;
    t.tmx = t.tmn + B * cap * 3;
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
    t.gpart = t.gred + (size_t)MAX_GROUP_LEVELS * G * 12;
    SSDR_CHECK_CUDA(cudaMemsetAsync(t.gbar, 0, (size_t)MAX_GROUP_LEVELS * G * 13 * sizeof(unsigned), s));
    t.root_red = t.error + 4 + (B + 3) / 4 + 1;
    t.item_flags = reinterpret_cast<unsigned char*>(t.error + 4);
    t.built = ctl + ctl_words - 4;
    t.skip_if_built = 0;
    t.n_flag = nullptr;
    t.item_needed = nullptr;
    t.tstamps = nullptr;
    *out = t;
    return SSDR_OK;
}
static unsigned char* tree_needed_flags(const Tree& t) { return t.item_flags; }

static int launch_build(Ctx* c, cudaStream_t s, const float* d_pts, const Tree& t) {
    const size_t smem = SM_TOTAL;
    SSDR_CHECK_CUDA(cudaFuncSetAttribute(build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    SSDR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, build_kernel, BT, smem));
    SSDR_REQUIRE(nb >= 1, SSDR_ERR_CUDA, "tree build kernel does not fit on an SM");
    const float* pts = d_pts;
    Tree tt = t;
    void* args[] = {(void*)&pts, (void*)&tt};
    // one CTA per 2048 points, at least one per item: a small cloud's build keeps to a few SMs (its barriers are
    // cheaper, and the branches of a pyramid that run beside it keep the others)
    size_t grid = ((size_t)t.B * t.N + 2047) / 2048;
    grid = grid < t.B ? t.B : grid;
    grid = grid > (size_t)c->sm_count ? (size_t)c->sm_count : grid;
    SSDR_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)build_kernel, dim3((unsigned)grid), dim3(BT), args, smem, s));
    return SSDR_OK;
}

// ---- tree reuse across calls ------------------------------------------------------------------------------------
// RandLA's pyramid asks for the same support cloud twice in a row (1-NN up-sampling onto level l+1, then the k=16
// self-query of level l+1), and a tree is a pure function of its points: the last trees built by this thread stay
// valid in their workspace until the next build, and a call with the same (B, N) first checks on the device that
// every tree it needs was built and that its position-ordered points are bit-identical to the new cloud
// (pp[i].xyz == pts[pp[i].index] for all i <=> same cloud, because the indices are a permutation).
struct TreeCache {
    const Ctx* owner = nullptr;
    bool valid = false, pending = false;
    size_t B = 0, N = 0;
    Tree t;
};
static thread_local TreeCache g_tree_cache;

static void invalidate_tree_cache(const Ctx* c) {
    if (g_tree_cache.owner == c) g_tree_cache.valid = g_tree_cache.pending = false;
}
// after the call's final synchronisation: the trees enqueued by this call exist iff some row was flagged (the build
// returns at once otherwise) and the build reported no error
static void settle_tree_cache(const Ctx* c, bool trees_exist) {
    if (g_tree_cache.owner == c && g_tree_cache.pending) {
        g_tree_cache.valid = trees_exist;
        g_tree_cache.pending = false;
    }
}

__global__ void verify_tree_kernel(const float* __restrict__ pts_all, const float4* __restrict__ pp, unsigned B,
                                   unsigned N, const unsigned char* __restrict__ needed,
                                   const unsigned char* __restrict__ built, unsigned* __restrict__ mismatch) {
    const unsigned long long total = (unsigned long long)B * N;
    bool bad = false;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned b = (unsigned)(i / N);
        if (!needed[b]) continue;
        if (!built[b]) {
            bad = true;
            continue;
        }
        const float4 v = pp[i];
        const unsigned w = __float_as_uint(v.w);
        if (w >= N) {
            bad = true;
            continue;
        }
        const float* p = pts_all + ((size_t)b * N + w) * 3;
        bad |= __float_as_uint(v.x) != __float_as_uint(__ldg(p)) || __float_as_uint(v.y) != __float_as_uint(__ldg(p + 1)) ||
               __float_as_uint(v.z) != __float_as_uint(__ldg(p + 2));
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(mismatch, 1u);
}

// Enqueue the whole tie path behind the main kernel WITHOUT a host round trip: mark the items that own flagged rows,
// build their trees, overwrite the flagged rows with nanoflann's exact answer.  Every kernel reads the flagged-row
// count on the device and returns at once when it is zero.  The caller checks t_out->error after its final sync and
// then calls settle_tree_cache.  *n_launch is advanced by the kernels launched here.
template <typename OutT>
static int enqueue_tie_path(Ctx* c, cudaStream_t s, const float* d_pts, size_t B, size_t N, const float* d_q, size_t Q,
                            size_t K, OutT* d_out, const unsigned* flag_list, const unsigned* d_flag_count,
                            Tree* t_out, cudaEvent_t ev_mid, unsigned long long* n_launch, bool* reused) {
    SSDR_REQUIRE(K <= (size_t)MAX_K, SSDR_ERR_UNSUPPORTED, "K=%zu > %d in the exact tie path", K, MAX_K);
    Tree t;
    *reused = false;
    TreeCache& tc = g_tree_cache;
    if (tc.valid && tc.owner == c && tc.B == B && tc.N == N) {
        SSDR_TRY(c->ws[TW_BASE + 5].reserve(B + 64));
        unsigned char* scratch = c->ws[TW_BASE + 5].as<unsigned char>();
        unsigned* mismatch = reinterpret_cast<unsigned*>(scratch);
        unsigned char* needed2 = scratch + 16;
        SSDR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, B + 64, s));
        mark_items_kernel<<<64, 256, 0, s>>>(flag_list, d_flag_count, (unsigned)Q, needed2);
        verify_tree_kernel<<<(unsigned)c->sm_count * 2, 256, 0, s>>>(d_pts, tc.t.pp, (unsigned)B, (unsigned)N, needed2,
                                                                    tree_needed_flags(tc.t), mismatch);
        SSDR_CHECK_CUDA(cudaGetLastError());
        *n_launch += 2;
        unsigned h_mismatch = 1;
        SSDR_TRY(d2h_sync(c, &h_mismatch, mismatch, sizeof(unsigned), s));
        if (!h_mismatch) {
            t = tc.t;
            t.n_flag = d_flag_count;
            t.tstamps = nullptr;
            *reused = true;
        } else {
            tc.valid = false;
        }
    }
    if (!*reused) {
        SSDR_TRY(alloc_tree(c, s, B, N, &t));
        unsigned char* needed = tree_needed_flags(t);
        t.item_needed = needed;
        t.n_flag = d_flag_count;
        if (B <= 8) {  // a handful of items: build them all (same latency), so that the next call can reuse every tree
            SSDR_CHECK_CUDA(cudaMemsetAsync(needed, 1, B, s));
        } else {
            mark_items_kernel<<<64, 256, 0, s>>>(flag_list, d_flag_count, (unsigned)Q, needed);
            *n_launch += 1;
        }
        SSDR_TRY(launch_build(c, s, d_pts, t));
        *n_launch += 1;
        tc.owner = c;
        tc.B = B;
        tc.N = N;
        tc.t = t;
        tc.valid = false;
        tc.pending = true;
    }
    if (ev_mid) SSDR_CHECK_CUDA(cudaEventRecord(ev_mid, s));
#define SSDR_EXACT(KCV)                                                                                       \
    exact_query_kernel<OutT, KCV><<<(unsigned)c->sm_count * 16, 32, 0, s>>>(d_q, t, (unsigned)Q, (int)K, flag_list, \
                                                                            d_flag_count, d_out)
    if (K == 1) SSDR_EXACT(1);
    else if (K <= 2) SSDR_EXACT(2);
    else if (K <= 4) SSDR_EXACT(4);
    else if (K <= 8) SSDR_EXACT(8);
    else if (K <= 16) SSDR_EXACT(16);
    else if (K <= 32) SSDR_EXACT(32);
    else SSDR_EXACT(64);
#undef SSDR_EXACT
    SSDR_CHECK_CUDA(cudaGetLastError());
    *n_launch += 1;
    *t_out = t;
    return SSDR_OK;
}

// The same without ANY host round trip, for callers that enqueue several KNN calls back to back (the pyramid): errors
// accumulate in the caller's persistent device word `status`; `shared` carries the tree workspace from one call to the
// next, and reuse_shared says that this call's support cloud IS the previous call's (known structurally, so no content
// check): the build then runs only if the previous call did not need the trees.
template <typename OutT>
static int enqueue_tie_path_async(Ctx* c, cudaStream_t s, const float* d_pts, size_t B, size_t N, const float* d_q,
                                  size_t Q, size_t K, OutT* d_out, const unsigned* flag_list,
                                  const unsigned* d_flag_count, Tree* shared, bool reuse_shared, unsigned* status,
                                  unsigned long long* n_launch) {
    SSDR_REQUIRE(K <= (size_t)MAX_K, SSDR_ERR_UNSUPPORTED, "K=%zu > %d in the exact tie path", K, MAX_K);
    Tree t;
    if (reuse_shared) {
        t = *shared;
        SSDR_REQUIRE(t.B == B && t.N == N, SSDR_ERR_INVALID, "shared tree geometry mismatch");
        t.skip_if_built = 1;
    } else {
        SSDR_TRY(alloc_tree(c, s, B, N, &t));  // also drops the thread's synchronous tree cache
        SSDR_CHECK_CUDA(cudaMemsetAsync(t.item_flags, 1, B, s));
        t.item_needed = nullptr;  // all items: the next call may need any of them
    }
    t.n_flag = d_flag_count;
    t.error = status;
    t.tstamps = nullptr;
    SSDR_TRY(launch_build(c, s, d_pts, t));
    *n_launch += 1;
#define SSDR_EXACT(KCV)                                                                                       \
    exact_query_kernel<OutT, KCV><<<(unsigned)c->sm_count * 16, 32, 0, s>>>(d_q, t, (unsigned)Q, (int)K, flag_list, \
                                                                            d_flag_count, d_out)
    if (K == 1) SSDR_EXACT(1);
    else if (K <= 2) SSDR_EXACT(2);
    else if (K <= 4) SSDR_EXACT(4);
    else if (K <= 8) SSDR_EXACT(8);
    else if (K <= 16) SSDR_EXACT(16);
    else if (K <= 32) SSDR_EXACT(32);
    else SSDR_EXACT(64);
#undef SSDR_EXACT
    SSDR_CHECK_CUDA(cudaGetLastError());
    *n_launch += 1;
    *shared = t;
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
