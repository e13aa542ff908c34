// kdtree.cuh -- nanoflann-identical KD-tree on the device, used to resolve KNN rows whose neighbour order depends on
// nanoflann's tree-visit order (equal / almost equal fp32 distances).
//
// build_kernel reproduces KDTreeSingleIndexAdaptor::buildIndex (utils/nearest_neighbors/nanoflann.hpp:1136-1147,
// divideTree :848-896, middleSplit_ :898-937, planeSplit :948-975) bit for bit -- same vind permutation, same
// divfeat/divlow/divhigh -- as ONE persistent cooperative launch for all batch items:
//   * the tree is grown level by level; a grid barrier separates levels, work lists carry the nodes to split;
//   * a node with more than SMALL_MAX points is split by a whole CTA (block-wide prefix counts over global memory),
//     a smaller node by ONE WARP entirely in shared memory (ballot/popc prefix counts), so deep levels with
//     thousands of nodes use every warp of the GPU;
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
constexpr int IPT = 4;
constexpr int LEAF = 10;
constexpr int SMALL_MAX = 512;
constexpr int MAX_LEVELS = 512;
constexpr int MAX_DEPTH = 96;
constexpr int MAX_K = 64;
constexpr int WARP_SMEM = SMALL_MAX * 4 + SMALL_MAX * 4 + SMALL_MAX + SMALL_MAX;  // sv, sval, sL(u16), sR(u16)

struct __align__(16) NodeRec {
    int c1, c2;            // children (per-item node ids), -1 = leaf          nanoflann.hpp:853-856
    int feat;              // split dimension                                   :877
    unsigned pad;
    float divlow, divhigh; // tight max of left child / min of right child      :886-887
    unsigned l, r;         // vind range [l, r)
};

struct Tree {
    unsigned N, cap, B;    // points per item, node capacity per item (2N+2), items
    unsigned lcap;         // capacity of one work list
    unsigned* vind;        // [B*N]   position -> point index (nanoflann's vind)
    unsigned* lpos;        // [B*N]   big-node scratch: k-th misplaced position from the left, at [l+k]
    unsigned* rpos;        // [B*N]   ... from the right
    unsigned* psat;        // [B*N]   inclusive prefix count of predicate-true positions (from the sweep start)
    unsigned* pfail;       // [B*N]   inclusive prefix count of predicate-false positions
    NodeRec* nodes;        // [B*cap]
    float* nlo;            // [B*cap*3] loose bbox (root bbox cut by the ancestors' planes) -- drives middleSplit_
    float* nhi;
    float* root_lo;        // [B*3] tight root bbox (computeBoundingBox :1241-1263)
    float* root_hi;
    unsigned* node_count;  // [B]
    unsigned* list;        // [2 parities][2 classes][lcap] global node ids (item*cap + node)
    unsigned* list_cnt;    // [(MAX_LEVELS+1)*2]
    unsigned* barrier;     // grid barrier counter
    unsigned* error;       // bit 0: node capacity, bit 1: level cap, bit 2: DFS stack
    const unsigned char* item_needed;  // [B] build only where a flagged row lives (nullptr = all)
};

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

// divideTree's bookkeeping for one split node (:877-892): children records, loose bboxes, next-level work lists
__device__ __forceinline__ void emit_children(const Tree& t, unsigned g, unsigned b, unsigned l, unsigned r,
                                              unsigned idx, int cf, float cv, float divlow, float divhigh,
                                              const float lo[3], const float hi[3], int level) {
    const unsigned a = atomicAdd(&t.node_count[b], 2u);
    if (a + 1 >= t.cap) {
        atomicOr(t.error, 1u);
        return;
    }
    const unsigned ga = b * t.cap + a, gb = ga + 1;
    NodeRec ra, rb;
    ra.c1 = ra.c2 = rb.c1 = rb.c2 = -1;
    ra.feat = rb.feat = 0;
    ra.pad = rb.pad = 0;
    ra.divlow = ra.divhigh = rb.divlow = rb.divhigh = 0.f;
    ra.l = l;
    ra.r = l + idx;
    rb.l = l + idx;
    rb.r = r;
    t.nodes[ga] = ra;
    t.nodes[gb] = rb;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        t.nlo[(size_t)ga * 3 + d] = lo[d];
        t.nhi[(size_t)ga * 3 + d] = (d == cf) ? cv : hi[d];
        t.nlo[(size_t)gb * 3 + d] = (d == cf) ? cv : lo[d];
        t.nhi[(size_t)gb * 3 + d] = hi[d];
    }
    NodeRec me;
    me.c1 = (int)a;
    me.c2 = (int)(a + 1);
    me.feat = cf;
    me.pad = 0;
    me.divlow = divlow;
    me.divhigh = divhigh;
    me.l = l;
    me.r = r;
    t.nodes[g] = me;
    if (level + 1 >= MAX_LEVELS) {
        atomicOr(t.error, 2u);
        return;
    }
    unsigned* next = t.list + (size_t)(((level + 1) & 1) * 2) * t.lcap;
    const unsigned cl = idx, cr = (r - l) - idx;
    if (cl > (unsigned)LEAF) {
        const int cls = cl > (unsigned)SMALL_MAX ? 0 : 1;
        const unsigned pos = atomicAdd(&t.list_cnt[(level + 1) * 2 + cls], 1u);
        next[(size_t)cls * t.lcap + pos] = ga;
    }
    if (cr > (unsigned)LEAF) {
        const int cls = cr > (unsigned)SMALL_MAX ? 0 : 1;
        const unsigned pos = atomicAdd(&t.list_cnt[(level + 1) * 2 + cls], 1u);
        next[(size_t)cls * t.lcap + pos] = gb;
    }
}

// ---- one warp splits one node of <= SMALL_MAX points in shared memory --------------------------------------
__device__ __forceinline__ void split_small(const float* __restrict__ pts_all, const Tree& t, unsigned g, int level,
                                            unsigned char* wsm) {
    const int lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    unsigned* sv = reinterpret_cast<unsigned*>(wsm);
    float* sval = reinterpret_cast<float*>(wsm + SMALL_MAX * 4);
    unsigned short* sL = reinterpret_cast<unsigned short*>(wsm + SMALL_MAX * 8);
    unsigned short* sR = sL + SMALL_MAX / 2;
    const unsigned b = g / t.cap;
    const float* pts = pts_all + (size_t)b * t.N * 3;
    unsigned* vind = t.vind + (size_t)b * t.N;
    const unsigned l = __ldcg(&t.nodes[g].l), r = __ldcg(&t.nodes[g].r);
    const unsigned count = r - l;
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = __ldcg(&t.nlo[(size_t)g * 3 + d]);
        hi[d] = __ldcg(&t.nhi[(size_t)g * 3 + d]);
    }
    const int nslot = (int)((count + 31) / 32);
    // computeMinMax over the node (:827-836)
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int s = 0; s < nslot; ++s) {
        const unsigned p = (unsigned)s * 32 + lane;
        if (p < count) {
            const unsigned v = __ldcg(vind + l + p);
            sv[p] = v;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float x = __ldg(pts + 3 * (size_t)v + d);
                mn[d] = fminf(mn[d], x);
                mx[d] = fmaxf(mx[d], x);
            }
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
    decide_split(lo, hi, mn, mx, &cf, &cv);
    __syncwarp();
    for (int s = 0; s < nslot; ++s) {
        const unsigned p = (unsigned)s * 32 + lane;
        if (p < count) sval[p] = __ldg(pts + 3 * (size_t)sv[p] + cf);
    }
    __syncwarp();
    // planeSplit (:948-975): sweep 0 uses "< cutval" from position 0, sweep 1 "<= cutval" from lim1
    unsigned start = 0, lim1 = 0, lim2 = 0;
    for (int sweep = 0; sweep < 2; ++sweep) {
        unsigned tot = 0;
        for (int s = 0; s < nslot; ++s) {
            const unsigned p = (unsigned)s * 32 + lane;
            const bool in = p < count && p >= start;
            const float v = in ? sval[p] : 0.f;
            const bool sat = in && (sweep == 0 ? (v < cv) : (v <= cv));
            tot += __popc(__ballot_sync(0xffffffffu, sat));
        }
        const unsigned lim = start + tot;
        unsigned sb = 0, fb = 0, m = 0;
        for (int s = 0; s < nslot; ++s) {
            const unsigned p = (unsigned)s * 32 + lane;
            const bool in = p < count && p >= start;
            const float v = in ? sval[p] : 0.f;
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
            const unsigned va = sv[a], vc = sv[c];
            sv[a] = vc;
            sv[c] = va;
            const float fa = sval[a], fc = sval[c];
            sval[a] = fc;
            sval[c] = fa;
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
            const float v = sval[p];
            if (p < idx) dlow = fmaxf(dlow, v);
            else dhigh = fminf(dhigh, v);
            vind[l + p] = sv[p];
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        dlow = fmaxf(dlow, __shfl_xor_sync(0xffffffffu, dlow, m));
        dhigh = fminf(dhigh, __shfl_xor_sync(0xffffffffu, dhigh, m));
    }
    if (lane == 0) emit_children(t, g, b, l, r, idx, cf, cv, dlow, dhigh, lo, hi, level);
    __syncwarp();
}

// ---- one CTA splits one node of > SMALL_MAX points over global memory -----------------------------------------
__device__ __forceinline__ void split_big(const float* __restrict__ pts_all, const Tree& t, unsigned g, int level,
                                          unsigned long long* s_warp, float* s_red, float* s_bc) {
    const unsigned tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned b = g / t.cap;
    const float* pts = pts_all + (size_t)b * t.N * 3;
    unsigned* vind = t.vind + (size_t)b * t.N;
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
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = l + tid; i < r; i += BT) {
        const unsigned v = __ldcg(vind + i);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float x = __ldg(pts + 3 * (size_t)v + d);
            mn[d] = fminf(mn[d], x);
            mx[d] = fmaxf(mx[d], x);
        }
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

    unsigned start = l, lim1 = l, lim2 = l;
    for (int sweep = 0; sweep < 2; ++sweep) {
        unsigned long long carry = 0;
        for (unsigned base = start; base < r; base += BT * IPT) {
            const unsigned i0 = base + tid * IPT;
            unsigned long long f[IPT];
            unsigned long long local = 0;
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const unsigned i = i0 + k;
                unsigned long long fl = 0;
                if (i < r) {
                    const float v = __ldg(pts + 3 * (size_t)__ldcg(vind + i) + cf);
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
            for (int k = 0; k < IPT; ++k) {
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
            const unsigned va = __ldcg(vind + a), vc = __ldcg(vind + c);
            vind[a] = vc;
            vind[c] = va;
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
    unsigned idx;
    if (l1 > count / 2) idx = l1;
    else if (l2 < count / 2) idx = l2;
    else idx = count / 2;
    float dlow = -INFINITY, dhigh = INFINITY;
    for (unsigned i = l + tid; i < r; i += BT) {
        const float v = __ldg(pts + 3 * (size_t)__ldcg(vind + i) + cf);
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
        emit_children(t, g, b, l, r, idx, cf, cv, dlow, dhigh, lo, hi, level);
    }
}

__global__ void __launch_bounds__(BT, 1) build_kernel(const float* __restrict__ pts_all, const Tree t) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ unsigned long long s_warp[32];
    __shared__ float s_red[6 * NW];
    __shared__ float s_bc[4];
    const unsigned tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned phase = 0;

    // ---- roots: vind = identity (init_vind :1232-1238), data bbox (computeBoundingBox :1241-1263)
    for (unsigned b = blockIdx.x; b < t.B; b += gridDim.x) {
        if (t.item_needed && !t.item_needed[b]) continue;
        const float* pts = pts_all + (size_t)b * t.N * 3;
        unsigned* vind = t.vind + (size_t)b * t.N;
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (unsigned i = tid; i < t.N; i += BT) {
            vind[i] = i;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float x = __ldg(pts + 3 * (size_t)i + d);
                mn[d] = fminf(mn[d], x);
                mx[d] = fmaxf(mx[d], x);
            }
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
        if (tid == 0) {
            const unsigned g = b * t.cap;
            for (int d = 0; d < 3; ++d) {
                float a = s_red[d * NW], c = s_red[(3 + d) * NW];
                for (int w = 1; w < NW; ++w) {
                    a = fminf(a, s_red[d * NW + w]);
                    c = fmaxf(c, s_red[(3 + d) * NW + w]);
                }
                t.nlo[(size_t)g * 3 + d] = a;
                t.nhi[(size_t)g * 3 + d] = c;
                t.root_lo[b * 3 + d] = a;
                t.root_hi[b * 3 + d] = c;
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
            if (t.N > (unsigned)LEAF) {
                const int cls = t.N > (unsigned)SMALL_MAX ? 0 : 1;
                const unsigned pos = atomicAdd(&t.list_cnt[cls], 1u);
                t.list[(size_t)cls * t.lcap + pos] = g;
            }
        }
        __syncthreads();
    }
    grid_sync(t.barrier, phase);

    for (int level = 0; level < MAX_LEVELS; ++level) {
        const unsigned nbig = __ldcg(&t.list_cnt[level * 2]), nsmall = __ldcg(&t.list_cnt[level * 2 + 1]);
        if (nbig == 0 && nsmall == 0) break;
        const unsigned* cur = t.list + (size_t)((level & 1) * 2) * t.lcap;
        for (unsigned i = blockIdx.x; i < nbig; i += gridDim.x)
            split_big(pts_all, t, __ldcg(cur + i), level, s_warp, s_red, s_bc);
        for (unsigned w = blockIdx.x * NW + warp; w < nsmall; w += gridDim.x * NW)
            split_small(pts_all, t, __ldcg(cur + t.lcap + w), level, dyn_smem + (size_t)warp * WARP_SMEM);
        grid_sync(t.barrier, phase);
    }
}

__global__ void mark_items_kernel(const unsigned* __restrict__ flag_list, unsigned n_flag, unsigned Q,
                                  unsigned char* __restrict__ item_needed) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_flag) item_needed[flag_list[i] / Q] = 1;
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
template <typename OutT>
__global__ void __launch_bounds__(32) exact_query_kernel(const float* __restrict__ pts_all,
                                                         const float* __restrict__ q_all, const Tree t, unsigned Q,
                                                         int K, const unsigned* __restrict__ flag_list,
                                                         unsigned n_flag, OutT* __restrict__ out) {
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_flag) return;
    const unsigned row = flag_list[f];
    const unsigned b = row / Q;
    const unsigned N = t.N;
    const float* pts = pts_all + (size_t)b * N * 3;
    const unsigned* vind = t.vind + (size_t)b * N;
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
        const float worst = rd[K - 1];  // snapshot once per leaf (:1277)
        for (unsigned i = nd.l; i < nd.r; ++i) {
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

// Carve the tree arrays out of workspace slabs and zero the small control block.
static int alloc_tree(Ctx* c, cudaStream_t s, size_t B, size_t N, Tree* out) {
    const size_t cap = 2 * N + 2;
    const size_t lcap = B * (N / 5 + 2) + 16;  // a level never holds more split-able nodes than that
    SSDR_TRY(c->ws[TW_BASE + 0].reserve(5 * B * N * sizeof(unsigned)));
    SSDR_TRY(c->ws[TW_BASE + 1].reserve(B * cap * (sizeof(NodeRec) + 6 * sizeof(float))));
    SSDR_TRY(c->ws[TW_BASE + 2].reserve(4 * lcap * sizeof(unsigned)));
    const size_t ctl_words = 6 * B + B + (size_t)(MAX_LEVELS + 1) * 2 + 8 + (B + 3) / 4 + 4;
    SSDR_TRY(c->ws[TW_BASE + 3].reserve(ctl_words * sizeof(unsigned)));
    Tree t;
    t.N = (unsigned)N;
    t.cap = (unsigned)cap;
    t.B = (unsigned)B;
    t.lcap = (unsigned)lcap;
    unsigned* pb = c->ws[TW_BASE + 0].as<unsigned>();
    t.vind = pb;
    t.lpos = pb + B * N;
    t.rpos = pb + 2 * B * N;
    t.psat = pb + 3 * B * N;
    t.pfail = pb + 4 * B * N;
    t.nodes = c->ws[TW_BASE + 1].as<NodeRec>();
    t.nlo = reinterpret_cast<float*>(t.nodes + B * cap);
    t.nhi = t.nlo + B * cap * 3;
    t.list = c->ws[TW_BASE + 2].as<unsigned>();
    unsigned* ctl = c->ws[TW_BASE + 3].as<unsigned>();
    SSDR_CHECK_CUDA(cudaMemsetAsync(ctl, 0, ctl_words * sizeof(unsigned), s));
    t.root_lo = reinterpret_cast<float*>(ctl);
    t.root_hi = reinterpret_cast<float*>(ctl + 3 * B);
    t.node_count = ctl + 6 * B;
    t.list_cnt = ctl + 7 * B;
    t.barrier = t.list_cnt + (size_t)(MAX_LEVELS + 1) * 2;
    t.error = t.barrier + 4;
    t.item_needed = nullptr;
    *out = t;
    return SSDR_OK;
}
static unsigned char* tree_needed_flags(const Tree& t) { return reinterpret_cast<unsigned char*>(t.error + 4); }

static int launch_build(Ctx* c, cudaStream_t s, const float* d_pts, const Tree& t) {
    const size_t smem = (size_t)NW * WARP_SMEM;
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
                 "than %d)", h_err, MAX_LEVELS, MAX_DEPTH);
    return SSDR_OK;
}

// Build the trees of the items that own flagged rows, then overwrite those rows with nanoflann's exact answer.
template <typename OutT>
static int resolve_flagged(Ctx* c, cudaStream_t s, const float* d_pts, size_t B, size_t N, const float* d_q, size_t Q,
                           size_t K, OutT* d_out, const unsigned* flag_list, unsigned n_flag,
                           unsigned long long* builds) {
    SSDR_REQUIRE(K <= (size_t)MAX_K, SSDR_ERR_UNSUPPORTED, "K=%zu > %d in the exact tie path", K, MAX_K);
    Tree t;
    SSDR_TRY(alloc_tree(c, s, B, N, &t));
    unsigned char* needed = tree_needed_flags(t);
    t.item_needed = needed;
    mark_items_kernel<<<(n_flag + 255) / 256, 256, 0, s>>>(flag_list, n_flag, (unsigned)Q, needed);
    SSDR_TRY(launch_build(c, s, d_pts, t));
    exact_query_kernel<OutT><<<(n_flag + 31) / 32, 32, 0, s>>>(d_pts, d_q, t, (unsigned)Q, (int)K, flag_list, n_flag,
                                                              d_out);
    SSDR_CHECK_CUDA(cudaGetLastError());
    SSDR_TRY(check_tree_error(c, s, t));
    if (builds) *builds = B;  // upper bound; items without flagged rows are skipped inside the kernel
    return SSDR_OK;
}

}  // namespace kdtree
}  // namespace ssdr
