// grid.cu -- voxel grid (barycentre) subsampling on sm_100a.
//
// Replaces grid_subsampling() (utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106),
// a single-threaded unordered_map loop, by a sort + segmented sequential reduce that reproduces the reference
// bit for bit:
//   1. min/max corners                         (cloud.cpp:27-67)                     -> minmax_kernel
//   2. origin = floor(min * (1/dl)) * dl, nX, nY (grid_subsampling.cpp:27-31)        -> setup_kernel
//   3. key = iX + nX*iY + nX*nY*iZ with i* = floor((p-origin)/dl) in IEEE fp32, true division
//      (grid_subsampling.cpp:53-56)                                                  -> key_kernel
//   4. STABLE LSD radix sort of (key, input index) over the significant key bits only -> prim::radix_sort_pairs
//   5. segment heads -> exclusive scan -> voxel start offsets, M                     -> head_flag/voxel_start kernels
//   6. one thread per voxel walks its points IN INPUT ORDER (the stable sort keeps ascending index inside a
//      voxel), accumulating fp32 sums exactly like SampledData::update_* (grid_subsampling.h:42-79), then
//      bary = sum * float(1.0/count), feat = sum / float(count) (grid_subsampling.cpp:87-95) and the label vote
//      with libstdc++'s unordered_map iteration order as tie-break (grid_subsampling.cpp:97-102)  -> reduce_kernel
//   7. optional (SSDR_GRID_ORDER_REFERENCE): rows permuted into the reference's libstdc++ hash-iteration order.
// Rows come out in ascending voxel-key order by default (SSDR_GRID_ORDER_KEY).
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
    unsigned long long max_key;  // largest key seen (key_kernel) -> number of radix bits worth sorting
    unsigned long long n_sel;    // points taking part (all of them, or the ones inside the requested slab)
    unsigned long long M;
};

enum { WS_META = 0, WS_PART = 1, WS_KEYS = 2, WS_KEYS2 = 3, WS_IDX = 4, WS_IDX2 = 5, WS_TEMP = 6, WS_STARTS = 7,
       WS_IN_P = 9, WS_IN_F = 10, WS_IN_C = 11, WS_HASH = 12, WS_CMIN = 13, WS_SEL = 14, WS_RAW = 15 };

constexpr int LABEL_CAP = 64;
constexpr int MM_BLOCK = 256;

// ---- 1. min / max ------------------------------------------------------------------------------------
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

// float -> size_t the way x86-64 gcc does it for |v| < 2^63 (truncate to int64, reinterpret)
__device__ __forceinline__ unsigned long long f2u64(float v) { return (unsigned long long)(long long)v; }

static int bits_for_value(unsigned long long v) {  // radix bits needed to order values <= v
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

// ---- 2. grid geometry ---------------------------------------------------------------------------------
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
        Meta m;
        const float inv = __fdiv_rn(1.0f, dl);
        for (int d = 0; d < 3; ++d) {
            m.mn[d] = r[d];
            m.mx[d] = r[3 + d];
            m.origin[d] = __fmul_rn(floorf(__fmul_rn(r[d], inv)), dl);
        }
        m.dl = dl;
        m.nX = f2u64(floorf(__fdiv_rn(__fsub_rn(m.mx[0], m.origin[0]), dl))) + 1ull;
        m.nY = f2u64(floorf(__fdiv_rn(__fsub_rn(m.mx[1], m.origin[1]), dl))) + 1ull;
        m.nZ = f2u64(floorf(__fdiv_rn(__fsub_rn(m.mx[2], m.origin[2]), dl))) + 1ull;
        m.error = 0;
        m.key_bits = 64;
        m.max_key = 0;
        m.n_sel = 0;
        m.M = 0;
        *meta = m;
    }
}

// ---- 3. voxel keys --------------------------------------------------------------------------------------
// Keys are computed in wrapping 64-bit arithmetic exactly like the reference's size_t expression, so even
// degenerate inputs (a point rounding to one cell below the origin, overflowing nX*nY*nZ) group identically.
__global__ void key_kernel(const float* __restrict__ pts, unsigned long long N, Meta* __restrict__ meta,
                           unsigned long long* __restrict__ keys, unsigned* __restrict__ idx,
                           const unsigned* __restrict__ sel /* nullable: slab members, meta->n_sel of them */) {
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = sel ? meta->n_sel : N;
    unsigned long long key = 0;
    if (j < n) {
        const unsigned long long i = sel ? sel[j] : j;
        const float dl = meta->dl;
        const unsigned long long iX = f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + 0), meta->origin[0]), dl)));
        const unsigned long long iY = f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + 1), meta->origin[1]), dl)));
        const unsigned long long iZ = f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + 2), meta->origin[2]), dl)));
        key = iX + meta->nX * iY + meta->nX * meta->nY * iZ;
        keys[j] = key;
        idx[j] = (unsigned)i;
    }
    unsigned long long mk = key;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, mk, m);
        mk = o > mk ? o : mk;
    }
    __shared__ unsigned long long smax[8];
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mk;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mk = smax[w] > mk ? smax[w] : mk;
        atomicMax(&meta->max_key, mk);
    }
}

// ---- slabs: a rank of a multi-GPU job subsamples only the voxel layers [lo, hi) along one axis, with the grid
// geometry of the WHOLE cloud, so every voxel is owned by exactly one rank and its value is bit-identical -----------
__global__ void slab_flag_kernel(const float* __restrict__ pts, unsigned long long N, const Meta* __restrict__ meta,
                                 int axis, unsigned long long lo, unsigned long long hi, unsigned* __restrict__ flags) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned long long layer =
        f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + axis), meta->origin[axis]), meta->dl)));
    flags[i] = (layer >= lo && layer < hi) ? 1u : 0u;
}
__global__ void slab_compact_kernel(const unsigned* __restrict__ flags, const unsigned* __restrict__ pos,
                                    unsigned long long N, unsigned* __restrict__ sel) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && flags[i]) sel[pos[i]] = (unsigned)i;
}
__global__ void point_layers_kernel(const float* __restrict__ pts, unsigned long long N,
                                    const Meta* __restrict__ meta, int axis, int* __restrict__ layers) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned long long layer =
        f2u64(floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + axis), meta->origin[axis]), meta->dl)));
    layers[i] = layer < 0x7FFFFFFFull ? (int)layer : 0x7FFFFFFF;
}
__global__ void bbox_fill_kernel(float* partials, float a0, float a1, float a2, float b0, float b1, float b2) {
    partials[0] = a0; partials[1] = a1; partials[2] = a2;
    partials[3] = b0; partials[4] = b1; partials[5] = b2;
}

// widen uint8 colours / labels on the device (exact in float32 / int32): the callers' arrays are uint8 and the
// reference converts them on the host (wrapper.cpp:100-106); uploading the bytes is 4x less PCIe and no host pass
__global__ void widen_u8_f32_kernel(const unsigned char* __restrict__ in, unsigned long long n, float* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
__global__ void widen_u8_i32_kernel(const unsigned char* __restrict__ in, unsigned long long n, int* __restrict__ out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

// ---- 5. voxel segments of the sorted keys ---------------------------------------------------------------------
__global__ void head_flag_kernel(const unsigned long long* __restrict__ keys, unsigned long long N,
                                 unsigned* __restrict__ flags) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void voxel_start_kernel(const unsigned* __restrict__ flags, const unsigned* __restrict__ vid,
                                   unsigned long long N, unsigned* __restrict__ starts) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && flags[i]) starts[vid[i]] = (unsigned)i;
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

// ---- 6. per-voxel sequential reduce -----------------------------------------------------------------------
template <int FD>  // FD = compile-time feature width, -1 = generic (accumulate in the output row)
__global__ void reduce_kernel(const float* __restrict__ pts, const float* __restrict__ feats,
                              const int* __restrict__ cls, int fdim, int ldim, const unsigned long long* __restrict__ keys,
                              const unsigned* __restrict__ idx, const unsigned* __restrict__ starts,
                              unsigned long long N, unsigned long long M, float* __restrict__ out_p,
                              float* __restrict__ out_f, int* __restrict__ out_c,
                              unsigned long long* __restrict__ out_k, int* __restrict__ out_n, Meta* meta) {
    const unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    const unsigned long long s = starts[v];
    const unsigned long long e = v + 1 < M ? (unsigned long long)starts[v + 1] : N;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    float fs[FD > 0 ? FD : 1];
#pragma unroll
    for (int j = 0; j < (FD > 0 ? FD : 1); ++j) fs[j] = 0.f;
    if (FD < 0)
        for (int j = 0; j < fdim; ++j) out_f[v * fdim + j] = 0.f;
    int labs[LABEL_CAP], cnts[LABEL_CAP];
    int nl = 0;
    bool overflow = false;
    for (unsigned long long i = s; i < e; ++i) {
        const unsigned long long p = idx[i];
        sx = __fadd_rn(sx, __ldg(pts + 3 * p + 0));
        sy = __fadd_rn(sy, __ldg(pts + 3 * p + 1));
        sz = __fadd_rn(sz, __ldg(pts + 3 * p + 2));
        if (FD > 0) {
#pragma unroll
            for (int j = 0; j < (FD > 0 ? FD : 1); ++j) fs[j] = __fadd_rn(fs[j], __ldg(feats + p * FD + j));
        } else if (FD < 0) {
            for (int j = 0; j < fdim; ++j)
                out_f[v * fdim + j] = __fadd_rn(out_f[v * fdim + j], __ldg(feats + p * fdim + j));
        }
        if (ldim >= 1) {  // first label column in the same pass
            const int lab = __ldg(cls + p * ldim);
            int q = 0;
            while (q < nl && labs[q] != lab) ++q;
            if (q == nl) {
                if (nl < LABEL_CAP) {
                    labs[nl] = lab;
                    cnts[nl] = 0;
                    ++nl;
                } else {
                    overflow = true;
                    q = 0;
                }
            }
            cnts[q] += 1;
        }
    }
    const int count = (int)(e - s);
    const float a = (float)(1.0 / (double)count);  // grid_subsampling.cpp:87: double reciprocal narrowed to float
    out_p[3 * v + 0] = __fmul_rn(sx, a);
    out_p[3 * v + 1] = __fmul_rn(sy, a);
    out_p[3 * v + 2] = __fmul_rn(sz, a);
    const float cf = (float)count;
    if (FD > 0) {
#pragma unroll
        for (int j = 0; j < (FD > 0 ? FD : 1); ++j) out_f[v * FD + j] = __fdiv_rn(fs[j], cf);
    } else if (FD < 0) {
        for (int j = 0; j < fdim; ++j) out_f[v * fdim + j] = __fdiv_rn(out_f[v * fdim + j], cf);
    }
    for (int col = 0; col < ldim; ++col) {
        if (col > 0) {  // further label columns: one more walk each (rare: ldim is 1 for every reference caller)
            nl = 0;
            for (unsigned long long i = s; i < e; ++i) {
                const int lab = __ldg(cls + (unsigned long long)idx[i] * ldim + col);
                int q = 0;
                while (q < nl && labs[q] != lab) ++q;
                if (q == nl) {
                    if (nl < LABEL_CAP) {
                        labs[nl] = lab;
                        cnts[nl] = 0;
                        ++nl;
                    } else {
                        overflow = true;
                        q = 0;
                    }
                }
                cnts[q] += 1;
            }
        }
        int best = -1, nbest = 0, arg = 0;
        for (int q = 0; q < nl; ++q) {
            if (cnts[q] > best) {
                best = cnts[q];
                nbest = 1;
                arg = q;
            } else if (cnts[q] == best)
                ++nbest;
        }
        out_c[v * ldim + col] = nbest == 1 ? labs[arg] : label_first_in_iteration_order(labs, cnts, nl, best);
    }
    if (overflow) meta->error = 2;
    out_k[v] = keys[s];
    out_n[v] = count;
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

static int bits_for_value(unsigned long long v);

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

struct Slab {
    int axis = -1;  // -1: whole cloud
    unsigned long long lo = 0, hi = 0;
};

static int run_dev(Ctx* c, cudaStream_t s, const float* d_p, const float* d_f, const int* d_c, size_t N, size_t fdim,
                   size_t ldim, float dl, int order, size_t* M_out, void** handle, Slab slab = Slab(),
                   const float* bbox = nullptr) {
    typedef unsigned long long KeyT;
    SSDR_REQUIRE(d_p && M_out && handle, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_EMPTY, "Error");
    SSDR_REQUIRE(N < 0x7FFFFFFFull, SSDR_ERR_UNSUPPORTED, "N=%zu exceeds 2^31-2 points per call", N);
    SSDR_REQUIRE(order == SSDR_GRID_ORDER_KEY || order == SSDR_GRID_ORDER_REFERENCE, SSDR_ERR_INVALID,
                 "order must be SSDR_GRID_ORDER_KEY or SSDR_GRID_ORDER_REFERENCE");
    SSDR_REQUIRE(dl > 0.0f, SSDR_ERR_INVALID, "sampleDl must be positive");
    if (!d_f) fdim = 0;
    if (!d_c) ldim = 0;
    SSDR_TRY(c->ws[WS_META].reserve(sizeof(Meta)));
    SSDR_TRY(c->ws[WS_KEYS].reserve(N * sizeof(KeyT)));
    SSDR_TRY(c->ws[WS_KEYS2].reserve(N * sizeof(KeyT)));
    SSDR_TRY(c->ws[WS_IDX].reserve(N * sizeof(unsigned)));
    SSDR_TRY(c->ws[WS_IDX2].reserve(N * sizeof(unsigned)));
    SSDR_TRY(c->ws[WS_STARTS].reserve((N + 1) * sizeof(unsigned)));
    Meta* meta = c->ws[WS_META].as<Meta>();
    KeyT* keys = c->ws[WS_KEYS].as<KeyT>();
    KeyT* keys2 = c->ws[WS_KEYS2].as<KeyT>();
    unsigned* idx = c->ws[WS_IDX].as<unsigned>();
    unsigned* idx2 = c->ws[WS_IDX2].as<unsigned>();
    unsigned* starts = c->ws[WS_STARTS].as<unsigned>();

    SSDR_TRY(c->ws[WS_TEMP].reserve((prim::rs_scratch_words(N) + prim::scan_scratch_words(N) + 8) * sizeof(unsigned)));
    unsigned* scratch = c->ws[WS_TEMP].as<unsigned>();
    SSDR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (prim::rs_scratch_words(N) + prim::scan_scratch_words(N) + 8) * sizeof(unsigned), s));
    SSDR_TRY(geometry(c, s, d_p, N, dl, bbox, meta));
    const unsigned* sel = nullptr;
    if (slab.axis >= 0) {  // members of the slab, in input order (stable compaction)
        SSDR_REQUIRE(slab.axis < 3 && slab.lo <= slab.hi, SSDR_ERR_INVALID, "bad slab");
        SSDR_TRY(c->ws[WS_SEL].reserve(N * sizeof(unsigned)));
        unsigned* flags = reinterpret_cast<unsigned*>(keys2);  // free until the sort
        unsigned* pos = flags + N;
        const unsigned nb0 = (unsigned)((N + 255) / 256);
        slab_flag_kernel<<<nb0, 256, 0, s>>>(d_p, N, meta, slab.axis, slab.lo, slab.hi, flags);
        SSDR_TRY(prim::exclusive_scan_u32(flags, pos, N, scratch + prim::rs_scratch_words(N),
                                          reinterpret_cast<unsigned*>(&meta->n_sel), s));
        slab_compact_kernel<<<nb0, 256, 0, s>>>(flags, pos, N, c->ws[WS_SEL].as<unsigned>());
        sel = c->ws[WS_SEL].as<unsigned>();
    }
    key_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(d_p, N, meta, keys, idx, sel);
    SSDR_CHECK_CUDA(cudaGetLastError());
    Meta hm;
    SSDR_TRY(d2h_sync(c, &hm, meta, sizeof(Meta), s));  // sync 1: how many key bits are worth sorting (and slab size)
    const int key_bits = bits_for_value(hm.max_key);
    if (sel) {
        N = (size_t)hm.n_sel;  // everything below works on the slab's points only
        if (N == 0) {          // an empty slab is a valid shard: zero rows
            Handle* h0 = new Handle();
            h0->stream = s;
            h0->device = c->device;
            h0->fdim = fdim;
            h0->ldim = ldim;
            *M_out = 0;
            *handle = h0;
            return SSDR_OK;
        }
    }

    // 4. stable radix sort over the significant bits; 5. heads -> voxel ids -> starts (M lands in meta->M)
    int cur = 0;
    SSDR_TRY(prim::radix_sort_pairs(keys, idx, keys2, idx2, N, key_bits, scratch, &cur, s));
    const KeyT* keys_sorted = cur ? keys2 : keys;
    const unsigned* idx_sorted = cur ? idx2 : idx;
    unsigned* flags = reinterpret_cast<unsigned*>(cur ? keys : keys2);  // the other key buffer is free now: 8 B/pt
    unsigned* vid = flags + N;
    const unsigned nb = (unsigned)((N + 255) / 256);
    head_flag_kernel<<<nb, 256, 0, s>>>(keys_sorted, N, flags);
    SSDR_TRY(prim::exclusive_scan_u32(flags, vid, N, scratch + prim::rs_scratch_words(N),
                                      reinterpret_cast<unsigned*>(&meta->M), s));
    voxel_start_kernel<<<nb, 256, 0, s>>>(flags, vid, N, starts);
    SSDR_CHECK_CUDA(cudaGetLastError());
    SSDR_TRY(d2h_sync(c, &hm, meta, sizeof(Meta), s));  // sync 2: M, to size the outputs
    const size_t M = (size_t)hm.M;
    SSDR_REQUIRE(M >= 1 && M <= N, SSDR_ERR_EMPTY, "Error");

    Handle* h = new Handle();
    h->stream = s;
    h->device = c->device;
    h->M = M;
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
    SSDR_ALLOC(h->d_p, M * 3 * sizeof(float));
    if (fdim) SSDR_ALLOC(h->d_f, M * fdim * sizeof(float));
    if (ldim) SSDR_ALLOC(h->d_c, M * ldim * sizeof(int));
    SSDR_ALLOC(h->d_k, M * sizeof(unsigned long long));
    SSDR_ALLOC(h->d_n, M * sizeof(int));
#undef SSDR_ALLOC
    const unsigned vblocks = (unsigned)((M + 127) / 128);
#define SSDR_REDUCE(FDV)                                                                                         \
    reduce_kernel<FDV><<<vblocks, 128, 0, s>>>(d_p, d_f, d_c, (int)fdim, (int)ldim, keys_sorted, idx_sorted, \
                                               starts, N, M, h->d_p, h->d_f, h->d_c, h->d_k, h->d_n, meta)
    switch (fdim) {
        case 0: SSDR_REDUCE(0); break;
        case 1: SSDR_REDUCE(1); break;
        case 2: SSDR_REDUCE(2); break;
        case 3: SSDR_REDUCE(3); break;
        case 4: SSDR_REDUCE(4); break;
        case 6: SSDR_REDUCE(6); break;
        case 8: SSDR_REDUCE(8); break;
        default: SSDR_REDUCE(-1); break;
    }
#undef SSDR_REDUCE
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(set_error(SSDR_ERR_CUDA, "reduce_kernel launch failed: %s", cudaGetErrorString(e)));
    }
    if (ldim) {  // the only late failure mode is a label-table overflow; surface it before results are used
        int rc = d2h_sync(c, &hm, meta, sizeof(Meta), s);
        if (rc != SSDR_OK) return fail(rc);
        if (hm.error == 2)
            return fail(set_error(SSDR_ERR_UNSUPPORTED, "more than %d distinct labels inside one voxel", LABEL_CAP));
    }
    if (order == SSDR_GRID_ORDER_REFERENCE && M > 1) {
        int rc = reorder_reference(c, s, h, idx_sorted, starts, N, scratch);
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
    return grid::run_dev(c, (cudaStream_t)stream, d_points, d_feats, d_classes, N, fdim, ldim,
                         sampleDl, order, M_out, handle);
}

int ssdr_grid_subsample_slab_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N,
                                 size_t fdim, size_t ldim, float sampleDl, int order, const float* bbox, int axis,
                                 unsigned long long layer_lo, unsigned long long layer_hi, void* stream,
                                 size_t* M_out, void** handle) {
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    SSDR_REQUIRE(axis >= -1 && axis <= 2, SSDR_ERR_INVALID, "axis must be -1 (no slab), 0, 1 or 2");
    SSDR_REQUIRE(order == SSDR_GRID_ORDER_KEY || (axis < 0 && !bbox), SSDR_ERR_UNSUPPORTED,
                 "the reference's hash-iteration order is defined for a whole cloud only; slabs come in key order");
    grid::Slab slab;
    slab.axis = axis;
    slab.lo = layer_lo;
    slab.hi = layer_hi;
    return grid::run_dev(c, (cudaStream_t)stream, d_points, d_feats, d_classes, N, fdim, ldim, sampleDl, order,
                         M_out, handle, slab, bbox);
}

int ssdr_grid_bbox_dev(const float* d_points, size_t N, void* stream, float* bbox_out) {
    SSDR_REQUIRE(d_points && bbox_out, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_EMPTY, "Error");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
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

int ssdr_grid_subsample(const float* points, const float* feats, const int32_t* classes, size_t N, size_t fdim,
                        size_t ldim, float sampleDl, int order, size_t* M_out, void** handle) {
    SSDR_REQUIRE(points && M_out && handle, SSDR_ERR_INVALID, "NULL pointer");
    SSDR_REQUIRE(N >= 1, SSDR_ERR_EMPTY, "Error");
    Ctx* c;
    SSDR_TRY(get_ctx(&c));
    if (!feats) fdim = 0;
    if (!classes) ldim = 0;
    SSDR_TRY(c->ws[grid::WS_IN_P].reserve(N * 3 * sizeof(float)));
    SSDR_TRY(h2d(c, c->ws[grid::WS_IN_P].p, points, N * 3 * sizeof(float), c->stream));
    const float* d_f = nullptr;
    const int* d_c = nullptr;
    if (fdim) {
        SSDR_TRY(c->ws[grid::WS_IN_F].reserve(N * fdim * sizeof(float)));
        SSDR_TRY(h2d(c, c->ws[grid::WS_IN_F].p, feats, N * fdim * sizeof(float), c->stream));
        d_f = c->ws[grid::WS_IN_F].as<float>();
    }
    if (ldim) {
        SSDR_TRY(c->ws[grid::WS_IN_C].reserve(N * ldim * sizeof(int)));
        SSDR_TRY(h2d(c, c->ws[grid::WS_IN_C].p, classes, N * ldim * sizeof(int), c->stream));
        d_c = c->ws[grid::WS_IN_C].as<int>();
    }
    return grid::run_dev(c, c->stream, c->ws[grid::WS_IN_P].as<float>(), d_f, d_c, N, fdim, ldim, sampleDl, order,
                         M_out, handle);
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
    SSDR_TRY(c->ws[grid::WS_IN_P].reserve(N * 3 * sizeof(float)));
    SSDR_TRY(h2d(c, c->ws[grid::WS_IN_P].p, points, N * 3 * sizeof(float), s));
    const size_t nf = N * fdim, nc = N * ldim;
    const size_t raw_f = (fdim && feats_dtype == SSDR_DTYPE_U8) ? align_up(nf, 256) : 0;
    const size_t raw_c = (ldim && classes_dtype == SSDR_DTYPE_U8) ? align_up(nc, 256) : 0;
    if (raw_f + raw_c) SSDR_TRY(c->ws[grid::WS_RAW].reserve(raw_f + raw_c));
    unsigned char* raw = c->ws[grid::WS_RAW].as<unsigned char>();
    const float* d_f = nullptr;
    const int* d_c = nullptr;
    if (fdim) {
        SSDR_TRY(c->ws[grid::WS_IN_F].reserve(nf * sizeof(float)));
        if (raw_f) {
            SSDR_TRY(h2d(c, raw, feats, nf, s));
            grid::widen_u8_f32_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, s>>>(raw, nf, c->ws[grid::WS_IN_F].as<float>());
        } else {
            SSDR_TRY(h2d(c, c->ws[grid::WS_IN_F].p, feats, nf * sizeof(float), s));
        }
        d_f = c->ws[grid::WS_IN_F].as<float>();
    }
    if (ldim) {
        SSDR_TRY(c->ws[grid::WS_IN_C].reserve(nc * sizeof(int)));
        if (raw_c) {
            SSDR_TRY(h2d(c, raw + raw_f, classes, nc, s));
            grid::widen_u8_i32_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(raw + raw_f, nc, c->ws[grid::WS_IN_C].as<int>());
        } else {
            SSDR_TRY(h2d(c, c->ws[grid::WS_IN_C].p, classes, nc * sizeof(int), s));
        }
        d_c = c->ws[grid::WS_IN_C].as<int>();
    }
    SSDR_CHECK_CUDA(cudaGetLastError());
    return grid::run_dev(c, s, c->ws[grid::WS_IN_P].as<float>(), d_f, d_c, N, fdim, ldim, sampleDl, order, M_out, handle);
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

int ssdr_grid_free(void* handle) {
    grid::free_handle((grid::Handle*)handle);
    return SSDR_OK;
}
}
