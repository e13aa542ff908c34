/*
 * ssdr_b200.h -- C ABI of libssdr_b200.so: the B200-native (sm_100a) drop-in for SSDR-AL's data-parallel
 * point-cloud hot path.  Plain pointers and sizes only; no torch / numpy / C++ types cross this boundary.
 *
 * Reference interfaces replaced (paths relative to SSDR_AL_s3dis/ of shaofeifei11/SSDR-AL):
 *   ssdr_knn, ssdr_knn_batch         <- cpp_knn / cpp_knn_omp / cpp_knn_batch / cpp_knn_batch_omp
 *                                       utils/nearest_neighbors/knn_.h:2-17 (bound by knn.pyx:8-23)
 *   ssdr_grid_subsample/_fetch/_free <- grid_subsampling()
 *                                       utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.h:84-91
 *                                       (bound by wrapper.cpp:58-286)
 *   ssdr_fps_f32 / ssdr_fps_f64      <- farthest_features_sample()        fps_gcn_cpu.py:119-147
 *   ssdr_kcenter_f32 / _f64          <- kCenterGreedy.select_batch_()     kcenterGreedy.py:84-128
 *
 * Conventions
 *   - Every function returns 0 (SSDR_OK) or a non-zero ssdr_status; ssdr_last_error() then returns a
 *     thread-local, NUL-terminated message.  Nothing throws, nothing aborts.
 *   - "host" entry points borrow caller-owned host buffers (row-major, C-contiguous) for the duration of the
 *     call and write caller-allocated outputs -- the same ownership rule as the reference prototypes.
 *   - "_dev" entry points take device pointers on the current device plus a cudaStream_t passed as void*
 *     (NULL = the CUDA default stream, exactly as in a kernel launch), so the work is ordered with the caller's own
 *     work on that stream.  Use one stream per calling thread: the per-thread workspaces are reused across calls.
 *   - The library is re-entrant per calling thread (per-thread stream + workspace).  It refuses to run in a
 *     fork()ed child of a process that already initialised it (SSDR_ERR_FORK) instead of hanging.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with SSDR_ERR_CUDA.
 */
#ifndef SSDR_B200_H
#define SSDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SSDR_OK = 0,
    SSDR_ERR_INVALID = 1,     /* bad argument (NULL pointer, dim != 3, ...) */
    SSDR_ERR_CUDA = 2,        /* CUDA runtime error or no device */
    SSDR_ERR_NOMEM = 3,       /* device or host allocation failed */
    SSDR_ERR_UNSUPPORTED = 4, /* valid in the reference but outside this implementation's documented limits */
    SSDR_ERR_FORK = 5,        /* called from a fork()ed child after CUDA initialisation in the parent */
    SSDR_ERR_EMPTY = 6        /* grid subsampling produced no point (reference: RuntimeError("Error")) */
} ssdr_status;

/* element types of arguments that exist in float32 and float64 flavours */
#define SSDR_F32 0
#define SSDR_F64 1

/* ---- runtime ------------------------------------------------------------------------------------ */
const char* ssdr_last_error(void);
int ssdr_version(void);               /* major*10000 + minor*100 + patch */
int ssdr_device_count(int* count);    /* number of visible CUDA devices */
int ssdr_set_device(int device);      /* device used by the calling thread's subsequent calls */
int ssdr_get_device(int* device);
int ssdr_device_sm_count(int* sms);
int ssdr_synchronize(void);           /* wait for the calling thread's library stream */
/* Pinned host memory helpers (optional; host entry points accept any host pointer but copy fastest from these). */
int ssdr_host_alloc(void** ptr, size_t bytes);
int ssdr_host_free(void* ptr);

/* ---- KNN ----------------------------------------------------------------------------------------- */
/* Exact K nearest neighbours with nanoflann-identical results (indices ascending by fp32 squared distance,
 * equal distances in nanoflann's tree-visit order).  dim must be 3.  If K > npts only the first npts slots
 * of each output row are written (the reference leaves the rest at the caller's zero initialisation). */
int ssdr_knn(const float* points, size_t npts, size_t dim, const float* queries, size_t nqueries, size_t K,
             int64_t* indices /* (nqueries, K) */);
int ssdr_knn_batch(const float* batch_data /* (B, npts, dim) */, size_t batch_size, size_t npts, size_t dim,
                   const float* queries /* (B, nqueries, dim) */, size_t nqueries, size_t K,
                   int64_t* batch_indices /* (B, nqueries, K) */);

typedef struct {
    uint64_t queries;         /* rows processed */
    uint64_t tie_rows;        /* rows whose top-(K+1) held an exact or near fp32 distance tie (re-resolved) */
    uint64_t tree_builds;     /* nanoflann-identical trees built on device for those rows */
    uint64_t dist_evals;      /* candidate distance evaluations of the main kernel */
    double grid_build_ms;     /* device time of stage A (bbox, cell counting sort of points and queries) */
    double main_kernel_ms;    /* device time of the main query kernel alone (CUDA events on the launch stream) */
    double tie_path_ms;       /* device time of the tie path (tree build + exact replay), 0 when unused */
    double tree_build_ms;     /* part of tie_path_ms spent before the exact replay starts (tree build) */
    uint64_t kernel_launches; /* kernels of this library the call launched (memsets and copies not counted) */
} ssdr_knn_stats;

/* knn_batch on batch items that are slices of larger arrays: item b of the points starts at
 * batch_data + b * data_item_stride floats (>= npts * dim; 0 = dense), likewise for the queries.  This is what
 * RandLA-Net's pyramid passes (s3dis_dataset.py:166 `batch_xyz[:, :N // ratio, :]`); the reference first packs such a
 * view on the host (np.ascontiguousarray, knn.pyx:95-96), here the copy engine packs it during the upload. */
int ssdr_knn_batch_strided(const float* batch_data, size_t batch_size, size_t npts, size_t dim, size_t data_item_stride,
                           const float* queries, size_t nqueries, size_t query_item_stride, size_t K,
                           int64_t* batch_indices);

/* Device-resident variant: d_points (B,npts,3), d_queries (B,nqueries,3), d_indices (B,nqueries,K) int64.
 * Enqueues on `stream`; stats (nullable, host struct) forces a stream synchronisation when requested. */
int ssdr_knn_batch_dev(const float* d_points, size_t batch_size, size_t npts, const float* d_queries,
                       size_t nqueries, size_t K, int64_t* d_indices, void* stream, ssdr_knn_stats* stats);
/* int32 output flavour for callers that feed TF int32 tensors (helper_tool.py:183 casts anyway). */
int ssdr_knn_batch_dev_i32(const float* d_points, size_t batch_size, size_t npts, const float* d_queries,
                           size_t nqueries, size_t K, int32_t* d_indices, void* stream, ssdr_knn_stats* stats);

/* RandLA-Net's whole input pyramid in ONE call -- the loop of s3dis_dataset.py:164-177 / helper_tool.py:173-183, which
 * the reference runs as 2 * n_levels knn_batch calls:  N_0 = npts, N_{l+1} = N_l / ratios[l]; level l+1 holds the first
 * N_{l+1} points of every item of level l (`batch_xyz[:, :N // ratio, :]`);
 *     d_neigh[l] (B, N_l, K) = knn_batch(level l, level l, K)      d_up[l] (B, N_l, 1) = knn_batch(level l+1, level l, 1)
 * d_points (B, npts, 3) and the output arrays are device pointers; `ratios` and the two pointer tables are host arrays.
 * Every level is enqueued on `stream` behind the previous one WITHOUT a host round trip (the exact tie path runs
 * speculatively on the device; the nanoflann-identical trees of a level are built at most once for its two queries).
 * The call returns as soon as the work is enqueued; ssdr_knn_status(stream) waits for the stream and reports (and
 * clears) an error of the asynchronous tie path. */
int ssdr_knn_pyramid_dev(const float* d_points, size_t batch_size, size_t npts, const int32_t* ratios, size_t n_levels,
                         size_t K, int64_t* const* d_neigh, int64_t* const* d_up, void* stream);
int ssdr_knn_status(void* stream);
/* The same pyramid with HOST arrays in and out, synchronous: batch_xyz (B, npts, dim == 3) float32, neigh[l] (B, N_l, K)
 * and up[l] (B, N_l, 1) int64 host arrays (pinned or pageable).  One upload of the points (the levels are prefixes),
 * the support clouds run as concurrent branches on the device and every branch's rows are copied back on the branch's
 * own stream while the others compute.  Requires K <= N_l for every level. */
int ssdr_knn_pyramid(const float* batch_xyz, size_t batch_size, size_t npts, size_t dim, const int32_t* ratios,
                     size_t n_levels, size_t K, int64_t* const* neigh, int64_t* const* up);
unsigned long long ssdr_knn_pyramid_launches(void); /* kernels launched by the calling thread's last pyramid call */

/* Diagnostic only: the nanoflann-identical tree built on the device for one cloud (node arrays: 3*npts+64 entries). */
int ssdr_knn_debug_tree(const float* points, size_t npts, uint32_t* vind_out, uint32_t* n_nodes_out, uint32_t* left,
                        uint32_t* right, int32_t* child1, int32_t* child2, int32_t* divfeat, float* divlow,
                        float* divhigh);

/* Diagnostic only: %globaltimer marks (ns) of the tree build for B clouds of npts points; see knn.cu. */
int ssdr_knn_debug_build_timing(const float* points, size_t B, size_t npts, uint64_t* marks16);

/* ---- grid subsampling ------------------------------------------------------------------------------ */
#define SSDR_GRID_ORDER_KEY 0       /* rows in ascending voxel key (canonical, deterministic) */
#define SSDR_GRID_ORDER_REFERENCE 1 /* rows in the reference's libstdc++ hash-iteration order */

/* Phase 1: run on the device, learn M.  feats / classes nullable (then fdim / ldim ignored).
 * Phase 2: ssdr_grid_fetch copies the M rows into caller-allocated host arrays (nullable each).
 * keys_out / counts_out are extras (voxel key, points per voxel) for diagnostics and tests. */
int ssdr_grid_subsample(const float* points /* (N,3) */, const float* feats /* (N,fdim) */,
                        const int32_t* classes /* (N,ldim) */, size_t N, size_t fdim, size_t ldim, float sampleDl,
                        int order, size_t* M_out, void** handle);
/* Same, with the element type of feats / classes stated: every data-prep caller of the reference passes uint8
 * colours and uint8 labels (e.g. utils/data_prepare_s3dis.py:58), which wrapper.cpp:100-106 widens on the host to
 * float32 / int32.  SSDR_DTYPE_U8 uploads the bytes as they are and widens them on the device (exact). */
#define SSDR_DTYPE_NATIVE 0 /* float32 features, int32 classes */
#define SSDR_DTYPE_U8 1
int ssdr_grid_subsample_typed(const float* points, const void* feats, int feats_dtype, const void* classes,
                              int classes_dtype, size_t N, size_t fdim, size_t ldim, float sampleDl, int order,
                              size_t* M_out, void** handle);
int ssdr_grid_fetch(void* handle, float* points_out, float* feats_out, int32_t* classes_out);
int ssdr_grid_fetch_ex(void* handle, float* points_out, float* feats_out, int32_t* classes_out,
                       uint64_t* keys_out, int32_t* counts_out);
int ssdr_grid_free(void* handle);
/* Diagnostic only: %globaltimer marks (ns) of the calling thread's last subsampling call -- [0] start, [1] geometry,
 * [2] keys, [3..10] end of radix pass k, [11] heads counted, [12] voxel starts written, [13] / [14] reduce start / end. */
int ssdr_grid_debug_timing(uint64_t* marks16, int* key_bits);

/* Device-resident variant: inputs are device pointers; results stay on the device inside the handle and can be
 * read back with ssdr_grid_fetch or borrowed with ssdr_grid_dev_ptrs (valid until ssdr_grid_free). */
int ssdr_grid_subsample_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N,
                            size_t fdim, size_t ldim, float sampleDl, int order, void* stream, size_t* M_out,
                            void** handle);
int ssdr_grid_dev_ptrs(void* handle, const float** d_points, const float** d_feats, const int32_t** d_classes);

/* One large cloud over several GPUs ("voxel-layer slabs, halo-free ownership").  The grid geometry
 * (origin, nX, nY of grid_subsampling.cpp:33-43) comes from the bounding box of the WHOLE cloud, and each rank reduces
 * only the voxel layers [layer_lo, layer_hi) along `axis`: every voxel has exactly one owner, sees its points in input
 * order, and so carries the same bits as in a single-GPU run.  With axis == 2 the ranks' rows concatenated in rank
 * order are the single-GPU SSDR_GRID_ORDER_KEY result.
 *   bbox : {minx,miny,minz,maxx,maxy,maxz} of the whole cloud (nullable = min/max of d_points, i.e. this call sees the
 *          whole cloud and only selects its slab); axis == -1 disables the slab selection (all points take part).
 * An empty slab succeeds with *M_out == 0.  Helpers: ssdr_grid_bbox_dev reduces a chunk's min/max (to be combined
 * across ranks by the caller), ssdr_grid_point_layers_dev writes each point's layer index along `axis` (clamped to
 * INT32_MAX) so the caller can histogram layers, balance the slabs and route points to their owner. */
int ssdr_grid_subsample_slab_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N,
                                 size_t fdim, size_t ldim, float sampleDl, int order, const float* bbox, int axis,
                                 unsigned long long layer_lo, unsigned long long layer_hi, void* stream,
                                 size_t* M_out, void** handle);
/* The same into caller-owned device arrays of `capacity` >= N rows each (the voxel count is known only on the device, so
 * the worst case of one voxel per point must fit): no handle, no allocation, ONE stream synchronisation (the read-back
 * of *M_out).  Rows in ascending voxel key.  bbox nullable, axis == -1 = whole cloud (see ssdr_grid_subsample_slab_dev);
 * d_keys_out / d_counts_out nullable. */
int ssdr_grid_subsample_into_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N,
                                 size_t fdim, size_t ldim, float sampleDl, const float* bbox, int axis,
                                 unsigned long long layer_lo, unsigned long long layer_hi, float* d_points_out,
                                 float* d_feats_out, int32_t* d_classes_out, uint64_t* d_keys_out, int32_t* d_counts_out,
                                 size_t capacity, void* stream, size_t* M_out);
int ssdr_grid_bbox_dev(const float* d_points, size_t N, void* stream, float* bbox_out /* 6 floats, host */);
int ssdr_grid_point_layers_dev(const float* d_points, size_t N, const float* bbox /* nullable */, float sampleDl,
                               int axis, int32_t* d_layers, void* stream, unsigned long long* n_layers_out);

/* The exchange step of slab sharding when every rank starts with a row chunk of the cloud: ssdr_grid_layer_hist_dev
 * counts this chunk's points per voxel layer (to be summed over the ranks and cut into balanced slabs by the caller),
 * ssdr_grid_route_dev groups the chunk's rows by destination rank -- rank r owns the layers [bounds[r], bounds[r+1]) --
 * with a stable device-side partition (input order kept inside every destination), ready for one all-to-all; it
 * returns the number of rows per destination in counts_out (host).  bbox = corners of the WHOLE cloud. */
int ssdr_grid_layer_hist_dev(const float* d_points, size_t N, const float* bbox, float sampleDl, int axis,
                             unsigned long long* d_hist /* n_layers, device */, size_t n_layers,
                             size_t sample_stride /* count every k-th point: balancing needs no exact counts */,
                             void* stream);
int ssdr_grid_route_dev(const float* d_points, const float* d_feats, const int32_t* d_classes, size_t N, size_t fdim,
                        size_t ldim, float sampleDl, const float* bbox, int axis,
                        const unsigned long long* bounds /* world + 1, host */, int world, float* d_points_out,
                        float* d_feats_out, int32_t* d_classes_out, unsigned long long* counts_out /* world, host */,
                        void* stream);

/* ---- farthest-feature sampling / k-center greedy ---------------------------------------------------- */
/* FPS: out[0] = first; out[s+1] = argmax_i min_{t<=s} sum_j (F[i,j]-F[out[t],j])^2, first index on ties,
 * distances in the input dtype with numpy's pairwise summation order (bit-exact picks). */
int ssdr_fps_f32(const float* F, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* out);
int ssdr_fps_f64(const double* F, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* out);
int ssdr_fps_f32_dev(const float* d_F, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* d_out,
                     void* stream);
int ssdr_fps_f64_dev(const double* d_F, size_t N, size_t D, int32_t first, size_t n_samples, int32_t* d_out,
                     void* stream);

/* k-center greedy with sklearn's euclidean distance (sqrt(max(0,|x|^2+|c|^2-2x.c)) accumulated in float64,
 * rounded to the input dtype).  selected: n_sel indices already chosen; out: n_pick new indices. */
int ssdr_kcenter_f32(const float* X, size_t N, size_t D, const int64_t* selected, size_t n_sel, size_t n_pick,
                     int64_t* out);
int ssdr_kcenter_f64(const double* X, size_t N, size_t D, const int64_t* selected, size_t n_sel, size_t n_pick,
                     int64_t* out);
int ssdr_kcenter_f32_dev(const float* d_X, size_t N, size_t D, const int64_t* d_selected, size_t n_sel,
                         size_t n_pick, int64_t* d_out, void* stream);
int ssdr_kcenter_f64_dev(const double* d_X, size_t N, size_t D, const int64_t* d_selected, size_t n_sel,
                         size_t n_pick, int64_t* d_out, void* stream);

/* ---- chamfer adjacency of the superpoints of one room ------------------------------------------------- */
/* create_cd / chamfer_distance (fps_gcn_cpu.py:12-38), the step that feeds the FPS loop: S small clouds, already
 * centred by the caller, concatenated in `points` (T,3) float64 with `offsets` (S+1, offsets[0] = 0);
 * out (S,S) float64: out[c][i] = mean_{p in i} min_{q in c} |p-q| + mean_{q in c} min_{p in i} |q-p|, 0 on the
 * diagonal.  Distances as sklearn's KDTree computes them (float64, sqrt(((dx^2+dy^2)+dz^2))), means in numpy's
 * pairwise order.  The _dev variant takes device pointers plus a host copy of the offsets (sizes the scratch). */
int ssdr_chamfer_matrix_f64(const double* points, const int64_t* offsets, size_t S, double* out);
/* farthest_superpoint_sample (sampler2.py:49-80): FPS over the S clouds with distance = squared distance of the
 * centroids (S,3 float64) + chamfer distance to the current pick (the chamfer row of each pick is computed on demand,
 * like the reference's loop); out[0] = trigger_idx, first arg-max on ties, strict '<' minimum update from 1e10. */
int ssdr_superpoint_fps_f64(const double* points, const int64_t* offsets, size_t S, const double* centroids,
                            int32_t trigger_idx, size_t n_samples, int32_t* out);
/* Same with the centroids in their own dtype (SSDR_F32 / SSDR_F64, defined below): the reference evaluates
 * np.sum((centroids - current) ** 2, axis=-1) in the dtype of the centroid array -- float32 for ply coordinates -- and
 * only widens when the float64 chamfer row is added (sampler2.py:68-73). */
int ssdr_superpoint_fps(const double* points, const int64_t* offsets, size_t S, const void* centroids, int centroid_dtype,
                        int32_t trigger_idx, size_t n_samples, int32_t* out);
int ssdr_chamfer_matrix_f64_dev(const double* d_points, const int64_t* d_offsets, const int64_t* h_offsets, size_t S,
                                double* d_out, void* stream);

/* ---- superpoint adjacency and feature propagation in front of the FPS loop ------------------------------ */
/* fps_adj_all's matrix arithmetic (fps_gcn_cpu.py:60-117) on the device.  Room b holds the superpoints
 * block_off[b] .. block_off[b+1]: ref[] = their rows in [0, N), centres (., 3) float64, and its chamfer block cd
 * (n_b x n_b float64, the blocks concatenated in room order).  S = A_ed + A_cd (1e10 + 1e10 outside the blocks),
 * adj = exp(-S) - I, column-scaled by 1 / rowsum (inf -> 0), + I.  The matrix stays on the device behind *handle;
 * ssdr_gcn_fetch copies it into a (N, N) float64 host array, ssdr_gcn_free releases it. */
int ssdr_gcn_adjacency_f64(size_t N, size_t n_blocks, const int64_t* block_off, const int64_t* ref, const double* centres,
                           const double* cd, void** handle);
int ssdr_gcn_fetch(void* handle, double* adj_out);
int ssdr_gcn_free(void* handle);
/* GCN_FPS_sampling's propagation (fps_gcn_cpu.py:153-167): optionally keep the gcn_top largest entries of every row of
 * adj (ties at the threshold: highest columns), then out = V + adj V + adj (adj V) + ... (gcn_number products).
 * The adjacency comes from `handle` (device resident) or, when handle is NULL, from the host array adj (N, N).
 * V and out are (N, D) float64 host arrays. */
int ssdr_gcn_propagate_f64(void* handle, const double* adj, size_t N, const double* V, size_t D, int gcn_number,
                           int gcn_top, double* out);

/* Row-sharded multi-GPU selection (one process per GPU).  Every rank holds the FULL matrix d_F (so the chosen
 * centre row is local) but scans only rows [row_begin,row_end); after each step the packed (distance, index)
 * candidates are combined across ranks by an 8-byte max all-reduce over NCCL.  `nccl_comm` is an ncclComm_t.
 * All ranks receive identical picks. */
int ssdr_fps_f32_sharded(const float* d_F, size_t N, size_t D, size_t row_begin, size_t row_end, int32_t first,
                         size_t n_samples, int32_t* d_out, void* nccl_comm, void* stream);
/* The same row-sharded selection with the per-pick exchange FUSED into the persistent kernel: CTA 0 of every rank
 * stores the rank's (distance, index) winner straight into every peer's mailbox over NVLink (peer memory mapped with
 * CUDA IPC), every CTA of every rank polls its local copy -- one launch for all picks, no collective call per pick.
 * Peer group life cycle (one process per GPU; `handles` = the world 64-byte handles in rank order, exchanged by the
 * caller, e.g. with torch.distributed.all_gather):
 *     ssdr_peer_group_create(world, rank, &g); ssdr_peer_group_export(g, h64); <all-gather h64>;
 *     ssdr_peer_group_connect(g, handles); <barrier>; ... sharded calls ...; <barrier>; ssdr_peer_group_destroy(g);
 * All ranks must issue the same sequence of sharded calls on a group.  FPS (farthest_features_sample) and k-center
 * (kCenterGreedy) in float32 or float64; every rank receives identical picks, bit-equal to the single-GPU entry points.
 * A rank whose peers do not answer within SSDR_PEER_TIMEOUT_MS (default 20000) fails with SSDR_ERR_CUDA instead of
 * hanging.  max_ctas = 0 uses every SM (tests run several virtual ranks on one device with a smaller grid each,
 * connected with ssdr_peer_group_connect_local).  These calls synchronise `stream` before returning. */
int ssdr_peer_group_create(int world, int rank, void** group);
int ssdr_peer_group_export(void* group, void* handle64 /* 64 bytes: cudaIpcMemHandle_t */);
int ssdr_peer_group_connect(void* group, const void* handles /* world x 64 bytes */);
int ssdr_peer_group_connect_local(void** groups, int world);
int ssdr_peer_group_destroy(void* group);
/* With SSDR_PEER_DEFER_CHECK=1 in the environment the two sharded entry points below return once their kernel is
 * enqueued; ssdr_peer_group_check then synchronises `stream` and reports a peer time-out. */
int ssdr_peer_group_check(void* group, void* stream);
int ssdr_fps_sharded_p2p(int dtype, const void* d_F, size_t N, size_t D, size_t row_begin, size_t row_end, int32_t first,
                         size_t n_samples, int32_t* d_out, void* group, void* stream, int max_ctas);
int ssdr_kcenter_sharded_p2p(int dtype, const void* d_X, size_t N, size_t D, size_t row_begin, size_t row_end,
                             const int64_t* d_selected, size_t n_sel, size_t n_pick, int64_t* d_out, void* group,
                             void* stream, int max_ctas);
/* NCCL bootstrap helpers so the host side can create the communicator without linking NCCL itself. */
int ssdr_nccl_unique_id(void* id128 /* 128 bytes */);
int ssdr_nccl_comm_init(void** comm, int nranks, const void* id128, int rank);
int ssdr_nccl_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* SSDR_B200_H */
