/*
 * b2points.h -- C ABI of the B200-native scan-matching linearization library (libb2points.so).
 *
 * This is the drop-in boundary for the per-point correspondence + linearization hot path of
 * koide3/gtsam_points.  Plain C: opaque handles, raw pointers, sizes.  No C++
 * types, no torch types, no exceptions; every function returns a b2_status and leaves a message
 * retrievable with b2_last_error() on failure.  The library never frees or retains caller memory:
 * host inputs are copied (to the device) before a call returns.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference root).
 * The GTSAM-typed adapter classes that keep the reference's C++ surface on top of this ABI live in
 * gtsam_points_b200/cpp/ ; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - poses / deltas: 4x4 row-major double, delta = T_target^-1 * T_source
 *     (src/gtsam_points/factors/integrated_matching_cost_factor.cpp:57-69).
 *   - points: `point_stride` doubles per point (3 = packed xyz, 4 = the reference's Vector4d (x,y,z,1),
 *     include/gtsam_points/types/point_cloud.hpp:106).
 *   - covariances: `cov_stride` doubles per point (9 = 3x3, 16 = the reference's Matrix4d with zero
 *     row/col 3, point_cloud.hpp:108); symmetric, so row- and column-major coincide.
 *   - tangent ordering of H, b: GTSAM Pose3 [rotation(3), translation(3)].
 *   - all 6x6 blocks are row-major; H_target_source = sum J_target^T M J_source.
 *   - b_* are the raw sums J^T M r; the HessianFactor takes -b
 *     (integrated_matching_cost_factor.cpp:46-52).
 */
#ifndef B2POINTS_H_
#define B2POINTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B2_API
#else
#define B2_API __attribute__((visibility("default")))
#endif

typedef enum b2_status {
  B2_OK = 0,
  B2_ERR_INVALID_ARGUMENT = 1,
  B2_ERR_CUDA = 2,
  B2_ERR_OUT_OF_MEMORY = 3,
  B2_ERR_INVALID_STATE = 4,
  B2_ERR_NO_DEVICE = 5
} b2_status;

typedef struct b2_ctx b2_ctx;               /* device + stream + staging buffers                                  */
typedef struct b2_cloud b2_cloud;           /* device copy of a PointCloud (points + covs)                        */
typedef struct b2_voxelmap b2_voxelmap;     /* GaussianVoxelMapGPU replacement                                    */
typedef struct b2_kdtree b2_kdtree;         /* NearestNeighborSearch / KdTree replacement on the device           */
typedef struct b2_factor b2_factor;         /* Integrated{VGICP,GICP}Factor device state                          */
typedef struct b2_factor_set b2_factor_set; /* NonlinearFactorSetGPU replacement: one batched launch over factors */

/* One linearized factor.  Replaces `LinearizedSystem6` (include/gtsam_points/cuda/kernels/linearized_system.cuh:10-71)
 * and the five out-parameters of IntegratedMatchingCostFactor::evaluate
 * (include/gtsam_points/factors/integrated_matching_cost_factor.hpp:74-83).  128 doubles = 1 KiB. */
typedef struct b2_linearized {
  double H_target[36];
  double H_source[36];
  double H_target_source[36];
  double b_target[6];
  double b_source[6];
  double error;       /* sum r^T M r (not halved)                      */
  double num_inliers; /* number of source points with a correspondence */
  double reserved[6];
} b2_linearized;

#define B2_LINEARIZED_DOUBLES 128

/* b2_cloud_create flags */
#define B2_CLOUD_DEFAULT 0u
#define B2_CLOUD_NO_REORDER 1u   /* keep the caller's point order on the device (default: Morton order)                       */
#define B2_CLOUD_COMPACT_F32 2u  /* store points AND covariances as float32 even if that rounds them: the reference's        \
                                    PointCloudGPU layout (point_cloud.hpp:115-117).  Default is lossless: float32 storage is \
                                    chosen per array only when every value is exactly float32-representable, else float64.   */
#define B2_CLOUD_FORCE_F64 4u    /* always store float64                                                                     */

typedef struct b2_cloud_info {
  uint64_t num_points;
  int32_t point_bytes;     /* 4 or 8: storage type of coordinates */
  int32_t cov_bytes;       /* 4 or 8: storage type of covariances, 0 = no covariances */
  int32_t reordered;       /* 1 if stored in Morton order */
  int32_t reserved;
  uint64_t device_bytes;
} b2_cloud_info;

typedef struct b2_voxelmap_info {
  uint64_t num_voxels;
  uint64_t num_buckets;
  double resolution;
  uint64_t device_bytes;
} b2_voxelmap_info;

/* ------------------------------------------------------------------------------------------------------------
 * Library / context
 * ---------------------------------------------------------------------------------------------------------- */

/* Message of the last failing call on this thread ("" if none). */
B2_API const char* b2_last_error(void);
B2_API const char* b2_version(void);

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t the caller owns (e.g. torch's current
 * stream) or NULL to let the context create its own non-blocking stream.  Replaces the reference's per-object
 * CUstream_st* arguments and StreamTempBufferRoundRobin (include/gtsam_points/cuda/stream_temp_buffer_roundrobin.hpp:49-65). */
B2_API b2_status b2_ctx_create(int device, void* stream, b2_ctx** out);
B2_API b2_status b2_ctx_destroy(b2_ctx* ctx);
B2_API b2_status b2_ctx_synchronize(b2_ctx* ctx);
B2_API void* b2_ctx_stream(b2_ctx* ctx);
/* Plain device memory on the context's device, for callers above the ABI that own result / exchange buffers without a CUDA
 * toolchain of their own (the reference keeps such blobs in thrust::device_vector, cuda/nonlinear_factor_set_gpu.hpp:110-118).
 * b2_memcpy_*: stream-ordered on the context's stream, then synchronised. */
B2_API b2_status b2_device_malloc(b2_ctx* ctx, size_t bytes, void** out);
B2_API b2_status b2_device_free(b2_ctx* ctx, void* ptr);
B2_API b2_status b2_memcpy_d2h(b2_ctx* ctx, void* dst_host, const void* src_device, size_t bytes);
B2_API b2_status b2_memcpy_h2d(b2_ctx* ctx, void* dst_device, const void* src_host, size_t bytes);

/* ------------------------------------------------------------------------------------------------------------
 * Point clouds.  Replaces PointCloudGPU's upload of points/covs (src/gtsam_points/types/point_cloud_gpu.cu)
 * as consumed through frame::traits (include/gtsam_points/types/frame_traits.hpp:27-168).
 * ---------------------------------------------------------------------------------------------------------- */
B2_API b2_status b2_cloud_create(b2_ctx* ctx, const double* points, int point_stride, const double* covs, int cov_stride, size_t n,
                                 unsigned flags, b2_cloud** out);
B2_API b2_status b2_cloud_destroy(b2_cloud* cloud);
B2_API b2_status b2_cloud_get_info(const b2_cloud* cloud, b2_cloud_info* info);

/* ------------------------------------------------------------------------------------------------------------
 * Gaussian voxel map.  Replaces GaussianVoxelMapGPU (include/gtsam_points/types/gaussian_voxelmap_gpu.hpp:39-108)
 * with the index semantics of GaussianVoxelMapCPU (src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:59-77):
 * voxel ids are first-touch order, no point is ever dropped, lookups are exact.
 * ---------------------------------------------------------------------------------------------------------- */

/* One-shot build on the device from a host cloud: GaussianVoxelMap::insert(const PointCloud&)
 * (include/gtsam_points/types/gaussian_voxelmap.hpp:27; CPU semantics ann/impl/incremental_voxelmap_impl.hpp:31-68). */
B2_API b2_status b2_voxelmap_create_from_points(b2_ctx* ctx, double resolution, const double* points, int point_stride, const double* covs,
                                                int cov_stride, size_t n, b2_voxelmap** out);
/* Incremental form: an empty map, then any number of insert() calls -- IncrementalVoxelMap<GaussianVoxel>::insert with its LRU
 * eviction (include/gtsam_points/ann/impl/incremental_voxelmap_impl.hpp:31-68; defaults lru_horizon = lru_clear_cycle = 10):
 * voxels that exist are continued (re-opened, summed on, re-finalized), new voxels get the next ids in first-touch order,
 * every lru_clear_cycle-th insert removes the voxels untouched for more than lru_horizon inserts and re-indexes the rest.
 * Ids, counts, means and covariances equal the CPU map's after every call.  (The reference's GPU map is one-shot only,
 * types/gaussian_voxelmap_gpu.hpp:63.)  Factors / factor sets built over the map see the new contents at their next call. */
B2_API b2_status b2_voxelmap_create(b2_ctx* ctx, double resolution, b2_voxelmap** out);
B2_API b2_status b2_voxelmap_set_lru(b2_voxelmap* vm, size_t lru_horizon, size_t lru_clear_cycle);
B2_API b2_status b2_voxelmap_insert(b2_voxelmap* vm, const double* points, int point_stride, const double* covs, int cov_stride, size_t n);
/* Upload an existing map (e.g. a GaussianVoxelMapCPU's flat_voxels or a save_compact file,
 * include/gtsam_points/types/gaussian_voxel_data.hpp:11-54): coords V x 3, means V x 3, covs V x 9, num_points V. */
B2_API b2_status b2_voxelmap_create_from_voxels(b2_ctx* ctx, double resolution, const int32_t* coords, const double* means, const double* covs,
                                                const int32_t* num_points, size_t num_voxels, b2_voxelmap** out);
B2_API b2_status b2_voxelmap_destroy(b2_voxelmap* vm);
B2_API b2_status b2_voxelmap_get_info(const b2_voxelmap* vm, b2_voxelmap_info* info);
/* download_voxel_means / _covs / _num_points / buckets (gaussian_voxelmap_gpu.hpp:110-114); any pointer may be NULL */
B2_API b2_status b2_voxelmap_download(const b2_voxelmap* vm, int32_t* coords, double* means, double* covs, int32_t* num_points);
/* GaussianVoxelMapCPU::save_compact / ::load (src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:79-135): the reference's wire
 * format -- 7 text header lines, then num_voxels packed 56-byte GaussianVoxelData records (types/gaussian_voxel_data.hpp:11-54:
 * int32 coord[3], int32 num_points, float mean[3], float cov[6] = (00,01,02,11,12,22), float intensity).  Files written by either
 * side load on the other. */
B2_API b2_status b2_voxelmap_save_compact(const b2_voxelmap* vm, const char* path);
B2_API b2_status b2_voxelmap_load(b2_ctx* ctx, const char* path, b2_voxelmap** out);
/* overlap_gpu (include/gtsam_points/types/gaussian_voxelmap_gpu.hpp:116-125, src/.../gaussian_voxelmap_gpu_funcs.cu:65-194):
 * fraction of the source points p for which T_j p lies in a voxel of target j for some j (first match counts).
 * Ts_target_source: num_targets x 16 doubles (row-major 4x4). */
B2_API b2_status b2_overlap(const b2_voxelmap* const* targets, size_t num_targets, const b2_cloud* source, const double* Ts_target_source, double* out_overlap);
/* merge_frames_gpu (src/gtsam_points/types/gaussian_voxelmap_gpu_funcs.cu:196-330; CPU twin gaussian_voxelmap_cpu_funcs.cpp:25-113):
 * merges num_frames posed frames (poses: num_frames x 16, row-major 4x4, world <- frame) into ONE cloud downsampled on a grid of
 * downsample_resolution laid out in the FIRST frame's coordinates: per occupied voxel the mean of the world points and of the
 * rotated covariances, voxels in ascending 63-bit key order (the CPU function's order and summation order: bit-identical).
 * out_points (m x 3) / out_covs (m x 9, row-major 3x3; may be NULL if no frame has covariances) must hold one entry per input
 * point (m <= total); *out_n = m.  Feed them to b2_cloud_create for a device-resident merged frame. */
B2_API b2_status b2_merge_frames(b2_ctx* ctx, const double* poses, const b2_cloud* const* frames, size_t num_frames, double downsample_resolution,
                                 double* out_points, double* out_covs, size_t* out_n);
/* voxel_coord + lookup_voxel_index for n host points (gaussian_voxelmap_cpu.cpp:59-69); out_index[i] = id or -1 */
B2_API b2_status b2_voxelmap_lookup(const b2_voxelmap* vm, const double* points, int point_stride, size_t n, int32_t* out_index);

/* ------------------------------------------------------------------------------------------------------------
 * Nearest-neighbour search.  Replaces NearestNeighborSearch::knn_search for k = 1
 * (include/gtsam_points/ann/nearest_neighbor_search.hpp:31-35, ann/kdtree2.hpp:52-61): exact.
 * ---------------------------------------------------------------------------------------------------------- */
B2_API b2_status b2_kdtree_create(b2_ctx* ctx, const double* points, int point_stride, size_t n, b2_kdtree** out);
B2_API b2_status b2_kdtree_destroy(b2_kdtree* tree);
/* queries: host, nq x query_stride doubles.  out_index[i] = index of the nearest target point with squared distance
 * < max_sq_dist, else -1; out_sq_dist[i] = its squared distance (max_sq_dist if none).  Either output may be NULL. */
B2_API b2_status b2_kdtree_knn1(const b2_kdtree* tree, const double* queries, int query_stride, size_t nq, double max_sq_dist, int64_t* out_index,
                                double* out_sq_dist);

/* General form: the k nearest neighbours, 1 <= k <= B2_KNN_MAX_K, of every query -- NearestNeighborSearch::knn_search with
 * KnnResult's output convention (include/gtsam_points/ann/knn_result.hpp:44-109): out_index / out_sq_dist are nq x k, sorted by
 * distance, entries beyond the number found hold (-1, max_sq_dist); exact (KnnSetting::epsilon = 0). */
#define B2_KNN_MAX_K 64
B2_API b2_status b2_kdtree_knn(const b2_kdtree* tree, const double* queries, int query_stride, size_t nq, int k, double max_sq_dist, int64_t* out_index,
                               double* out_sq_dist);
/* estimate_covariances (include/gtsam_points/features/covariance_estimation.hpp:14-66, src/.../covariance_estimation.cpp:18-77):
 * per point the covariance of its k nearest neighbours (itself included), EIG-regularised: eigenvalues replaced by
 * eigen_values[3] in ascending-eigenvalue order (NULL = the reference's default (1e-3, 1, 1)).  out_cov3x3: host, n x 9
 * (row-major 3x3; embed into Matrix4d with zero row / column 3).  Points with fewer than k neighbours get identity.
 * The first form builds its own kd-tree like the reference does; the second re-uses an existing tree over the same points. */
B2_API b2_status b2_estimate_covariances(b2_ctx* ctx, const double* points, int point_stride, size_t n, int k_neighbors, const double* eigen_values, double* out_cov3x3);
B2_API b2_status b2_kdtree_estimate_covariances(const b2_kdtree* tree, int k_neighbors, const double* eigen_values, double* out_cov3x3);

/* ------------------------------------------------------------------------------------------------------------
 * Factors.  Replace IntegratedVGICPFactor_ / IntegratedVGICPFactorGPU and IntegratedGICPFactor_
 * (include/gtsam_points/factors/integrated_vgicp_factor.hpp:37-54, integrated_vgicp_factor_gpu.hpp:43-66,
 *  integrated_gicp_factor.hpp:44-78).  A factor retains its target / source handles (they must outlive it).
 * ---------------------------------------------------------------------------------------------------------- */
B2_API b2_status b2_vgicp_factor_create(b2_ctx* ctx, const b2_voxelmap* target, const b2_cloud* source, b2_factor** out);
/* target_cloud must carry covariances; tree must have been built over the same target points. */
B2_API b2_status b2_gicp_factor_create(b2_ctx* ctx, const b2_cloud* target_cloud, const b2_kdtree* tree, const b2_cloud* source, b2_factor** out);
/* IntegratedICPFactor_ / IntegratedPointToPlaneICPFactor_ (include/gtsam_points/factors/integrated_icp_factor.hpp:27-145,
 * impl/integrated_icp_factor_impl.hpp:131-248): the same kd-tree search, residual r = mu_B - T p weighted with the identity
 * (point-to-point) or, when use_point_to_plane != 0, scaled row-wise by the target point's normal (target_normals: host, n x 3,
 * caller order).  Neither cloud needs covariances. */
B2_API b2_status b2_icp_factor_create(b2_ctx* ctx, const b2_cloud* target_cloud, const b2_kdtree* tree, const b2_cloud* source, int use_point_to_plane,
                                      const double* target_normals, b2_factor** out);
B2_API b2_status b2_factor_destroy(b2_factor* f);
/* IntegratedGICPFactor_::set_max_correspondence_distance (integrated_gicp_factor.hpp:98-101); default 1.0 */
B2_API b2_status b2_factor_set_max_correspondence_distance(b2_factor* f, double dist);
/* IntegratedGICPFactor_::set_correspondence_update_tolerance (integrated_gicp_factor.hpp:103-109, impl:135-147): a linearize()
 * whose pose lies within (angle [rad], trans [m]) of the pose of the last correspondence update KEEPS those correspondences and
 * linearizes them at the new pose; otherwise it re-associates.  Default 0 / 0 = always re-associate.  Honoured by the entry
 * points that receive HOST poses (b2_factor_linearize, b2_factor_set_linearize, *_issue_linearize); the device-pose variants
 * cannot see the pose and always re-associate.  Set it before the factor joins a user-built set.  kd-tree factors only. */
B2_API b2_status b2_factor_set_correspondence_update_tolerance(b2_factor* f, double angle, double trans);
B2_API size_t b2_factor_num_points(const b2_factor* f);
/* Correspondences frozen at the last linearize, in the caller's point order: VGICP voxel id / GICP target index, -1 = none
 * (IntegratedVGICPFactor_::correspondences, integrated_vgicp_factor.hpp:107; IntegratedGICPFactor_::correspondences :147). */
B2_API b2_status b2_factor_correspondences(const b2_factor* f, int64_t* out);

/* linearize(): update_correspondences(delta) + evaluate(delta, H..., b...) in one fused pass
 * (integrated_matching_cost_factor.cpp:37-55).  Single-factor convenience wrappers around a factor set of size 1. */
B2_API b2_status b2_factor_linearize(b2_factor* f, const double* delta, b2_linearized* out);
/* error(): evaluate(delta_eval) re-using the correspondences and fused covariances frozen at the last linearize
 * (integrated_matching_cost_factor.cpp:32-35, integrated_vgicp_factor_impl.hpp:183-224).  If the factor has never been
 * linearized, correspondences are first established at delta_eval (integrated_vgicp_factor_impl.hpp:183-185). */
B2_API b2_status b2_factor_error(b2_factor* f, const double* delta_eval, double* out_error);

/* Per-factor asynchronous protocol: the device half of NonlinearFactorGPU::issue_linearize / issue_compute_error / sync
 * (include/gtsam_points/factors/nonlinear_factor_gpu.hpp:85-121).  delta is a HOST pose (captured before the call returns),
 * d_out a DEVICE pointer (1 KiB record / one double); the kernel is enqueued on the factor's context stream. */
B2_API b2_status b2_factor_issue_linearize(b2_factor* f, const double* delta, double* d_out);
B2_API b2_status b2_factor_issue_error(b2_factor* f, const double* delta_eval, double* d_out_error);
B2_API b2_status b2_factor_sync(b2_factor* f);

/* ------------------------------------------------------------------------------------------------------------
 * Factor sets.  Replace NonlinearFactorSetGPU::linearize / ::error
 * (src/gtsam_points/cuda/nonlinear_factor_set_gpu.cpp:64-139, :141-218): all factors of the set are evaluated by
 * ONE kernel launch per (factor kind, storage type) group, one H2D of the poses and one D2H of the results.
 * ---------------------------------------------------------------------------------------------------------- */
B2_API b2_status b2_factor_set_create(b2_ctx* ctx, b2_factor* const* factors, size_t num_factors, b2_factor_set** out);
B2_API b2_status b2_factor_set_destroy(b2_factor_set* set);
B2_API size_t b2_factor_set_size(const b2_factor_set* set);
/* deltas: host, F x 16 doubles; out: host, F records.  Blocks until the results are on the host. */
B2_API b2_status b2_factor_set_linearize(b2_factor_set* set, const double* deltas, b2_linearized* out);
/* deltas_eval: host, F x 16; out_errors: host, F doubles. */
B2_API b2_status b2_factor_set_error(b2_factor_set* set, const double* deltas_eval, double* out_errors);
/* Device-resident variants for callers that keep poses / results on the GPU (multi-GPU all-reduce of the result
 * buffer, kernel-only timing): d_deltas F x 16 doubles, d_out F x 128 doubles, both device pointers valid on the
 * context's stream.  Asynchronous: no host synchronisation; ordering is the stream's. */
B2_API b2_status b2_factor_set_linearize_device(b2_factor_set* set, const double* d_deltas, double* d_out);
B2_API b2_status b2_factor_set_error_device(b2_factor_set* set, const double* d_deltas_eval, double* d_out_errors);
/* Asynchronous split of the two calls above -- the issue / sync / store protocol of NonlinearFactorGPU
 * (include/gtsam_points/factors/nonlinear_factor_gpu.hpp:49-121: set_linearization_point + issue_linearize, sync,
 * store_linearized; set_evaluation_point + issue_compute_error, store_computed_error) at factor-set granularity, which is how
 * NonlinearFactorSetGPU drives it (src/gtsam_points/cuda/nonlinear_factor_set_gpu.cpp:91-133).
 *   issue_*: deltas are HOST poses (F x 16), captured before the call returns (a set of one factor passes its pose BY VALUE as
 *            a kernel parameter: no copy operation at all); the kernels are enqueued on the context's stream and the results
 *            stay on the device -- in d_out / d_out_errors (device pointers) or, if NULL, in the set's own buffer.
 *   sync:    waits for everything issued on the set's stream.
 *   store_*: copies the results of the last issue_* out of the set's own buffer to the host (synchronises itself). */
B2_API b2_status b2_factor_set_issue_linearize(b2_factor_set* set, const double* deltas, double* d_out);
B2_API b2_status b2_factor_set_issue_error(b2_factor_set* set, const double* deltas_eval, double* d_out_errors);
B2_API b2_status b2_factor_set_sync(b2_factor_set* set);
B2_API b2_status b2_factor_set_store_linearized(b2_factor_set* set, b2_linearized* out);
B2_API b2_status b2_factor_set_store_errors(b2_factor_set* set, double* out_errors);
/* Multi-GPU exchange fused into the kernel (one process per GPU, all GPUs of one NVLink / NVSwitch node): like
 * b2_factor_set_linearize_device, and in the same launch the epilogue that finishes a factor also stores its record into
 * the same slot of every peer GPU's result buffer (peer_out[p]: device pointer, valid on THIS device, to peer p's buffer at
 * the offset that corresponds to d_out; entry my_rank is ignored), and the CTA that finishes the last local factor raises
 * peer_flags[p][my_rank] = seq on every GPU p (own included).  b2_exchange_wait enqueues a wait, on the context's stream,
 * until flags[r] == seq for all r < n_peers: after it this GPU's buffer holds every rank's records.  Replaces the single
 * all-reduce of the [F x 128] block the reference would need; callers double-buffer by step parity.
 * seq must be non-zero and change from step to step. */
B2_API b2_status b2_factor_set_linearize_exchange(b2_factor_set* set, const double* d_deltas, double* d_out, double* const* peer_out,
                                                  unsigned int* const* peer_flags, int n_peers, int my_rank, unsigned int seq);
B2_API b2_status b2_exchange_wait(b2_ctx* ctx, const unsigned int* d_flags, int n_peers, unsigned int seq);
/* For a rank that owns no factor in this step: only raises peer_flags[p][my_rank] = seq on every GPU (stream-ordered). */
B2_API b2_status b2_exchange_signal(b2_ctx* ctx, unsigned int* const* peer_flags, int n_peers, int my_rank, unsigned int seq);
/* Peer-buffer setup for that exchange WITHOUT any framework above the ABI (what a C++ host such as ISAM2Ext needs,
 * src/gtsam_points/optimizers/isam2_ext.cpp:110-129): every rank (one process per GPU) creates an exchange object holding its
 * own double-buffered result block [2 x num_records x 128 doubles] + flag words, exports a 64-byte CUDA IPC handle, hands it to
 * its peers over whatever channel it has (MPI, a pipe, shared memory ...), and imports theirs; b2_exchange_enable_peer maps a
 * peer that lives in the SAME process (one thread per GPU) through cudaDeviceEnablePeerAccess instead.  b2_exchange_linearize is
 * then one call per step: the factors of `set` are linearized, their records stored at slots first_slot.. of EVERY rank's
 * block, flags raised, and the kernel itself waits for all ranks' flags before it completes -- when the stream reaches the
 * end of this launch, b2_exchange_records(ex) holds all num_records records of the step.  step must increase by 1 per call,
 * starting at 1, identically on every rank. */
typedef struct b2_exchange b2_exchange;
#define B2_IPC_HANDLE_BYTES 64
B2_API b2_status b2_exchange_create(b2_ctx* ctx, int n_ranks, int my_rank, size_t num_records, b2_exchange** out);
B2_API b2_status b2_exchange_destroy(b2_exchange* ex);
B2_API b2_status b2_exchange_export(const b2_exchange* ex, unsigned char handle[B2_IPC_HANDLE_BYTES]);
B2_API b2_status b2_exchange_import(b2_exchange* ex, int peer_rank, const unsigned char handle[B2_IPC_HANDLE_BYTES]);
B2_API b2_status b2_exchange_enable_peer(b2_exchange* ex, int peer_rank, const b2_exchange* peer_in_this_process);
/* deltas: HOST poses of the set's factors (F x 16), or NULL for a rank that owns no factor in this step (set may then be NULL) */
B2_API b2_status b2_exchange_linearize(b2_exchange* ex, b2_factor_set* set, const double* deltas, size_t first_slot, unsigned int step);
/* device pointer to the [num_records x 128] block of `step` (valid until step + 2 is issued) */
/* Same step with HOST delivery: besides the device block, all num_records records of the step are written to out_records
 * (host, num_records x 128 doubles) -- the CTA that waited for the peers' flags copies them into pinned mapped memory and raises a
 * mapped completion word the call spins on: one launch, no copy operation, no stream synchronisation.  Returns when they are in. */
B2_API b2_status b2_exchange_linearize_host(b2_exchange* ex, b2_factor_set* set, const double* deltas, size_t first_slot, unsigned int step,
                                            double* out_records);
B2_API const double* b2_exchange_records(const b2_exchange* ex, unsigned int step);
/* Stream-ordered rendezvous of the exchange's GPUs (one tiny kernel on the context's stream: raise a word on every GPU, wait for
 * every rank's word): work enqueued after it starts on all GPUs within a few microseconds of each other, whatever the skew
 * between the host processes -- every rank must call it the same number of times.  No host synchronisation.
 * (The reference has no multi-GPU path; NCCL's barrier costs a collective launch and a host-side wait.) */
B2_API b2_status b2_exchange_barrier(b2_exchange* ex);
/* Number of kernel launches issued by this set since creation (for bench.py's gpu_launches). */
B2_API uint64_t b2_factor_set_launch_count(const b2_factor_set* set);

#ifdef __cplusplus
}
#endif
#endif /* B2POINTS_H_ */
