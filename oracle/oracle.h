/*
 * oracle.h -- C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs may load this library, and there only
 * as the checker or as the timed CPU baseline.  Nothing under gtsam_points_b200/ may
 * include, link or dlopen anything from oracle/.
 *
 * PARITY UNPINNED: the reference (koide3/gtsam_points v1.2.1) ships no golden vectors
 * for H, b or correspondences, and cannot be compiled in this image (needs GTSAM, Eigen,
 * Boost -- none installed, no network).  The restatement in oracle.cpp follows the
 * reference line by line (citations there) and is cross-checked against an independent
 * numpy float64 brute-force implementation (tests/np_ref.py) and against the behavioural
 * gates of the reference's own tests (tests/test_oracle_*.py).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_cloud orc_cloud;
typedef struct orc_voxelmap orc_voxelmap;
typedef struct orc_kdtree orc_kdtree;
typedef struct orc_factor orc_factor;

/* Result of one linearize(): all 6x6 blocks row-major, tangent order [rot(3), trans(3)].
 * H_target_source is J_t^T M J_s (rows: target tangent, cols: source tangent).
 * b_* are the raw sums J^T M r (the HessianFactor receives -b, see
 * src/gtsam_points/factors/integrated_matching_cost_factor.cpp:46-52). */
typedef struct orc_linearized {
  double H_target[36];
  double H_source[36];
  double H_target_source[36];
  double b_target[6];
  double b_source[6];
  double error;
  double num_inliers;
} orc_linearized;

/* cloud: xyz = N x 3 doubles, cov3x3 = N x 9 doubles (row-major, symmetric); cov may be NULL.
 * Stored internally exactly like the reference: Vector4d (x,y,z,1), Matrix4d with row/col 3 zero
 * (include/gtsam_points/types/point_cloud.hpp:106-108). */
orc_cloud* orc_cloud_create(const double* xyz, const double* cov3x3, size_t n);
void orc_cloud_destroy(orc_cloud*);
size_t orc_cloud_size(const orc_cloud*);

/* GaussianVoxelMapCPU (src/gtsam_points/types/gaussian_voxelmap_cpu.cpp) */
orc_voxelmap* orc_voxelmap_create(double resolution);
void orc_voxelmap_destroy(orc_voxelmap*);
void orc_voxelmap_set_lru(orc_voxelmap*, int lru_horizon, int lru_clear_cycle);
void orc_voxelmap_insert(orc_voxelmap*, const orc_cloud*);
size_t orc_voxelmap_num_voxels(const orc_voxelmap*);
/* coords V x 3 int32, means V x 3, covs V x 9 (row-major 3x3), num_points V (int32); any may be NULL */
void orc_voxelmap_export(const orc_voxelmap*, int32_t* coords, double* means, double* covs, int32_t* num_points);
/* voxel_coord + lookup_voxel_index for n points (xyz n x 3); out idx n (int32, -1 = none) */
void orc_voxelmap_lookup(const orc_voxelmap*, const double* xyz, size_t n, int32_t* out_idx);

/* KdTree2 / UnsafeKdTree (include/gtsam_points/ann/small_kdtree.hpp) */
orc_kdtree* orc_kdtree_create(const orc_cloud* target, int build_num_threads);
void orc_kdtree_destroy(orc_kdtree*);
/* queries nq x 3; out_idx nq x k (uint64; SIZE_MAX = invalid), out_sqd nq x k; returns per-query found counts in out_found (may be NULL) */
void orc_kdtree_knn(const orc_kdtree*, const double* queries, size_t nq, int k, double max_sq_dist, uint64_t* out_idx, double* out_sqd,
                    int32_t* out_found, int num_threads);
size_t orc_kdtree_num_nodes(const orc_kdtree*);

/* estimate_covariances with EIG regularisation (src/gtsam_points/features/covariance_estimation.cpp:18-77):
 * out_cov3x3 = n x 9; eigen_values3 in ascending-eigenvalue order (reference default 1e-3, 1, 1). */
void orc_estimate_covariances(const orc_cloud* cloud, int k_neighbors, const double* eigen_values3, int num_threads, double* out_cov3x3);

/* IntegratedVGICPFactor_ / IntegratedGICPFactor_ in FusedCovCacheMode::FULL */
orc_factor* orc_vgicp_create(const orc_voxelmap* target, const orc_cloud* source);
orc_factor* orc_gicp_create(const orc_cloud* target, const orc_kdtree* tree, const orc_cloud* source);
/* IntegratedICPFactor_ (include/gtsam_points/factors/impl/integrated_icp_factor_impl.hpp:131-248); point-to-plane needs
 * target normals (orc_cloud_set_normals, n x 3) */
orc_factor* orc_icp_create(const orc_cloud* target, const orc_kdtree* tree, const orc_cloud* source, int use_point_to_plane);
void orc_cloud_set_normals(orc_cloud*, const double* nxyz);
/* FusedCovCacheMode 0 FULL / 1 COMPACT / 2 NONE (integrated_gicp_factor.hpp:20-24,104-105) */
void orc_factor_set_fused_cov_cache_mode(orc_factor*, int mode);
/* integrated_gicp_factor.hpp:106-109, impl:135-147 */
void orc_factor_set_correspondence_update_tolerance(orc_factor*, double angle, double trans);
/* overlap (src/gtsam_points/types/gaussian_voxelmap_cpu_funcs.cpp:126-173) */
double orc_overlap(const orc_voxelmap* target, const orc_cloud* source, const double* T_target_source_rm16);
double orc_overlap_multi(const orc_voxelmap* const* targets, int num_targets, const orc_cloud* source, const double* Ts_rm16);
/* save_compact / load (src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:79-135, types/gaussian_voxel_data.hpp:11-54) */
int orc_voxelmap_save_compact(const orc_voxelmap*, const char* path);
orc_voxelmap* orc_voxelmap_load(const char* path);
/* merge_frames (gaussian_voxelmap_cpu_funcs.cpp:25-113); out arrays sized for the total number of input points */
size_t orc_merge_frames(const double* poses_rm16, const orc_cloud* const* frames, int num_frames, double downsample_resolution, double* out_xyz, double* out_cov3x3);
void orc_factor_destroy(orc_factor*);
void orc_factor_set_num_threads(orc_factor*, int n);
void orc_factor_set_max_correspondence_distance(orc_factor*, double dist); /* GICP only */
/* delta: 4x4 row-major T_target^-1 * T_source.  linearize = update_correspondences + evaluate(H,b). */
void orc_factor_linearize(orc_factor*, const double* delta_rm16, orc_linearized* out);
/* error at delta re-using correspondences + Mahalanobis frozen at the last linearize */
double orc_factor_error(orc_factor*, const double* delta_rm16);
/* out n (int64): VGICP voxel id / GICP target index, -1 = none */
void orc_factor_correspondences(const orc_factor*, int64_t* out);

/* delta = T_target^-1 * T_source for 4x4 row-major rigid transforms */
void orc_calc_delta(const double* T_target_rm16, const double* T_source_rm16, double* delta_rm16);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
