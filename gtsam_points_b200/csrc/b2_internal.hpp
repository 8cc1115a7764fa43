// b2_internal.hpp -- host-side object model behind the C ABI (include/b2points.h).
// Nothing here crosses the library boundary.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b2points.h"

namespace b2 {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
b2_status fail(b2_status st, const char* fmt, ...);

#define B2_CUDA(expr)                                                                                          \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess) {                                                                                   \
      return ::b2::fail(_e == cudaErrorMemoryAllocation ? B2_ERR_OUT_OF_MEMORY : B2_ERR_CUDA, "%s:%d: %s -> %s", \
                        __FILE__, __LINE__, #expr, cudaGetErrorString(_e));                                    \
    }                                                                                                          \
  } while (0)

#define B2_REQUIRE(cond, ...)                                              \
  do {                                                                     \
    if (!(cond)) return ::b2::fail(B2_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define B2_TRY(expr)               \
  do {                             \
    b2_status _s = (expr);         \
    if (_s != B2_OK) return _s;    \
  } while (0)

inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }
inline uint64_t next_pow2(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ---- device data layouts (shared with kernels) -------------------------------------------------

// One hash bucket of a voxel map: integer voxel coordinate -> voxel id.  16 bytes = one LDG.128.
struct alignas(16) VoxelBucket {
  int32_t x, y, z;
  int32_t id;  // -1 = empty
};

// One target record: mean (3), upper-triangular covariance (6), number of points (1).  80 bytes = 5 x LDG.128.
// Used for voxels (VGICP) and for target points (GICP, count = 1).
constexpr int kRecordDoubles = 10;

// kd-tree node, 16 bytes = one LDG.128.
//   internal: thresh = split value, a = index of the left child (right child = a + 1), b = axis (0..2)
//   leaf:     a = first point position (leaf order), b = 4 + number of points in the leaf
struct alignas(16) KdNodeGPU {
  double thresh;
  uint32_t a;
  uint32_t b;
};

}  // namespace b2

// ---- opaque handle definitions -----------------------------------------------------------------

struct b2_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int sm_count = 0;
  // pinned staging for poses (H2D) and results (D2H), grown on demand
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  void* d_stage = nullptr;
  size_t d_stage_bytes = 0;

  // completion word of zero-copy host calls (pinned + mapped) and the device counter behind it
  volatile unsigned int* h_done = nullptr;
  volatile unsigned int* d_done_flag = nullptr;  // device alias of h_done
  unsigned int* d_done_counter = nullptr;
  unsigned int done_seq = 0;

  b2_status ensure_stage(size_t host_bytes, size_t dev_bytes);
};

struct b2_cloud {
  b2_ctx* ctx = nullptr;
  size_t n = 0;
  size_t n_pad = 0;      // plane stride in elements (multiple of 32)
  int point_bytes = 0;   // 4 or 8
  int cov_bytes = 0;     // 0, 4 or 8
  bool reordered = false;
  void* d_points = nullptr;  // 3 planes (x, y, z) of n_pad elements
  void* d_covs = nullptr;    // 6 planes (c00, c01, c02, c11, c12, c22) of n_pad elements
  uint32_t* d_perm = nullptr;  // stored position -> caller index (nullptr if not reordered)
  size_t device_bytes = 0;
};

struct b2_voxelmap {
  b2_ctx* ctx = nullptr;
  double resolution = 0.0;
  double inv_resolution = 0.0;
  size_t num_voxels = 0;
  size_t num_buckets = 0;  // power of two
  b2::VoxelBucket* d_buckets = nullptr;
  double* d_records = nullptr;   // num_voxels x 10
  int32_t* d_coords = nullptr;   // num_voxels x 3 (id order), kept for download
  size_t device_bytes = 0;
  // incremental insertion state (ann/incremental_voxelmap.hpp:13-28): LRU bookkeeping as in the CPU map
  uint32_t* d_lru = nullptr;     // per voxel: lru_counter value of the last insert() that touched it
  size_t capacity = 0;           // voxels the record / coord / lru arrays can hold
  size_t lru_horizon = 10, lru_clear_cycle = 10, lru_counter = 0;
  uint64_t generation = 0;       // bumped whenever the device pointers / voxel ids change: factor sets refresh their descriptors
};

struct b2_kdtree {
  b2_ctx* ctx = nullptr;
  size_t n = 0;
  size_t num_nodes = 0;
  b2::KdNodeGPU* d_nodes = nullptr;
  void* d_leaf_points = nullptr;     // leaf-order records: float4 (x, y, z, 0) if leaf_f32 else 4 doubles (x, y, z, 0)
  bool leaf_f32 = false;             // every coordinate is exactly float32-representable
  size_t n_pad = 0;
  uint32_t* d_leaf_index = nullptr;  // leaf position -> caller index
  size_t device_bytes = 0;
};

enum b2_factor_kind { B2_FACTOR_VGICP = 0, B2_FACTOR_GICP = 1, B2_FACTOR_ICP = 2, B2_FACTOR_ICP_PLANE = 3 };

struct b2_factor {
  b2_ctx* ctx = nullptr;
  b2_factor_kind kind = B2_FACTOR_VGICP;
  const b2_voxelmap* voxelmap = nullptr;
  const b2_cloud* target = nullptr;
  const b2_kdtree* tree = nullptr;
  const b2_cloud* source = nullptr;
  double max_corr_sq = 1.0;
  // integrated_gicp_factor.hpp:103-109: skip re-association while the pose stays within these of the last association point
  double corr_tol_rot = 0.0, corr_tol_trans = 0.0;
  double last_corr_delta[16] = {0};
  bool has_corr = false;  // correspondences have been established at last_corr_delta
  uint64_t params_gen = 0;  // bumped by every setter that changes a value factor sets have copied into their device descriptors
  int32_t* d_corr = nullptr;       // per stored source position: voxel id / target leaf position, -1 = none
  double* d_target_records = nullptr;  // GICP: target records in leaf order (Nt x 10), owned
  double* d_lin_pose = nullptr;        // 16 doubles: the linearization point, written by the linearize kernel's epilogue
  bool linearized = false;
  double lin_delta[16] = {0};
  b2_factor_set* self_set = nullptr;  // lazily created set of size 1 for the single-factor entry points
};
