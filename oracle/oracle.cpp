/*
 * oracle.cpp -- dependency-free CPU restatement of the gtsam_points scan-matching hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).  PARITY UNPINNED (see oracle.h).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference
 * root, koide3/gtsam_points v1.2.1).  Storage mirrors the reference so that this file also
 * serves as the timed CPU baseline: 32-byte (x,y,z,1) points, 128-byte column-major 4x4
 * covariances, std::unordered_map + XOR spatial hash -> flat_voxels, FusedCovCacheMode::FULL
 * (two passes), kd-tree with <=20-point leaves, `omp parallel for schedule(guided, 8)`.
 *
 * Third-party arithmetic that is not in the reference tree is restated from its published
 * definition: Eigen fixed-size products (coefficient sums in index order), Eigen 3x3
 * inverse (cofactors * 1/det), gtsam::SO3::Hat (skew-symmetric matrix), Isometry3d * Vector4d.
 * Build with -ffp-contract=off: no FMA contraction, like the reference's default x86-64 build
 * (CMakeLists.txt:21, BUILD_WITH_MARCH_NATIVE=OFF).
 */
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <unordered_map>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// Minimal fixed-size linear algebra with Eigen's storage conventions (column-major).
// ---------------------------------------------------------------------------------------------
struct alignas(32) Vec4 {
  double v[4];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};

struct alignas(32) Mat4 {  // column-major: m[c * 4 + r]
  double m[16];
  double& operator()(int r, int c) { return m[c * 4 + r]; }
  double operator()(int r, int c) const { return m[c * 4 + r]; }
  void set_zero() { std::memset(m, 0, sizeof(m)); }
};

struct Mat66 {  // column-major 6x6
  double m[36];
  double& operator()(int r, int c) { return m[c * 6 + r]; }
  double operator()(int r, int c) const { return m[c * 6 + r]; }
  void set_zero() { std::memset(m, 0, sizeof(m)); }
  Mat66& operator+=(const Mat66& o) {
    for (int i = 0; i < 36; i++) m[i] += o.m[i];
    return *this;
  }
};

struct Vec6 {
  double v[6];
  void set_zero() { std::memset(v, 0, sizeof(v)); }
  Vec6& operator+=(const Vec6& o) {
    for (int i = 0; i < 6; i++) v[i] += o.v[i];
    return *this;
  }
};

struct Mat46 {  // 4 rows x 6 cols, column-major
  double m[24];
  double& operator()(int r, int c) { return m[c * 4 + r]; }
  double operator()(int r, int c) const { return m[c * 4 + r]; }
};

struct Mat64 {  // 6 rows x 4 cols, column-major
  double m[24];
  double& operator()(int r, int c) { return m[c * 6 + r]; }
  double operator()(int r, int c) const { return m[c * 6 + r]; }
};

inline Mat4 mul44(const Mat4& a, const Mat4& b) {
  Mat4 r;
  for (int c = 0; c < 4; c++) {
    for (int i = 0; i < 4; i++) {
      double s = a(i, 0) * b(0, c);
      for (int k = 1; k < 4; k++) s += a(i, k) * b(k, c);
      r(i, c) = s;
    }
  }
  return r;
}

inline Mat4 transpose44(const Mat4& a) {
  Mat4 r;
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 4; i++) r(i, c) = a(c, i);
  return r;
}

// Isometry3d * Vector4d: top three rows = affine(3x4) * v, last row copied (Eigen Transform.h,
// transform_right_product_impl for a (Dim+1)-vector).  Coefficient sum in index order.
inline Vec4 transform_point(const Mat4& d, const Vec4& p) {
  Vec4 r;
  for (int i = 0; i < 3; i++) {
    r[i] = ((d(i, 0) * p[0] + d(i, 1) * p[1]) + d(i, 2) * p[2]) + d(i, 3) * p[3];
  }
  r[3] = p[3];
  return r;
}

// Eigen fixed-size 3x3 inverse (Eigen/src/LU/InverseImpl.h, compute_inverse<.., 3>):
// cofactor(i,j) = m(i1,j1) m(i2,j2) - m(i1,j2) m(i2,j1) with i1=(i+1)%3, i2=(i+2)%3;
// det = sum_i cofactor(i,0) m(i,0) ; inverse(i,j) = cofactor(j,i) / det (multiplication by 1/det).
inline double cofactor3(const double a[3][3], int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return a[i1][j1] * a[i2][j2] - a[i1][j2] * a[i2][j1];
}

inline void inverse3(const double a[3][3], double inv[3][3]) {
  const double c00 = cofactor3(a, 0, 0), c10 = cofactor3(a, 1, 0), c20 = cofactor3(a, 2, 0);
  const double det = (c00 * a[0][0] + c10 * a[1][0]) + c20 * a[2][0];
  const double invdet = 1.0 / det;
  inv[0][0] = c00 * invdet;
  inv[0][1] = c10 * invdet;
  inv[0][2] = c20 * invdet;
  inv[1][0] = cofactor3(a, 0, 1) * invdet;
  inv[1][1] = cofactor3(a, 1, 1) * invdet;
  inv[1][2] = cofactor3(a, 2, 1) * invdet;
  inv[2][0] = cofactor3(a, 0, 2) * invdet;
  inv[2][1] = cofactor3(a, 1, 2) * invdet;
  inv[2][2] = cofactor3(a, 2, 2) * invdet;
}

// ---------------------------------------------------------------------------------------------
// util/fast_floor.hpp:12-15 : truncate, then subtract one where the value is below the truncation.
// ---------------------------------------------------------------------------------------------
struct Vec3i {
  int32_t x, y, z;
  bool operator==(const Vec3i& o) const { return x == o.x && y == o.y && z == o.z; }
};

inline int32_t fast_floor1(double v) {
  const int32_t n = static_cast<int32_t>(v);
  return n - (v < static_cast<double>(n) ? 1 : 0);
}

// util/vector3i_hash.hpp:29-37 (XORVector3iHash): int -> size_t conversion is sign-extending, products wrap.
struct XORVec3iHash {
  size_t operator()(const Vec3i& x) const {
    const size_t p1 = 9132043225175502913ull;
    const size_t p2 = 7277549399757405689ull;
    const size_t p3 = 6673468629021231217ull;
    return static_cast<size_t>((static_cast<size_t>(static_cast<int64_t>(x.x)) * p1) ^ (static_cast<size_t>(static_cast<int64_t>(x.y)) * p2) ^
                               (static_cast<size_t>(static_cast<int64_t>(x.z)) * p3));
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// types/point_cloud.hpp:103-118 : raw-pointer SoA cloud (CPU side: Vector4d points, Matrix4d covs)
// ---------------------------------------------------------------------------------------------
struct orc_cloud {
  std::vector<Vec4> points;
  std::vector<Mat4> covs;
  std::vector<Vec4> normals;  // (nx, ny, nz, 0), types/point_cloud.hpp:107
  bool covs_given = false;
  bool has_normals() const { return normals.size() == points.size() && !points.empty(); }
  size_t size() const { return points.size(); }
  bool has_covs() const { return covs_given; }
};

// ---------------------------------------------------------------------------------------------
// types/gaussian_voxelmap_cpu.hpp:13-52, src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:23-77,
// ann/incremental_voxelmap.hpp:13-28, ann/impl/incremental_voxelmap_impl.hpp:13-68
// ---------------------------------------------------------------------------------------------
namespace {

struct VoxelInfo {
  size_t lru;
  Vec3i coord;
};

struct GaussianVoxel {
  bool finalized = false;
  size_t num_points = 0;
  Vec4 mean = {{0, 0, 0, 0}};
  Mat4 cov;
  double intensity = 0.0;
  GaussianVoxel() { cov.set_zero(); }

  // gaussian_voxelmap_cpu.cpp:23-37
  void add(const orc_cloud& points, size_t i) {
    if (finalized) {
      finalized = false;
      for (int k = 0; k < 4; k++) mean[k] *= static_cast<double>(num_points);
      for (int k = 0; k < 16; k++) cov.m[k] *= static_cast<double>(num_points);
    }
    num_points++;
    for (int k = 0; k < 4; k++) mean[k] += points.points[i][k];
    for (int k = 0; k < 16; k++) cov.m[k] += points.covs[i].m[k];
  }

  // gaussian_voxelmap_cpu.cpp:39-47
  void finalize() {
    if (finalized) return;
    for (int k = 0; k < 4; k++) mean[k] /= static_cast<double>(num_points);
    for (int k = 0; k < 16; k++) cov.m[k] /= static_cast<double>(num_points);
    finalized = true;
  }
};

}  // namespace

struct orc_voxelmap {
  double inv_leaf_size;
  size_t lru_horizon = 10;
  size_t lru_clear_cycle = 10;
  size_t lru_counter = 0;
  std::vector<std::shared_ptr<std::pair<VoxelInfo, GaussianVoxel>>> flat_voxels;
  std::unordered_map<Vec3i, size_t, XORVec3iHash> voxels;

  explicit orc_voxelmap(double leaf) : inv_leaf_size(1.0 / leaf) {}

  // gaussian_voxelmap_cpu.cpp:59-61 : fast_floor(x * inv_leaf_size).head<3>()
  Vec3i voxel_coord(const Vec4& x) const {
    return Vec3i{fast_floor1(x[0] * inv_leaf_size), fast_floor1(x[1] * inv_leaf_size), fast_floor1(x[2] * inv_leaf_size)};
  }

  // gaussian_voxelmap_cpu.cpp:63-69
  int lookup_voxel_index(const Vec3i& coord) const {
    auto found = voxels.find(coord);
    if (found == voxels.end()) return -1;
    return static_cast<int>(found->second);
  }

  // gaussian_voxelmap_cpu.cpp:71-73
  const GaussianVoxel& lookup_voxel(int id) const { return flat_voxels[id]->second; }

  // ann/impl/incremental_voxelmap_impl.hpp:31-68
  void insert(const orc_cloud& points) {
    for (size_t i = 0; i < points.size(); i++) {
      const Vec3i coord = voxel_coord(points.points[i]);
      auto found = voxels.find(coord);
      if (found == voxels.end()) {
        auto voxel = std::make_shared<std::pair<VoxelInfo, GaussianVoxel>>(VoxelInfo{lru_counter, coord}, GaussianVoxel());
        found = voxels.emplace_hint(found, coord, flat_voxels.size());
        flat_voxels.emplace_back(voxel);
      }
      auto& entry = *flat_voxels[found->second];
      entry.first.lru = lru_counter;
      entry.second.add(points, i);
    }

    if ((++lru_counter) % lru_clear_cycle == 0) {
      auto remove_counter = std::remove_if(flat_voxels.begin(), flat_voxels.end(), [&](const auto& voxel) {
        return voxel->first.lru + lru_horizon < lru_counter;
      });
      flat_voxels.erase(remove_counter, flat_voxels.end());
      voxels.clear();
      for (size_t i = 0; i < flat_voxels.size(); i++) voxels[flat_voxels[i]->first.coord] = i;
    }

    for (auto& voxel : flat_voxels) voxel->second.finalize();
  }
};

// ---------------------------------------------------------------------------------------------
// ann/knn_result.hpp:11-117 (KnnSetting, KnnResult<-1> dynamic container)
// ---------------------------------------------------------------------------------------------
namespace {

struct KnnResultDyn {
  static constexpr size_t INVALID = std::numeric_limits<size_t>::max();
  int capacity;
  int num_found_neighbors = 0;
  size_t* indices;
  double* distances;

  // knn_result.hpp:44-72 : buffers pre-filled with INVALID / max_sq_dist
  KnnResultDyn(size_t* idx, double* dist, int k, double max_sq_dist) : capacity(k), indices(idx), distances(dist) {
    std::fill(indices, indices + capacity, INVALID);
    std::fill(distances, distances + capacity, max_sq_dist);
  }
  size_t num_found() const { return num_found_neighbors; }
  double worst_distance() const { return distances[capacity - 1]; }

  // knn_result.hpp:89-109 : strict '<' against the current worst, insertion sort
  void push(size_t index, double distance) {
    if (distance >= worst_distance()) return;
    int insert_loc = std::min<int>(num_found_neighbors, capacity - 1);
    for (; insert_loc > 0 && distance < distances[insert_loc - 1]; insert_loc--) {
      indices[insert_loc] = indices[insert_loc - 1];
      distances[insert_loc] = distances[insert_loc - 1];
    }
    indices[insert_loc] = index;
    distances[insert_loc] = distance;
    num_found_neighbors = std::min<int>(num_found_neighbors + 1, capacity);
  }
};

// knn_result.hpp:11-23 with epsilon = 0 and max_nn = INT_MAX (the kNN path never early-terminates,
// except when the worst distance becomes < 0, which cannot happen).
inline bool fulfilled(const KnnResultDyn& r) {
  return r.worst_distance() < 0.0 || r.num_found() >= static_cast<size_t>(std::numeric_limits<int>::max());
}

constexpr uint32_t INVALID_NODE = std::numeric_limits<uint32_t>::max();

// ann/small_kdtree.hpp:106-121
struct KdNode {
  union {
    struct {
      uint32_t first, last;
    } lr;
    struct {
      int axis;
      double thresh;
    } sub;
  } node_type;
  uint32_t left = INVALID_NODE;
  uint32_t right = INVALID_NODE;
};

}  // namespace

// ann/small_kdtree.hpp:351-533 (UnsafeKdTree), builders :124-274, projection :58-100
struct orc_kdtree {
  const orc_cloud* points;
  std::vector<size_t> indices;
  uint32_t root = 0;
  std::vector<KdNode> nodes;
  int max_leaf_size = 20;
  int max_scan_count = 128;

  // small_kdtree.hpp:77-96 : axis of largest variance over <=128 strided samples
  int find_axis(const size_t* first, const size_t* last) const {
    const size_t N = last - first;
    Vec4 sum_pt = {{0, 0, 0, 0}}, sum_sq = {{0, 0, 0, 0}};
    const size_t step = N < static_cast<size_t>(max_scan_count) ? 1 : N / max_scan_count;
    const size_t num_steps = N / step;
    for (size_t i = 0; i < num_steps; i++) {
      const Vec4& pt = points->points[first[step * i]];
      for (int k = 0; k < 4; k++) {
        sum_pt[k] += pt[k];
        sum_sq[k] += pt[k] * pt[k];
      }
    }
    double var[4];
    for (int k = 0; k < 4; k++) {
      const double mean = sum_pt[k] / sum_pt[3];
      var[k] = sum_sq[k] - mean * sum_pt[k];
    }
    return var[0] > var[1] ? (var[0] > var[2] ? 0 : 2) : (var[1] > var[2] ? 1 : 2);
  }

  // small_kdtree.hpp:224-268 (KdTreeBuilderOMP::create_node; :141-177 is the serial twin)
  uint32_t create_node(std::atomic<uint64_t>& node_count, size_t* global_first, size_t* first, size_t* last) {
    const size_t N = last - first;
    const uint32_t node_index = static_cast<uint32_t>(node_count++);
    KdNode& node = nodes[node_index];

    if (N <= static_cast<size_t>(max_leaf_size)) {
      node.node_type.lr.first = static_cast<uint32_t>(first - global_first);
      node.node_type.lr.last = static_cast<uint32_t>(last - global_first);
      return node_index;
    }

    const int axis = find_axis(first, last);
    size_t* median_itr = first + N / 2;
    std::nth_element(first, median_itr, last, [&](size_t i, size_t j) { return points->points[i][axis] < points->points[j][axis]; });

    node.node_type.sub.axis = axis;
    node.node_type.sub.thresh = points->points[*median_itr][axis];

#pragma omp task default(shared) if (N > 512)
    node.left = create_node(node_count, global_first, first, median_itr);
#pragma omp task default(shared) if (N > 512)
    node.right = create_node(node_count, global_first, median_itr, last);
#pragma omp taskwait
    return node_index;
  }

  // small_kdtree.hpp:196-217
  void build(int num_threads) {
    indices.resize(points->size());
    std::iota(indices.begin(), indices.end(), 0);
    std::atomic<uint64_t> node_count{0};
    nodes.resize(points->size());
    if (points->size() == 0) {
      nodes.clear();
      return;
    }
#pragma omp parallel num_threads(num_threads)
    {
#pragma omp single nowait
      { root = create_node(node_count, indices.data(), indices.data(), indices.data() + indices.size()); }
    }
    nodes.resize(node_count);
  }

  // small_kdtree.hpp:436-476
  bool knn_search(const Vec4& query, uint32_t node_index, KnnResultDyn& result) const {
    const KdNode& node = nodes[node_index];
    if (node.left == INVALID_NODE) {
      for (size_t i = node.node_type.lr.first; i < node.node_type.lr.last; i++) {
        const Vec4& p = points->points[indices[i]];
        // (p - query).squaredNorm() over four components (the w's cancel)
        const double d0 = p[0] - query[0], d1 = p[1] - query[1], d2 = p[2] - query[2], d3 = p[3] - query[3];
        const double sq_dist = ((d0 * d0 + d1 * d1) + d2 * d2) + d3 * d3;
        result.push(indices[i], sq_dist);
      }
      return !fulfilled(result);
    }

    const double val = query[node.node_type.sub.axis];
    const double diff = val - node.node_type.sub.thresh;
    const double cut_sq_dist = diff * diff;

    uint32_t best_child, other_child;
    if (diff < 0.0) {
      best_child = node.left;
      other_child = node.right;
    } else {
      best_child = node.right;
      other_child = node.left;
    }

    if (!knn_search(query, best_child, result)) return false;
    if (result.worst_distance() > cut_sq_dist) return knn_search(query, other_child, result);
    return true;
  }

  // ann/kdtree2.hpp:52-61 -> small_kdtree.hpp:389-394 (dynamic KnnResult<-1>)
  size_t knn_search(const double* pt, size_t k, size_t* k_indices, double* k_sq_dists, double max_sq_dist) const {
    const Vec4 query = {{pt[0], pt[1], pt[2], 1.0}};
    KnnResultDyn result(k_indices, k_sq_dists, static_cast<int>(k), max_sq_dist);
    if (nodes.empty()) return 0;
    knn_search(query, root, result);
    return result.num_found();
  }
};

// ---------------------------------------------------------------------------------------------
// factors/impl/scan_matching_reduction.hpp:16-68 (scan_matching_reduce_omp)
// ---------------------------------------------------------------------------------------------
namespace {

template <typename F>
double scan_matching_reduce_omp(const F& f, int num_points, int num_threads, Mat66* H_target, Mat66* H_source, Mat66* H_target_source, Vec6* b_target,
                                Vec6* b_source) {
  double sum_errors = 0.0;
  const int num_Hs = H_target ? num_threads : 0;
  Mat66 zero66;
  zero66.set_zero();
  Vec6 zero6;
  zero6.set_zero();
  std::vector<Mat66> Hs_target(num_Hs, zero66), Hs_source(num_Hs, zero66), Hs_target_source(num_Hs, zero66);
  std::vector<Vec6> bs_target(num_Hs, zero6), bs_source(num_Hs, zero6);

#pragma omp parallel for num_threads(num_threads) schedule(guided, 8) reduction(+ : sum_errors)
  for (int i = 0; i < num_points; i++) {
    int thread_num = 0;
#ifdef _OPENMP
    thread_num = omp_get_thread_num();
#endif
    double error = 0.0;
    if (Hs_target.empty()) {
      error = f(i, nullptr, nullptr, nullptr, nullptr, nullptr);
    } else {
      error = f(i, &Hs_target[thread_num], &Hs_source[thread_num], &Hs_target_source[thread_num], &bs_target[thread_num], &bs_source[thread_num]);
    }
    sum_errors += error;
  }

  if (H_target) {
    *H_target = Hs_target[0];
    *H_source = Hs_source[0];
    *H_target_source = Hs_target_source[0];
    *b_target = bs_target[0];
    *b_source = bs_source[0];
    for (int i = 1; i < num_threads; i++) {
      *H_target += Hs_target[i];
      *H_source += Hs_source[i];
      *H_target_source += Hs_target_source[i];
      *b_target += bs_target[i];
      *b_source += bs_source[i];
    }
  }
  return sum_errors;
}

// gtsam::SO3::Hat(v): the skew-symmetric matrix [v]x (gtsam/geometry/SO3.cpp); written into a 3x3 block.
inline void hat3(const double v[3], double h[3][3]) {
  h[0][0] = 0.0;
  h[0][1] = -v[2];
  h[0][2] = v[1];
  h[1][0] = v[2];
  h[1][1] = 0.0;
  h[1][2] = -v[0];
  h[2][0] = -v[1];
  h[2][1] = v[0];
  h[2][2] = 0.0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// IntegratedVGICPFactor_ (factors/impl/integrated_vgicp_factor_impl.hpp:99-257) and
// IntegratedGICPFactor_  (factors/impl/integrated_gicp_factor_impl.hpp:132-296), FULL cache mode.
// ---------------------------------------------------------------------------------------------
struct orc_factor {
  bool is_vgicp;
  int icp_mode = 0;  // 0: (V)GICP, 1: point-to-point ICP, 2: point-to-plane ICP (integrated_icp_factor_impl.hpp)
  int num_threads = 1;
  double max_correspondence_distance_sq = 1.0;  // integrated_gicp_factor_impl.hpp:30
  int cache_mode = 0;                           // FusedCovCacheMode: 0 FULL, 1 COMPACT, 2 NONE (integrated_gicp_factor.hpp:20-24)
  double correspondence_update_tolerance_rot = 0.0, correspondence_update_tolerance_trans = 0.0;  // integrated_gicp_factor_impl.hpp:31-32

  const orc_voxelmap* target_voxels = nullptr;
  const orc_cloud* target = nullptr;
  const orc_kdtree* target_tree = nullptr;
  const orc_cloud* source = nullptr;

  Mat4 linearization_point;
  Mat4 last_correspondence_point;
  std::vector<const GaussianVoxel*> corr_voxels;  // VGICP: integrated_vgicp_factor.hpp:107
  std::vector<long> corr_indices;                 // GICP:  integrated_gicp_factor.hpp:147
  std::vector<Mat4> mahalanobis_full;
  struct Compact6 {
    float v[6];
  };
  std::vector<Compact6> mahalanobis_compact;  // util/compact.hpp:9-25: (00, 10, 11, 20, 21, 22) as float

  // 3x3 inverse((cov_B + delta cov_A delta^T).topLeft3x3): vgicp_impl:139-143 / gicp_impl:177-183
  static void fused_mahalanobis3(const Mat4& delta, const Mat4& cov_B, const Mat4& cov_A, double inv[3][3]) {
    const Mat4 dC = mul44(delta, cov_A);
    const Mat4 dCdT = mul44(dC, transpose44(delta));
    double rcr[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) rcr[r][c] = cov_B(r, c) + dCdT(r, c);
    inverse3(rcr, inv);
  }
  static void fused_mahalanobis(const Mat4& delta, const Mat4& cov_B, const Mat4& cov_A, Mat4& out) {
    double inv[3][3];
    fused_mahalanobis3(delta, cov_B, cov_A, inv);
    out.set_zero();
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) out(r, c) = inv[r][c];
  }
  // compact_cov / uncompact_cov (util/compact.hpp:9-25)
  static Compact6 compact_cov(const double m[3][3]) {
    return Compact6{{static_cast<float>(m[0][0]), static_cast<float>(m[1][0]), static_cast<float>(m[1][1]), static_cast<float>(m[2][0]), static_cast<float>(m[2][1]),
                     static_cast<float>(m[2][2])}};
  }
  static void uncompact_cov(const Compact6& c, Mat4& out) {
    out.set_zero();
    out(0, 0) = c.v[0];
    out(1, 0) = out(0, 1) = c.v[1];
    out(1, 1) = c.v[2];
    out(2, 0) = out(0, 2) = c.v[3];
    out(2, 1) = out(1, 2) = c.v[4];
    out(2, 2) = c.v[5];
  }
  void resize_caches(size_t N) {
    if (cache_mode == 0) mahalanobis_full.resize(N);
    if (cache_mode == 1) mahalanobis_compact.resize(N);
  }
  // the cache entry of point i (both factor families: vgicp_impl:131-163, gicp_impl:174-199)
  void store_mahalanobis(size_t i, bool valid, const Mat4& delta, const Mat4* cov_B) {
    if (cache_mode == 0) {
      if (!valid)
        mahalanobis_full[i].set_zero();
      else
        fused_mahalanobis(delta, *cov_B, source->covs[i], mahalanobis_full[i]);
    } else if (cache_mode == 1) {
      if (!valid) {
        mahalanobis_compact[i] = Compact6{{0, 0, 0, 0, 0, 0}};
      } else {
        double inv[3][3];
        fused_mahalanobis3(delta, *cov_B, source->covs[i], inv);
        mahalanobis_compact[i] = compact_cov(inv);
      }
    }
  }

  // Eigen::AngleAxisd(R).angle() as Eigen computes it (Geometry/AngleAxis.h: via Quaternion(R), angle = 2 atan2(|vec|, |w|))
  static double rotation_angle(const Mat4& T) {
    const double m00 = T(0, 0), m11 = T(1, 1), m22 = T(2, 2);
    double w, x, y, z;
    const double t = m00 + m11 + m22;
    if (t > 0.0) {
      double tt = std::sqrt(t + 1.0);
      w = 0.5 * tt;
      tt = 0.5 / tt;
      x = (T(2, 1) - T(1, 2)) * tt;
      y = (T(0, 2) - T(2, 0)) * tt;
      z = (T(1, 0) - T(0, 1)) * tt;
    } else {
      int i = 0;
      if (m11 > m00) i = 1;
      if (m22 > T(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      double tt = std::sqrt(T(i, i) - T(j, j) - T(k, k) + 1.0);
      double q[3];
      q[i] = 0.5 * tt;
      tt = 0.5 / tt;
      w = (T(k, j) - T(j, k)) * tt;
      q[j] = (T(j, i) + T(i, j)) * tt;
      q[k] = (T(k, i) + T(i, k)) * tt;
      x = q[0], y = q[1], z = q[2];
    }
    const double n = std::sqrt(x * x + y * y + z * z);
    return n != 0.0 ? 2.0 * std::atan2(n, std::fabs(w)) : 0.0;
  }
  static Mat4 rigid_inverse(const Mat4& T) {
    Mat4 inv;
    inv.set_zero();
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) inv(r, c) = T(c, r);
    for (int r = 0; r < 3; r++) inv(r, 3) = -(inv(r, 0) * T(0, 3) + inv(r, 1) * T(1, 3) + inv(r, 2) * T(2, 3));
    inv(3, 3) = 1.0;
    return inv;
  }
  // gicp_impl:135-147 / icp_impl:132-140: skip re-association when the pose moved less than the tolerances
  bool want_correspondence_update(const Mat4& delta) const {
    if (corr_indices.size() == source->size() && (correspondence_update_tolerance_trans > 0.0 || correspondence_update_tolerance_rot > 0.0)) {
      const Mat4 diff = mul44(rigid_inverse(delta), last_correspondence_point);
      const double diff_rot = rotation_angle(diff);
      const double diff_trans = std::sqrt(diff(0, 3) * diff(0, 3) + diff(1, 3) * diff(1, 3) + diff(2, 3) * diff(2, 3));
      if (diff_rot < correspondence_update_tolerance_rot && diff_trans < correspondence_update_tolerance_trans) return false;
    }
    return true;
  }

  // vgicp_impl:99-172
  void update_correspondences_vgicp(const Mat4& delta) {
    linearization_point = delta;
    const int N = static_cast<int>(source->size());
    corr_voxels.resize(N);
    resize_caches(N);

#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
    for (int i = 0; i < N; i++) {
      const Vec4 pt = transform_point(delta, source->points[i]);
      const Vec3i coord = target_voxels->voxel_coord(pt);
      const int voxel_id = target_voxels->lookup_voxel_index(coord);
      if (voxel_id < 0) {
        corr_voxels[i] = nullptr;
        store_mahalanobis(i, false, delta, nullptr);
      } else {
        const GaussianVoxel* voxel = &target_voxels->lookup_voxel(voxel_id);
        corr_voxels[i] = voxel;
        store_mahalanobis(i, true, delta, &voxel->cov);
      }
    }
  }

  // gicp_impl:132-215
  void update_correspondences_gicp(const Mat4& delta) {
    linearization_point = delta;
    const bool do_update = want_correspondence_update(delta);
    if (do_update) last_correspondence_point = delta;
    const int N = static_cast<int>(source->size());
    corr_indices.resize(N);
    resize_caches(N);

#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
    for (int i = 0; i < N; i++) {
      if (do_update) {
        const Vec4 pt = transform_point(delta, source->points[i]);
        size_t k_index = static_cast<size_t>(-1);
        double k_sq_dist = -1;
        const size_t num_found = target_tree->knn_search(pt.v, 1, &k_index, &k_sq_dist, max_correspondence_distance_sq);
        corr_indices[i] = (num_found && k_sq_dist < max_correspondence_distance_sq) ? static_cast<long>(k_index) : -1;
      }
      store_mahalanobis(i, corr_indices[i] >= 0, delta, corr_indices[i] >= 0 ? &target->covs[corr_indices[i]] : nullptr);
    }
  }

  // icp_impl:131-182 (no Mahalanobis cache; the tolerance check returns before anything is touched)
  void update_correspondences_icp(const Mat4& delta) {
    if (!want_correspondence_update(delta)) return;
    last_correspondence_point = delta;
    const int N = static_cast<int>(source->size());
    corr_indices.resize(N);
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
    for (int i = 0; i < N; i++) {
      const Vec4 pt = transform_point(delta, source->points[i]);
      size_t k_index = static_cast<size_t>(-1);
      double k_sq_dist = -1;
      const size_t num_found = target_tree->knn_search(pt.v, 1, &k_index, &k_sq_dist, max_correspondence_distance_sq);
      corr_indices[i] = (num_found == 0 || k_sq_dist > max_correspondence_distance_sq) ? -1 : static_cast<long>(k_index);
    }
  }

  void update_correspondences(const Mat4& delta) {
    if (is_vgicp)
      update_correspondences_vgicp(delta);
    else if (icp_mode)
      update_correspondences_icp(delta);
    else
      update_correspondences_gicp(delta);
  }

  // vgicp_impl:175-257 / gicp_impl:218-296
  double evaluate(const Mat4& delta, Mat66* H_target, Mat66* H_source, Mat66* H_target_source, Vec6* b_target, Vec6* b_source) {
    const size_t ncorr = is_vgicp ? corr_voxels.size() : corr_indices.size();
    if (ncorr != source->size()) update_correspondences(delta);

    const auto perpoint_task = [&](int i, Mat66* H_target, Mat66* H_source, Mat66* H_target_source, Vec6* b_target, Vec6* b_source) -> double {
      const Vec4* mean_B;
      if (is_vgicp) {
        const GaussianVoxel* voxel = corr_voxels[i];
        if (voxel == nullptr) return 0.0;
        mean_B = &voxel->mean;
      } else {
        const long target_index = corr_indices[i];
        if (target_index < 0) return 0.0;
        mean_B = &target->points[target_index];
      }
      const Vec4& mean_A = source->points[i];

      const Vec4 transed_mean_A = transform_point(delta, mean_A);
      Vec4 residual;
      for (int k = 0; k < 4; k++) residual[k] = (*mean_B)[k] - transed_mean_A[k];

      // vgicp_impl:209-224 / gicp_impl:252-266: the cached, uncompacted or (NONE) re-derived fused Mahalanobis matrix
      Mat4 mahalanobis;
      const Vec4* normal_B = nullptr;
      if (icp_mode) {
        // icp_impl:205-248: residual^T residual, J^T J -- written as M = I; point-to-plane scales residual and the Jacobian
        // rows by the target normal (normal_B.array() * residual.array(), normal_B.asDiagonal() * J)
        mahalanobis.set_zero();
        for (int k = 0; k < 4; k++) mahalanobis(k, k) = 1.0;
        if (icp_mode == 2) {
          normal_B = &target->normals[corr_indices[i]];
          for (int k = 0; k < 4; k++) residual[k] = (*normal_B)[k] * residual[k];
        }
      } else if (cache_mode == 0) {
        mahalanobis = mahalanobis_full[i];
      } else if (cache_mode == 1) {
        uncompact_cov(mahalanobis_compact[i], mahalanobis);
      } else {
        const Mat4& cov_B = is_vgicp ? corr_voxels[i]->cov : target->covs[corr_indices[i]];
        fused_mahalanobis(linearization_point, cov_B, source->covs[i], mahalanobis);
      }

      // error = residual^T * mahalanobis * residual
      double Mr[4];
      for (int r = 0; r < 4; r++) {
        double s = mahalanobis(r, 0) * residual[0];
        for (int k = 1; k < 4; k++) s += mahalanobis(r, k) * residual[k];
        Mr[r] = s;
      }
      double error = residual[0] * Mr[0];
      for (int k = 1; k < 4; k++) error += residual[k] * Mr[k];

      if (!H_target) return error;

      // J_target = [-Hat(q) | I ; 0], J_source = [R Hat(p) | -R ; 0]   (4x6)
      Mat46 J_target, J_source;
      std::memset(J_target.m, 0, sizeof(J_target.m));
      std::memset(J_source.m, 0, sizeof(J_source.m));
      double hq[3][3], hp[3][3];
      hat3(transed_mean_A.v, hq);
      hat3(mean_A.v, hp);
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) {
          J_target(r, c) = -hq[r][c];
          double s = delta(r, 0) * hp[0][c];
          for (int k = 1; k < 3; k++) s += delta(r, k) * hp[k][c];
          J_source(r, c) = s;
          J_source(r, 3 + c) = -delta(r, c);
        }
        J_target(r, 3 + r) = 1.0;
      }
      if (normal_B) {
        for (int r = 0; r < 4; r++)
          for (int c = 0; c < 6; c++) {
            J_target(r, c) = (*normal_B)[r] * J_target(r, c);
            J_source(r, c) = (*normal_B)[r] * J_source(r, c);
          }
      }

      // J^T * mahalanobis (6x4)
      Mat64 JtM, JsM;
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 4; c++) {
          double st = J_target(0, r) * mahalanobis(0, c);
          double ss = J_source(0, r) * mahalanobis(0, c);
          for (int k = 1; k < 4; k++) {
            st += J_target(k, r) * mahalanobis(k, c);
            ss += J_source(k, r) * mahalanobis(k, c);
          }
          JtM(r, c) = st;
          JsM(r, c) = ss;
        }
      }

      for (int c = 0; c < 6; c++) {
        for (int r = 0; r < 6; r++) {
          double tt = JtM(r, 0) * J_target(0, c);
          double ss = JsM(r, 0) * J_source(0, c);
          double ts = JtM(r, 0) * J_source(0, c);
          for (int k = 1; k < 4; k++) {
            tt += JtM(r, k) * J_target(k, c);
            ss += JsM(r, k) * J_source(k, c);
            ts += JtM(r, k) * J_source(k, c);
          }
          (*H_target)(r, c) += tt;
          (*H_source)(r, c) += ss;
          (*H_target_source)(r, c) += ts;
        }
      }
      for (int r = 0; r < 6; r++) {
        double bt = JtM(r, 0) * residual[0];
        double bs = JsM(r, 0) * residual[0];
        for (int k = 1; k < 4; k++) {
          bt += JtM(r, k) * residual[k];
          bs += JsM(r, k) * residual[k];
        }
        b_target->v[r] += bt;
        b_source->v[r] += bs;
      }
      return error;
    };

    return scan_matching_reduce_omp(perpoint_task, static_cast<int>(source->size()), num_threads, H_target, H_source, H_target_source, b_target, b_source);
  }
};

// ---------------------------------------------------------------------------------------------
// C interface
// ---------------------------------------------------------------------------------------------
namespace {
Mat4 from_rm16(const double* rm) {
  Mat4 m;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = rm[r * 4 + c];
  return m;
}
}  // namespace

extern "C" {

orc_cloud* orc_cloud_create(const double* xyz, const double* cov3x3, size_t n) {
  auto* c = new orc_cloud;
  c->points.resize(n);
  for (size_t i = 0; i < n; i++) c->points[i] = Vec4{{xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], 1.0}};
  if (cov3x3) {
    c->covs_given = true;
    c->covs.resize(n);
    for (size_t i = 0; i < n; i++) {
      c->covs[i].set_zero();
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) c->covs[i](r, k) = cov3x3[i * 9 + r * 3 + k];
    }
  }
  return c;
}
void orc_cloud_destroy(orc_cloud* c) { delete c; }
size_t orc_cloud_size(const orc_cloud* c) { return c->size(); }

orc_voxelmap* orc_voxelmap_create(double resolution) { return new orc_voxelmap(resolution); }
void orc_voxelmap_destroy(orc_voxelmap* v) { delete v; }
void orc_voxelmap_set_lru(orc_voxelmap* v, int lru_horizon, int lru_clear_cycle) {
  v->lru_horizon = lru_horizon;
  v->lru_clear_cycle = lru_clear_cycle;
}
void orc_voxelmap_insert(orc_voxelmap* v, const orc_cloud* c) {
  if (!c->has_covs()) {
    std::fprintf(stderr, "orc_voxelmap_insert: cloud has no covs\n");
    std::abort();
  }
  v->insert(*c);
}
size_t orc_voxelmap_num_voxels(const orc_voxelmap* v) { return v->flat_voxels.size(); }
void orc_voxelmap_export(const orc_voxelmap* v, int32_t* coords, double* means, double* covs, int32_t* num_points) {
  for (size_t i = 0; i < v->flat_voxels.size(); i++) {
    const auto& e = *v->flat_voxels[i];
    if (coords) {
      coords[i * 3] = e.first.coord.x;
      coords[i * 3 + 1] = e.first.coord.y;
      coords[i * 3 + 2] = e.first.coord.z;
    }
    if (means)
      for (int k = 0; k < 3; k++) means[i * 3 + k] = e.second.mean[k];
    if (covs)
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) covs[i * 9 + r * 3 + c] = e.second.cov(r, c);
    if (num_points) num_points[i] = static_cast<int32_t>(e.second.num_points);
  }
}
void orc_voxelmap_lookup(const orc_voxelmap* v, const double* xyz, size_t n, int32_t* out_idx) {
  for (size_t i = 0; i < n; i++) {
    const Vec4 p = {{xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], 1.0}};
    out_idx[i] = v->lookup_voxel_index(v->voxel_coord(p));
  }
}

orc_kdtree* orc_kdtree_create(const orc_cloud* target, int build_num_threads) {
  auto* t = new orc_kdtree;
  t->points = target;
  t->build(build_num_threads < 1 ? 1 : build_num_threads);
  return t;
}
void orc_kdtree_destroy(orc_kdtree* t) { delete t; }
size_t orc_kdtree_num_nodes(const orc_kdtree* t) { return t->nodes.size(); }
void orc_kdtree_knn(const orc_kdtree* t, const double* queries, size_t nq, int k, double max_sq_dist, uint64_t* out_idx, double* out_sqd,
                    int32_t* out_found, int num_threads) {
  if (num_threads < 1) num_threads = 1;
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (long i = 0; i < static_cast<long>(nq); i++) {
    static_assert(sizeof(size_t) == sizeof(uint64_t), "size_t must be 64-bit");
    const size_t found = t->knn_search(queries + i * 3, k, reinterpret_cast<size_t*>(out_idx + i * k), out_sqd + i * k, max_sq_dist);
    if (out_found) out_found[i] = static_cast<int32_t>(found);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// estimate_covariances (src/gtsam_points/features/covariance_estimation.cpp:18-77), EIG regularisation:
//   k nearest neighbours of every point (the point itself included), cov = (sum p p^T - mean sum p^T) / k over the 4-vectors
//   (top-left 3x3 used), C_i = V diag(eigen_values) V^-1 with V the eigenvectors of cov in ASCENDING eigenvalue order
//   (SelfAdjointEigenSolver::computeDirect), fewer than k neighbours => identity.
// Eigen is not available: the symmetric 3x3 eigen decomposition is restated with cyclic Jacobi rotations in float64 (the
// eigenvectors of a symmetric matrix are unique up to sign when the eigenvalues are distinct, and V diag V^T does not
// depend on the signs), so the result agrees with Eigen's closed-form solver to rounding wherever the spectrum is not
// degenerate.
// ---------------------------------------------------------------------------------------------------------------
namespace {
void jacobi_eigen3(const double A_in[3][3], double w[3], double V[3][3]) {
  double A[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[i][j] = A_in[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - sn * akq;
          A[k][q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - sn * aqk;
          A[q][k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {  // V <- V J
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - sn * vkq;
          V[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 3; i++) w[i] = A[i][i];
  std::sort(order, order + 3, [&](int a, int b) { return w[a] < w[b]; });
  double ws[3], Vs[3][3];
  for (int j = 0; j < 3; j++) {
    ws[j] = w[order[j]];
    for (int i = 0; i < 3; i++) Vs[i][j] = V[i][order[j]];
  }
  for (int j = 0; j < 3; j++) {
    w[j] = ws[j];
    for (int i = 0; i < 3; i++) V[i][j] = Vs[i][j];
  }
}
}  // namespace

void orc_estimate_covariances(const orc_cloud* cloud, int k_neighbors, const double* eigen_values3, int num_threads, double* out_cov3x3) {
  const size_t n = orc_cloud_size(cloud);
  orc_kdtree* tree = orc_kdtree_create(cloud, num_threads);  // covariance_estimation.cpp:19
  if (num_threads < 1) num_threads = 1;
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (long i = 0; i < static_cast<long>(n); i++) {
    std::vector<size_t> k_indices(k_neighbors);
    std::vector<double> k_sq_dists(k_neighbors);
    const Vec4& pi = cloud->points[i];
    const double q[3] = {pi[0], pi[1], pi[2]};
    const size_t found = tree->knn_search(q, k_neighbors, k_indices.data(), k_sq_dists.data(), std::numeric_limits<double>::max());
    double* out = out_cov3x3 + static_cast<size_t>(i) * 9;
    if (found < static_cast<size_t>(k_neighbors)) {  // :27-31 (warning + identity)
      for (int a = 0; a < 9; a++) out[a] = (a % 4 == 0) ? 1.0 : 0.0;
      continue;
    }
    double sum_p[3] = {0, 0, 0}, sum_c[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (size_t j = 0; j < found; j++) {  // :36-40
      const Vec4& pt = cloud->points[k_indices[j]];
      for (int a = 0; a < 3; a++) {
        sum_p[a] += pt[a];
        for (int b = 0; b < 3; b++) sum_c[a][b] += pt[a] * pt[b];
      }
    }
    double cov[3][3];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) cov[a][b] = (sum_c[a][b] - (sum_p[a] / found) * sum_p[b]) / found;  // :42-43
    double w[3], V[3][3];
    jacobi_eigen3(cov, w, V);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {  // :49-53, V orthonormal => V^-1 = V^T
        double v = 0.0;
        for (int c = 0; c < 3; c++) v += V[a][c] * eigen_values3[c] * V[b][c];
        out[a * 3 + b] = v;
      }
  }
  orc_kdtree_destroy(tree);
}

orc_factor* orc_vgicp_create(const orc_voxelmap* target, const orc_cloud* source) {
  if (!source->has_covs()) {
    std::fprintf(stderr, "error: source don't have covs!!\n");  // vgicp_impl:37-40
    std::abort();
  }
  auto* f = new orc_factor;
  f->is_vgicp = true;
  f->target_voxels = target;
  f->source = source;
  return f;
}
orc_factor* orc_gicp_create(const orc_cloud* target, const orc_kdtree* tree, const orc_cloud* source) {
  if (!source->has_covs() || !target->has_covs()) {
    std::fprintf(stderr, "error: target or source don't have covs!!\n");  // gicp_impl:37-65
    std::abort();
  }
  auto* f = new orc_factor;
  f->is_vgicp = false;
  f->target = target;
  f->target_tree = tree;
  f->source = source;
  return f;
}
// integrated_icp_factor_impl.hpp:20-52: point-to-point (use_point_to_plane = 0) or point-to-plane ICP over the same kd-tree
orc_factor* orc_icp_create(const orc_cloud* target, const orc_kdtree* tree, const orc_cloud* source, int use_point_to_plane) {
  if (use_point_to_plane && !target->has_normals()) {
    std::fprintf(stderr, "error: target frame doesn't have required attributes for icp\n");  // icp_impl:37-40
    std::abort();
  }
  auto* f = new orc_factor;
  f->is_vgicp = false;
  f->icp_mode = use_point_to_plane ? 2 : 1;
  f->target = target;
  f->target_tree = tree;
  f->source = source;
  return f;
}
void orc_cloud_set_normals(orc_cloud* c, const double* nxyz) {
  c->normals.resize(c->size());
  for (size_t i = 0; i < c->size(); i++) c->normals[i] = Vec4{{nxyz[i * 3], nxyz[i * 3 + 1], nxyz[i * 3 + 2], 0.0}};
}
void orc_factor_set_fused_cov_cache_mode(orc_factor* f, int mode) { f->cache_mode = mode; }  // integrated_gicp_factor.hpp:104-105
void orc_factor_set_correspondence_update_tolerance(orc_factor* f, double angle, double trans) {  // integrated_gicp_factor.hpp:106-109
  f->correspondence_update_tolerance_rot = angle;
  f->correspondence_update_tolerance_trans = trans;
}

// src/gtsam_points/types/gaussian_voxelmap_cpu_funcs.cpp:126-143 / :145-173
double orc_overlap(const orc_voxelmap* target, const orc_cloud* source, const double* T_rm16) {
  const Mat4 T = from_rm16(T_rm16);
  int num_overlap = 0;
  for (size_t i = 0; i < source->size(); i++) {
    const Vec4 pt = transform_point(T, source->points[i]);
    if (target->lookup_voxel_index(target->voxel_coord(pt)) >= 0) num_overlap++;
  }
  return static_cast<double>(num_overlap) / source->size();
}
double orc_overlap_multi(const orc_voxelmap* const* targets, int num_targets, const orc_cloud* source, const double* Ts_rm16) {
  int num_overlap = 0;
  for (size_t i = 0; i < source->size(); i++) {
    for (int j = 0; j < num_targets; j++) {
      const Mat4 T = from_rm16(Ts_rm16 + 16 * j);
      const Vec4 pt = transform_point(T, source->points[i]);
      if (targets[j]->lookup_voxel_index(targets[j]->voxel_coord(pt)) >= 0) {
        num_overlap++;
        break;
      }
    }
  }
  return static_cast<double>(num_overlap) / source->size();
}

// GaussianVoxelData (types/gaussian_voxel_data.hpp:11-54): 56-byte packed record {coord int32 x3, num_points int32, mean float x3,
// cov float x6 (00, 01, 02, 11, 12, 22), intensity float}; save_compact / load: src/.../gaussian_voxelmap_cpu.cpp:79-135
namespace {
struct GaussianVoxelData {
  int32_t coord[3];
  int32_t num_points;
  float mean[3];
  float cov[6];
  float intensity;
};
static_assert(sizeof(GaussianVoxelData) == 56, "GaussianVoxelData is 56 bytes in the reference");
}  // namespace
int orc_voxelmap_save_compact(const orc_voxelmap* v, const char* path) {
  std::vector<GaussianVoxelData> serial(v->flat_voxels.size());
  for (size_t i = 0; i < serial.size(); i++) {
    const auto& e = *v->flat_voxels[i];
    GaussianVoxelData& d = serial[i];
    d.coord[0] = e.first.coord.x, d.coord[1] = e.first.coord.y, d.coord[2] = e.first.coord.z;
    d.num_points = static_cast<int32_t>(e.second.num_points);
    for (int k = 0; k < 3; k++) d.mean[k] = static_cast<float>(e.second.mean[k]);
    const Mat4& c = e.second.cov;
    d.cov[0] = static_cast<float>(c(0, 0)), d.cov[1] = static_cast<float>(c(0, 1)), d.cov[2] = static_cast<float>(c(0, 2));
    d.cov[3] = static_cast<float>(c(1, 1)), d.cov[4] = static_cast<float>(c(1, 2)), d.cov[5] = static_cast<float>(c(2, 2));
    d.intensity = static_cast<float>(e.second.intensity);
  }
  std::FILE* fp = std::fopen(path, "wb");
  if (!fp) return -1;
  // operator<< of a double prints with 6 significant digits (%g), like the reference's ofstream
  std::fprintf(fp, "compact 1\nresolution %g\nlru_count %zu\nlru_cycle %zu\nlru_thresh %zu\nvoxel_bytes %zu\nnum_voxels %zu\n", 1.0 / v->inv_leaf_size, v->lru_counter,
               v->lru_clear_cycle, v->lru_horizon, sizeof(GaussianVoxelData), serial.size());
  std::fwrite(serial.data(), sizeof(GaussianVoxelData), serial.size(), fp);
  std::fclose(fp);
  return 0;
}
orc_voxelmap* orc_voxelmap_load(const char* path) {
  std::FILE* fp = std::fopen(path, "rb");
  if (!fp) return nullptr;
  char tok[64];
  int compact = 0;
  double resolution = 1.0;
  size_t lru_count = 0, lru_cycle = 0, lru_thresh = 0, voxel_bytes = 0, num_voxels = 0;
  if (std::fscanf(fp, "%63s %d %63s %lf %63s %zu %63s %zu %63s %zu %63s %zu %63s %zu", tok, &compact, tok, &resolution, tok, &lru_count, tok, &lru_cycle, tok, &lru_thresh, tok,
                  &voxel_bytes, tok, &num_voxels) != 14 ||
      voxel_bytes != sizeof(GaussianVoxelData)) {
    std::fclose(fp);
    return nullptr;
  }
  int ch;
  while ((ch = std::fgetc(fp)) != EOF && ch != '\n') {
  }
  std::vector<GaussianVoxelData> serial(num_voxels);
  const size_t got = std::fread(serial.data(), sizeof(GaussianVoxelData), num_voxels, fp);
  std::fclose(fp);
  if (got != num_voxels) return nullptr;
  auto* v = new orc_voxelmap(resolution);
  v->lru_counter = lru_count, v->lru_clear_cycle = lru_cycle, v->lru_horizon = lru_thresh;
  for (const auto& d : serial) {  // GaussianVoxelData::uncompact (gaussian_voxel_data.hpp:27-46)
    auto e = std::make_shared<std::pair<VoxelInfo, GaussianVoxel>>();
    e->first.lru = 0;
    e->first.coord = Vec3i{d.coord[0], d.coord[1], d.coord[2]};
    GaussianVoxel& g = e->second;
    g.finalized = true;
    g.num_points = static_cast<size_t>(d.num_points);
    g.mean = Vec4{{d.mean[0], d.mean[1], d.mean[2], 1.0}};
    g.cov.set_zero();
    g.cov(0, 0) = d.cov[0];
    g.cov(0, 1) = g.cov(1, 0) = d.cov[1];
    g.cov(0, 2) = g.cov(2, 0) = d.cov[2];
    g.cov(1, 1) = d.cov[3];
    g.cov(1, 2) = g.cov(2, 1) = d.cov[4];
    g.cov(2, 2) = d.cov[5];
    g.intensity = d.intensity;
    v->flat_voxels.emplace_back(e);
    v->voxels[e->first.coord] = v->flat_voxels.size() - 1;
  }
  return v;
}

// merge_frames (src/gtsam_points/types/gaussian_voxelmap_cpu_funcs.cpp:25-113): 21-bit voxel keys relative to the FIRST pose,
// sorted (std::sort is not stable, but every point of a voxel receives the same destination, so the result does not depend on
// the order inside a key), sums taken in WORLD coordinates (poses[i], not relative) in frame / point order, divided by the
// count (the homogeneous w).  Returns the number of merged points; out_* sized by the caller for the total point count.
size_t orc_merge_frames(const double* poses_rm16, const orc_cloud* const* frames, int num_frames, double downsample_resolution, double* out_xyz, double* out_cov3x3) {
  constexpr int coord_bits = 21;
  constexpr uint64_t coord_bitmask = (1ull << coord_bits) - 1;
  const int coord_offset = 1 << (coord_bits - 1);
  const double inv_resolution = 1.0 / downsample_resolution;
  std::vector<std::pair<uint64_t, uint64_t>> coords_indices;
  const Mat4 first_inv = orc_factor::rigid_inverse(from_rm16(poses_rm16));
  for (int fid = 0; fid < num_frames; fid++) {
    const Mat4 pose = mul44(first_inv, from_rm16(poses_rm16 + 16 * fid));
    for (size_t pid = 0; pid < frames[fid]->size(); pid++) {
      const Vec4 point = transform_point(pose, frames[fid]->points[pid]);
      int64_t c[3];
      bool out_of_range = false;
      for (int k = 0; k < 3; k++) {
        c[k] = static_cast<int64_t>(fast_floor1(point[k] * inv_resolution)) + coord_offset;
        // the reference compares a signed Array4i against the unsigned mask: negative values convert to huge unsigned ones
        if (c[k] < 0 || static_cast<uint64_t>(c[k]) > coord_bitmask) out_of_range = true;
      }
      if (out_of_range) continue;
      const uint64_t index = (static_cast<uint64_t>(fid) << 32) | (static_cast<uint64_t>(pid) & 0xffffffffull);
      coords_indices.emplace_back((static_cast<uint64_t>(c[0]) & coord_bitmask) | ((static_cast<uint64_t>(c[1]) & coord_bitmask) << coord_bits) |
                                    ((static_cast<uint64_t>(c[2]) & coord_bitmask) << (2 * coord_bits)),
                                  index);
    }
  }
  std::sort(coords_indices.begin(), coords_indices.end(), [](const auto& l, const auto& r) { return l.first < r.first; });
  std::vector<std::vector<size_t>> dest(num_frames);
  for (int i = 0; i < num_frames; i++) dest[i].assign(frames[i]->size(), 0);
  size_t num_voxels = 0;
  for (size_t i = 0; i < coords_indices.size(); i++) {
    if (i && coords_indices[i - 1].first != coords_indices[i].first) num_voxels++;
    const uint64_t index = coords_indices[i].second;
    dest[(index >> 32) & 0xffffffffull][index & 0xffffffffull] = num_voxels;
  }
  num_voxels++;
  std::vector<Vec4> pts(num_voxels, Vec4{{0, 0, 0, 0}});
  Mat4 zero;
  zero.set_zero();
  std::vector<Mat4> covs(num_voxels, zero);
  for (int i = 0; i < num_frames; i++) {
    const Mat4 pose = from_rm16(poses_rm16 + 16 * i);
    const Mat4 poseT = transpose44(pose);
    for (size_t j = 0; j < frames[i]->size(); j++) {
      const size_t d = dest[i][j];
      const Vec4 tp = transform_point(pose, frames[i]->points[j]);
      for (int k = 0; k < 4; k++) pts[d][k] += tp[k];
      const Mat4 c = mul44(mul44(pose, frames[i]->covs[j]), poseT);
      for (int k = 0; k < 16; k++) covs[d].m[k] += c.m[k];
    }
  }
  for (size_t i = 0; i < num_voxels; i++) {
    const double w = pts[i][3];
    for (int k = 0; k < 16; k++) covs[i].m[k] /= w;
    for (int k = 0; k < 4; k++) pts[i][k] /= w;
    for (int k = 0; k < 3; k++) out_xyz[i * 3 + k] = pts[i][k];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) out_cov3x3[i * 9 + r * 3 + c] = covs[i](r, c);
  }
  return num_voxels;
}

void orc_factor_destroy(orc_factor* f) { delete f; }
void orc_factor_set_num_threads(orc_factor* f, int n) { f->num_threads = n < 1 ? 1 : n; }
void orc_factor_set_max_correspondence_distance(orc_factor* f, double dist) { f->max_correspondence_distance_sq = dist * dist; }

// src/gtsam_points/factors/integrated_matching_cost_factor.cpp:37-55
void orc_factor_linearize(orc_factor* f, const double* delta_rm16, orc_linearized* out) {
  const Mat4 delta = from_rm16(delta_rm16);
  f->update_correspondences(delta);
  Mat66 H_target, H_source, H_target_source;
  Vec6 b_target, b_source;
  const double error = f->evaluate(delta, &H_target, &H_source, &H_target_source, &b_target, &b_source);
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) {
      out->H_target[r * 6 + c] = H_target(r, c);
      out->H_source[r * 6 + c] = H_source(r, c);
      out->H_target_source[r * 6 + c] = H_target_source(r, c);
    }
    out->b_target[r] = b_target.v[r];
    out->b_source[r] = b_source.v[r];
  }
  out->error = error;
  size_t inliers = 0;
  if (f->is_vgicp) {
    for (const auto* v : f->corr_voxels) inliers += (v != nullptr);  // integrated_vgicp_factor.hpp:78-85
  } else {
    for (long c : f->corr_indices) inliers += (c >= 0);  // integrated_gicp_factor.hpp:112-116
  }
  out->num_inliers = static_cast<double>(inliers);
}

// src/gtsam_points/factors/integrated_matching_cost_factor.cpp:32-35
double orc_factor_error(orc_factor* f, const double* delta_rm16) {
  const Mat4 delta = from_rm16(delta_rm16);
  return f->evaluate(delta, nullptr, nullptr, nullptr, nullptr, nullptr);
}

void orc_factor_correspondences(const orc_factor* f, int64_t* out) {
  const size_t n = f->source->size();
  if (f->is_vgicp) {
    // pointer -> flat index (the reference stores `const GaussianVoxel*`); recover ids through the map
    std::unordered_map<const GaussianVoxel*, int64_t> ids;
    ids.reserve(f->target_voxels->flat_voxels.size());
    for (size_t i = 0; i < f->target_voxels->flat_voxels.size(); i++) ids[&f->target_voxels->flat_voxels[i]->second] = static_cast<int64_t>(i);
    for (size_t i = 0; i < n; i++) out[i] = (i < f->corr_voxels.size() && f->corr_voxels[i]) ? ids[f->corr_voxels[i]] : -1;
  } else {
    for (size_t i = 0; i < n; i++) out[i] = i < f->corr_indices.size() ? f->corr_indices[i] : -1;
  }
}

// src/gtsam_points/factors/integrated_matching_cost_factor.cpp:57-69 : Pose3::inverse() * Pose3
// gtsam::Pose3::inverse() = (R^T, -R^T t); compose = (R1 R2, t1 + R1 t2).
void orc_calc_delta(const double* Tt, const double* Ts, double* d) {
  double Rt[3][3], tt[3], Rs[3][3], ts[3];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) {
      Rt[r][c] = Tt[r * 4 + c];
      Rs[r][c] = Ts[r * 4 + c];
    }
    tt[r] = Tt[r * 4 + 3];
    ts[r] = Ts[r * 4 + 3];
  }
  double Ri[3][3], ti[3];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Ri[r][c] = Rt[c][r];
  }
  for (int r = 0; r < 3; r++) ti[r] = -((Ri[r][0] * tt[0] + Ri[r][1] * tt[1]) + Ri[r][2] * tt[2]);
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) d[r * 4 + c] = (Ri[r][0] * Rs[0][c] + Ri[r][1] * Rs[1][c]) + Ri[r][2] * Rs[2][c];
    d[r * 4 + 3] = ti[r] + ((Ri[r][0] * ts[0] + Ri[r][1] * ts[1]) + Ri[r][2] * ts[2]);
  }
  d[12] = d[13] = d[14] = 0.0;
  d[15] = 1.0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
