// gtsam_points_b200.hpp -- header-only C++ adapters that keep gtsam_points' C++ surface for the scan-matching path on
// top of the C ABI (include/b2points.h, libb2points.so).
//
//   gtsam_points_b200::PointCloudGPU            <- gtsam_points::PointCloudGPU / PointCloud     (types/point_cloud.hpp:19-119)
//   gtsam_points_b200::GaussianVoxelMap[GPU]    <- gtsam_points::GaussianVoxelMap[GPU]          (types/gaussian_voxelmap.hpp:16-50, gaussian_voxelmap_gpu.hpp:39-108)
//   gtsam_points_b200::NearestNeighborSearch,
//                      KdTreeGPU                <- gtsam_points::NearestNeighborSearch / KdTree (ann/nearest_neighbor_search.hpp:16-57)
//   gtsam_points_b200::NonlinearFactorGPU       <- gtsam_points::NonlinearFactorGPU, the 10-call asynchronous protocol (factors/nonlinear_factor_gpu.hpp:37-122)
//   gtsam_points_b200::IntegratedVGICPFactor    <- IntegratedVGICPFactor_ / IntegratedVGICPFactorGPU (factors/integrated_vgicp_factor.hpp:25-113, _gpu.hpp:29-155)
//   gtsam_points_b200::IntegratedGICPFactor     <- IntegratedGICPFactor_                         (factors/integrated_gicp_factor.hpp:32-152)
//   gtsam_points_b200::NonlinearFactorSetGPU    <- gtsam_points::NonlinearFactorSetGPU : NonlinearFactorSet (optimizers/linearization_hook.hpp:11-29)
//
// With GTSAM on the include path (B2_HAVE_GTSAM) the factors derive from gtsam::NonlinearFactor: linearize() returns a
// gtsam::HessianFactor, clone() a NonlinearFactor::shared_ptr, so they drop into LevenbergMarquardtOptimizerExt / ISAM2Ext
// unchanged; with gtsam_points' linearization_hook.hpp also there, NonlinearFactorSetGPU derives from
// gtsam_points::NonlinearFactorSet and create_nonlinear_factor_set_gpu() can be handed to LinearizationHook::register_hook
// (INTEGRATION.md).  Smart pointers follow GTSAM's own (std:: in 4.3, boost:: in 4.2): everything goes through
// FactorBase::shared_ptr and an ADL dynamic_pointer_cast.  This branch is compiled in the CPU test suite against header mocks
// (tests/cpp/mock_gtsam) and run on the GPU box through the same mocks (tests/test_cpp_adapters.py).
// Without GTSAM the same classes are built on a small Pose / Values / HessianFactor stand-in.
//
// Frames: constructors take the reference's own frame shape -- any type with `points` (Vector4d*), `covs` (Matrix4d*) and
// size(), e.g. gtsam_points::PointCloud (types/point_cloud.hpp:103-118) -- through std::shared_ptr<const Frame>; the device
// copy is made once per frame and shared between factors (PointCloudGPU::from_frame).
//
// Error behaviour mirrors the reference: constructor preconditions print the reference's message and abort()
// (factors/impl/integrated_vgicp_factor_impl.hpp:32-45); runtime CUDA failures throw std::runtime_error with b2_last_error().
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "b2points.h"

#if defined(__has_include)
#if __has_include(<gtsam/nonlinear/NonlinearFactor.h>) && !defined(B2_NO_GTSAM)
#define B2_HAVE_GTSAM 1
#include <gtsam/geometry/Pose3.h>
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/nonlinear/NonlinearFactor.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
#if __has_include(<gtsam_points/optimizers/linearization_hook.hpp>)
#define B2_HAVE_LINEARIZATION_HOOK 1
#include <gtsam_points/optimizers/linearization_hook.hpp>
#endif
#endif
#endif

namespace gtsam_points_b200 {

inline void check(b2_status st, const char* what) {
  if (st != B2_OK) throw std::runtime_error(std::string(what) + ": " + b2_last_error());
}

// ---------------------------------------------------------------------------------------------------------------------
// Pose / Values: GTSAM's when available, otherwise a minimal stand-in (row-major 4x4)
// ---------------------------------------------------------------------------------------------------------------------
using Mat4 = std::array<double, 16>;  // row-major

#ifdef B2_HAVE_GTSAM
using Key = gtsam::Key;
using Values = gtsam::Values;
inline Mat4 pose_matrix(const gtsam::Pose3& pose) {
  const gtsam::Matrix4 m = pose.matrix();
  Mat4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r[i * 4 + j] = m(i, j);
  return r;
}
inline Mat4 pose_matrix(const Values& values, Key key) { return pose_matrix(values.at<gtsam::Pose3>(key)); }
#else
using Key = std::uint64_t;
struct Values {
  std::map<Key, Mat4> poses;
  void insert(Key k, const Mat4& T) { poses[k] = T; }
  const Mat4& at(Key k) const { return poses.at(k); }
};
inline Mat4 pose_matrix(const Mat4& pose) { return pose; }
inline Mat4 pose_matrix(const Values& values, Key key) { return values.at(key); }
#endif

// delta = T_target^-1 * T_source with gtsam::Pose3 semantics (inverse = (R^T, -R^T t)); integrated_matching_cost_factor.cpp:57-69
inline Mat4 calc_delta(const Mat4& Tt, const Mat4& Ts) {
  Mat4 d{};
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) d[r * 4 + c] = Tt[0 * 4 + r] * Ts[0 * 4 + c] + Tt[1 * 4 + r] * Ts[1 * 4 + c] + Tt[2 * 4 + r] * Ts[2 * 4 + c];
    double ti = 0.0, ts = 0.0;
    for (int k = 0; k < 3; k++) {
      ti -= Tt[k * 4 + r] * Tt[k * 4 + 3];
      ts += Tt[k * 4 + r] * Ts[k * 4 + 3];
    }
    d[r * 4 + 3] = ti + ts;
  }
  d[15] = 1.0;
  return d;
}

// ---------------------------------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------------------------------
class PointCloudGPU;
class Context {
public:
  using Ptr = std::shared_ptr<Context>;
  explicit Context(int device = 0, void* stream = nullptr) { check(b2_ctx_create(device, stream, &ctx_), "b2_ctx_create"); }
  ~Context() { b2_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  b2_ctx* get() const { return ctx_; }
  static Ptr default_context() {
    static Ptr ctx = std::make_shared<Context>(0);
    return ctx;
  }

private:
  friend class PointCloudGPU;
  b2_ctx* ctx_ = nullptr;
  // device copies of host frames, keyed by (frame address, points address, size): one upload per frame however many factors use it
  std::mutex frame_mutex_;
  std::map<std::tuple<const void*, const void*, std::size_t>, std::weak_ptr<const PointCloudGPU>> frame_cache_;
};

// ---------------------------------------------------------------------------------------------------------------------
// Types
// ---------------------------------------------------------------------------------------------------------------------
class PointCloudGPU {
public:
  using Ptr = std::shared_ptr<PointCloudGPU>;
  using ConstPtr = std::shared_ptr<const PointCloudGPU>;

  // points: the reference's Vector4d array (x,y,z,1), covs: its Matrix4d array (may be null) -- types/point_cloud.hpp:106-108
  PointCloudGPU(const double* points_xyz1, const double* covs_4x4, std::size_t n, Context::Ptr ctx = Context::default_context(), unsigned flags = B2_CLOUD_DEFAULT)
  : ctx_(ctx), num_points_(n), has_covs_(covs_4x4 != nullptr) {
    check(b2_cloud_create(ctx_->get(), points_xyz1, 4, covs_4x4, 16, n, flags, &cloud_), "b2_cloud_create");
  }
  // packed variants (xyz / 3x3)
  static Ptr from_packed(const double* xyz, const double* cov3x3, std::size_t n, Context::Ptr ctx = Context::default_context(), unsigned flags = B2_CLOUD_DEFAULT) {
    Ptr p(new PointCloudGPU());
    p->ctx_ = ctx;
    p->num_points_ = n;
    p->has_covs_ = cov3x3 != nullptr;
    check(b2_cloud_create(ctx->get(), xyz, 3, cov3x3, 9, n, flags, &p->cloud_), "b2_cloud_create");
    return p;
  }
  // Any frame with the reference's PointCloud shape: `points` -> array of 4 doubles per point, `covs` -> array of 16 doubles
  // per point or null, size().  (Eigen::Vector4d / Matrix4d are exactly that.)  Uploaded once per frame and context.
  template <typename Frame>
  static ConstPtr from_frame(const std::shared_ptr<const Frame>& frame, Context::Ptr ctx = Context::default_context()) {
    if (!frame || !frame->points) return nullptr;
    const auto key = std::make_tuple(static_cast<const void*>(frame.get()), static_cast<const void*>(frame->points), static_cast<std::size_t>(frame->size()));
    std::lock_guard<std::mutex> lock(ctx->frame_mutex_);
    auto it = ctx->frame_cache_.find(key);
    if (it != ctx->frame_cache_.end()) {
      if (auto alive = it->second.lock()) return alive;
    }
    auto cloud = std::make_shared<PointCloudGPU>(reinterpret_cast<const double*>(frame->points), reinterpret_cast<const double*>(frame->covs), frame->size(), ctx);
    ctx->frame_cache_[key] = cloud;
    return cloud;
  }
  ~PointCloudGPU() { b2_cloud_destroy(cloud_); }
  PointCloudGPU(const PointCloudGPU&) = delete;

  std::size_t size() const { return num_points_; }
  bool has_points() const { return true; }
  bool has_covs() const { return has_covs_; }
  b2_cloud* handle() const { return cloud_; }
  Context::Ptr context() const { return ctx_; }

private:
  PointCloudGPU() = default;
  Context::Ptr ctx_;
  b2_cloud* cloud_ = nullptr;
  std::size_t num_points_ = 0;
  bool has_covs_ = false;
};

// types/gaussian_voxelmap.hpp:16-50: the abstract map the factor constructors take
class GaussianVoxelMap {
public:
  using Ptr = std::shared_ptr<GaussianVoxelMap>;
  using ConstPtr = std::shared_ptr<const GaussianVoxelMap>;
  virtual ~GaussianVoxelMap() {}
  virtual double voxel_resolution() const = 0;
  virtual void save_compact(const std::string& path) const = 0;
};

class GaussianVoxelMapGPU : public GaussianVoxelMap {
public:
  using Ptr = std::shared_ptr<GaussianVoxelMapGPU>;
  using ConstPtr = std::shared_ptr<const GaussianVoxelMapGPU>;

  explicit GaussianVoxelMapGPU(double resolution, Context::Ptr ctx = Context::default_context()) : ctx_(ctx), resolution_(resolution) {}
  ~GaussianVoxelMapGPU() override { b2_voxelmap_destroy(vm_); }
  GaussianVoxelMapGPU(const GaussianVoxelMapGPU&) = delete;

  double voxel_resolution() const override { return resolution_; }

  // GaussianVoxelMap::insert(const PointCloud&).  Unlike the reference's GPU map (one-shot, types/gaussian_voxelmap_gpu.hpp:63)
  // this is the CPU map's INCREMENTAL insert (ann/impl/incremental_voxelmap_impl.hpp:31-68): repeated calls extend the map,
  // ids stay first-touch, voxels untouched for lru_horizon inserts are dropped every lru_clear_cycle inserts.
  void insert(const double* points, int point_stride, const double* covs, int cov_stride, std::size_t n) {
    if (!vm_) check(b2_voxelmap_create(ctx_->get(), resolution_, &vm_), "b2_voxelmap_create");
    check(b2_voxelmap_insert(vm_, points, point_stride, covs, cov_stride, n), "b2_voxelmap_insert");
  }
  // IncrementalVoxelMap::set_lru_horizon / set_lru_clear_cycle (ann/incremental_voxelmap.hpp:41-46); defaults 10 / 10
  void set_lru(std::size_t lru_horizon, std::size_t lru_clear_cycle) {
    if (!vm_) check(b2_voxelmap_create(ctx_->get(), resolution_, &vm_), "b2_voxelmap_create");
    check(b2_voxelmap_set_lru(vm_, lru_horizon, lru_clear_cycle), "b2_voxelmap_set_lru");
  }
  template <typename Frame>
  void insert(const Frame& frame) {  // frame: the reference's PointCloud shape
    if (!frame.points || !frame.covs) {
      std::cerr << "error: points/covs have not been allocated!!" << std::endl;  // gaussian_voxelmap_gpu.cu:218-222
      abort();
    }
    insert(reinterpret_cast<const double*>(frame.points), 4, reinterpret_cast<const double*>(frame.covs), 16, frame.size());
  }
  void save_compact(const std::string& path) const override {
    if (!vm_) throw std::runtime_error("GaussianVoxelMapGPU::save_compact: empty map");
    check(b2_voxelmap_save_compact(vm_, path.c_str()), "b2_voxelmap_save_compact");
  }
  static Ptr load(const std::string& path, Context::Ptr ctx = Context::default_context()) {
    b2_voxelmap* vm = nullptr;
    if (b2_voxelmap_load(ctx->get(), path.c_str(), &vm) != B2_OK) return nullptr;
    b2_voxelmap_info info{};
    check(b2_voxelmap_get_info(vm, &info), "b2_voxelmap_get_info");
    Ptr p = std::make_shared<GaussianVoxelMapGPU>(info.resolution, ctx);
    p->vm_ = vm;
    return p;
  }
  std::size_t num_voxels() const {
    b2_voxelmap_info info{};
    if (vm_) check(b2_voxelmap_get_info(vm_, &info), "b2_voxelmap_get_info");
    return info.num_voxels;
  }
  b2_voxelmap* handle() const { return vm_; }
  Context::Ptr context() const { return ctx_; }

private:
  Context::Ptr ctx_;
  double resolution_;
  b2_voxelmap* vm_ = nullptr;
};

// overlap_gpu (types/gaussian_voxelmap_gpu.hpp:116-125): fraction of source points that fall into a voxel of the target
inline double overlap_gpu(const GaussianVoxelMap::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const Mat4& T_target_source) {
  auto t = std::dynamic_pointer_cast<const GaussianVoxelMapGPU>(target);
  if (!t || !t->handle() || !source) {
    std::cerr << "error: Failed to cast target voxelmap to GaussianVoxelMapGPU!!" << std::endl;  // gaussian_voxelmap_gpu_funcs.cu:70-75
    abort();
  }
  double v = 0.0;
  const b2_voxelmap* maps[1] = {t->handle()};
  check(b2_overlap(maps, 1, source->handle(), T_target_source.data(), &v), "b2_overlap");
  return v;
}
inline double overlap_gpu(const std::vector<GaussianVoxelMap::ConstPtr>& targets, const PointCloudGPU::ConstPtr& source, const std::vector<Mat4>& Ts_target_source) {
  std::vector<const b2_voxelmap*> maps;
  std::vector<double> Ts;
  for (std::size_t i = 0; i < targets.size(); i++) {
    auto t = std::dynamic_pointer_cast<const GaussianVoxelMapGPU>(targets[i]);
    if (!t || !t->handle()) {
      std::cerr << "error: Failed to cast target voxelmap to GaussianVoxelMapGPU!!" << std::endl;
      abort();
    }
    maps.push_back(t->handle());
    Ts.insert(Ts.end(), Ts_target_source[i].begin(), Ts_target_source[i].end());
  }
  double v = 0.0;
  check(b2_overlap(maps.data(), maps.size(), source->handle(), Ts.data(), &v), "b2_overlap");
  return v;
}

// ann/nearest_neighbor_search.hpp:16-57
class NearestNeighborSearch {
public:
  virtual ~NearestNeighborSearch() {}
  virtual std::size_t knn_search(const double* pt, std::size_t k, std::size_t* k_indices, double* k_sq_dists, double max_sq_dist = std::numeric_limits<double>::max()) const { return 0; }
  virtual std::size_t radius_search(const double* pt, double radius, std::vector<std::size_t>& indices, std::vector<double>& sq_dists, int max_num_neighbors = std::numeric_limits<int>::max()) const {
    return 0;
  }
};

// exact k-nearest-neighbour search on the device; batch entry points preferred (one launch for all queries)
class KdTreeGPU : public NearestNeighborSearch {
public:
  using Ptr = std::shared_ptr<KdTreeGPU>;
  using ConstPtr = std::shared_ptr<const KdTreeGPU>;
  KdTreeGPU(const double* points, int point_stride, std::size_t n, Context::Ptr ctx = Context::default_context()) : ctx_(ctx) {
    check(b2_kdtree_create(ctx_->get(), points, point_stride, n, &tree_), "b2_kdtree_create");
  }
  ~KdTreeGPU() override { b2_kdtree_destroy(tree_); }
  KdTreeGPU(const KdTreeGPU&) = delete;

  // NearestNeighborSearch::knn_search (ann/nearest_neighbor_search.hpp:31-35): indices / distances sorted by distance,
  // slots beyond the number found hold (size_t)-1 / max_sq_dist (ann/knn_result.hpp:44-72)
  std::size_t knn_search(const double* pt, std::size_t k, std::size_t* k_indices, double* k_sq_dists, double max_sq_dist = std::numeric_limits<double>::max()) const override {
    std::vector<std::int64_t> idx(k);
    check(b2_kdtree_knn(tree_, pt, 3, 1, static_cast<int>(k), max_sq_dist, idx.data(), k_sq_dists), "b2_kdtree_knn");
    std::size_t found = 0;
    for (std::size_t i = 0; i < k; i++) {
      k_indices[i] = idx[i] < 0 ? static_cast<std::size_t>(-1) : static_cast<std::size_t>(idx[i]);
      found += idx[i] >= 0;
    }
    return found;
  }
  // ann/nearest_neighbor_search.hpp:45-56: all points within `radius`, at most max_num_neighbors, sorted by distance
  std::size_t radius_search(const double* pt, double radius, std::vector<std::size_t>& indices, std::vector<double>& sq_dists, int max_num_neighbors = std::numeric_limits<int>::max()) const override {
    int k = 16;
    while (true) {
      const int kk = std::min(k, max_num_neighbors);
      std::vector<std::int64_t> idx(kk);
      std::vector<double> sq(kk);
      check(b2_kdtree_knn(tree_, pt, 3, 1, kk, radius * radius, idx.data(), sq.data()), "b2_kdtree_knn");
      int found = 0;
      while (found < kk && idx[found] >= 0) found++;
      if (found < kk || kk == max_num_neighbors || kk >= B2_KNN_MAX_K) {  // every neighbour inside the radius has been seen
        indices.assign(idx.begin(), idx.begin() + found);
        sq_dists.assign(sq.begin(), sq.begin() + found);
        return static_cast<std::size_t>(found);
      }
      k *= 2;
    }
  }
  void knn_search_batch(const double* queries, int stride, std::size_t n, double max_sq_dist, std::int64_t* indices, double* sq_dists, int k = 1) const {
    check(b2_kdtree_knn(tree_, queries, stride, n, k, max_sq_dist, indices, sq_dists), "b2_kdtree_knn");
  }
  b2_kdtree* handle() const { return tree_; }

private:
  Context::Ptr ctx_;
  b2_kdtree* tree_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------
// Factors
// ---------------------------------------------------------------------------------------------------------------------
#ifdef B2_HAVE_GTSAM
using FactorBase = gtsam::NonlinearFactor;
using LinearFactorPtr = gtsam::GaussianFactor::shared_ptr;
using FactorBasePtr = gtsam::NonlinearFactor::shared_ptr;
using KeyFormatter = gtsam::KeyFormatter;
inline std::string default_key_format(Key k) { return gtsam::DefaultKeyFormatter(k); }
#else
struct HessianFactor {  // stand-in for gtsam::HessianFactor(k_t, k_s, G11, G12, g1, G22, g2, f) / (k_s, G22, g2, f)
  std::vector<Key> keys;
  double G11[36], G12[36], G22[36], g1[6], g2[6], f;
};
using LinearFactorPtr = std::shared_ptr<HessianFactor>;
using KeyFormatter = std::function<std::string(Key)>;
inline std::string default_key_format(Key k) { return std::to_string(k); }
struct FactorBase {
  using shared_ptr = std::shared_ptr<FactorBase>;
  FactorBase() {}
  template <typename CONTAINER>
  explicit FactorBase(const CONTAINER& keys) : keys_(keys.begin(), keys.end()) {}
  virtual ~FactorBase() {}
  virtual void print(const std::string& = "", const KeyFormatter& = &default_key_format) const {}
  virtual double error(const Values& c) const = 0;
  virtual std::size_t dim() const = 0;
  virtual LinearFactorPtr linearize(const Values& c) const = 0;
  virtual shared_ptr clone() const = 0;
  const std::vector<Key>& keys() const { return keys_; }
  std::vector<Key> keys_;
};
using FactorBasePtr = FactorBase::shared_ptr;
#endif

// dynamic_pointer_cast that follows GTSAM's smart-pointer flavour (std:: or, in GTSAM 4.2, boost:: found through ADL)
template <typename T, typename P>
inline auto factor_pointer_cast(const P& p) {
  using std::dynamic_pointer_cast;
  return dynamic_pointer_cast<T>(p);
}

// factors/nonlinear_factor_gpu.hpp:37-122 -- the asynchronous device protocol.  NonlinearFactorSetGPU (the reference's, or the
// batched one below) drives it: set_*_point fills a host blob, issue_* enqueues device work that leaves its result in a
// device blob, sync() waits, store_* reads the downloaded blob.  Blob contents are private to the factor.
class NonlinearFactorGPU : public FactorBase {
public:
  template <typename CONTAINER>
  explicit NonlinearFactorGPU(const CONTAINER& keys) : FactorBase(keys) {}
  ~NonlinearFactorGPU() override {}
  virtual std::size_t linearization_input_size() const = 0;
  virtual std::size_t linearization_output_size() const = 0;
  virtual std::size_t evaluation_input_size() const = 0;
  virtual std::size_t evaluation_output_size() const = 0;
  virtual void set_linearization_point(const Values& values, void* lin_input_cpu) = 0;
  virtual void issue_linearize(const void* lin_input_cpu, const void* lin_input_gpu, void* lin_output_gpu) = 0;
  virtual void store_linearized(const void* lin_output_cpu) = 0;
  virtual void set_evaluation_point(const Values& values, void* eval_input_cpu) = 0;
  virtual void issue_compute_error(const void* lin_input_cpu, const void* eval_input_cpu, const void* lin_input_gpu, const void* eval_input_gpu, void* eval_output_gpu) = 0;
  virtual void store_computed_error(const void* eval_output_cpu) = 0;
  virtual void sync() = 0;
};

class NonlinearFactorSetGPU;

// factors/integrated_gicp_factor.hpp:20-24.  The device path recomputes M = (C_B + R C_A R^T)^-1 from the rotation of the
// linearization point inside the kernel (the reference's NONE behaviour: no cache to keep in HBM).  FULL gives the same
// numbers; COMPACT in the reference additionally rounds the cached matrix to float32 (its own test accepts 1e-3 between the
// modes, src/test/test_compact_mahalanobis.cpp:146-158), which is not reproduced.
enum class FusedCovCacheMode { FULL, COMPACT, NONE };

class IntegratedMatchingCostFactorB200 : public NonlinearFactorGPU {
public:
  ~IntegratedMatchingCostFactorB200() override { b2_factor_destroy(factor_); }
  std::size_t dim() const override { return 6; }

  Mat4 calc_delta(const Values& values) const {
    if (is_binary_) return gtsam_points_b200::calc_delta(pose_matrix(values, this->keys()[0]), pose_matrix(values, this->keys()[1]));
    return gtsam_points_b200::calc_delta(fixed_target_pose_, pose_matrix(values, this->keys()[0]));
  }

  // NonlinearFactor::error -- re-uses the correspondences / fused covariances frozen at the last linearize()
  double error(const Values& values) const override {
    if (has_evaluation_) {  // filled by NonlinearFactorSetGPU::error, like IntegratedVGICPFactorGPU::evaluation_result
      has_evaluation_ = false;
      return evaluation_;
    }
    const Mat4 d = calc_delta(values);
    double e = 0.0;
    check(b2_factor_error(factor_, d.data(), &e), "b2_factor_error");
    return e;
  }

  LinearFactorPtr linearize(const Values& values) const override {
    if (!has_linearization_) {  // not batched through a NonlinearFactorSetGPU: evaluate now
      const Mat4 d = calc_delta(values);
      check(b2_factor_linearize(factor_, d.data(), &linearized_), "b2_factor_linearize");
    }
    has_linearization_ = false;
    return make_hessian(linearized_);
  }

  // ---- NonlinearFactorGPU protocol: input blob = delta (16 doubles), linearization output blob = b2_linearized (1 KiB),
  //      evaluation output blob = one double.  The pose travels with the launch, so the *_input_gpu blobs are not read. ----
  std::size_t linearization_input_size() const override { return 16 * sizeof(double); }
  std::size_t linearization_output_size() const override { return sizeof(b2_linearized); }
  std::size_t evaluation_input_size() const override { return 16 * sizeof(double); }
  std::size_t evaluation_output_size() const override { return sizeof(double); }
  void set_linearization_point(const Values& values, void* lin_input_cpu) override {
    const Mat4 d = calc_delta(values);
    std::memcpy(lin_input_cpu, d.data(), sizeof(double) * 16);
  }
  void issue_linearize(const void* lin_input_cpu, const void* /*lin_input_gpu*/, void* lin_output_gpu) override {
    check(b2_factor_issue_linearize(factor_, static_cast<const double*>(lin_input_cpu), static_cast<double*>(lin_output_gpu)), "b2_factor_issue_linearize");
  }
  void store_linearized(const void* lin_output_cpu) override {
    std::memcpy(&linearized_, lin_output_cpu, sizeof(b2_linearized));
    has_linearization_ = true;
    evaluation_ = linearized_.error;  // integrated_vgicp_factor_gpu.cpp:239-245: the linearization also primes the evaluation result
    has_evaluation_ = true;
  }
  void set_evaluation_point(const Values& values, void* eval_input_cpu) override {
    const Mat4 d = calc_delta(values);
    std::memcpy(eval_input_cpu, d.data(), sizeof(double) * 16);
  }
  void issue_compute_error(const void* /*lin_input_cpu*/, const void* eval_input_cpu, const void* /*lin_input_gpu*/, const void* /*eval_input_gpu*/, void* eval_output_gpu) override {
    check(b2_factor_issue_error(factor_, static_cast<const double*>(eval_input_cpu), static_cast<double*>(eval_output_gpu)), "b2_factor_issue_error");
  }
  void store_computed_error(const void* eval_output_cpu) override {
    std::memcpy(&evaluation_, eval_output_cpu, sizeof(double));
    has_evaluation_ = true;
  }
  void sync() override { check(b2_factor_sync(factor_), "b2_factor_sync"); }

  int num_inliers() const { return static_cast<int>(linearized_.num_inliers); }
  double inlier_fraction() const { return linearized_.num_inliers / static_cast<double>(b2_factor_num_points(factor_)); }
  const b2_linearized& last_linearized() const { return linearized_; }
  std::vector<std::int64_t> correspondences() const {
    std::vector<std::int64_t> c(b2_factor_num_points(factor_));
    check(b2_factor_correspondences(factor_, c.data()), "b2_factor_correspondences");
    return c;
  }
  b2_factor* handle() const { return factor_; }

protected:
  IntegratedMatchingCostFactorB200(Key target_key, Key source_key) : NonlinearFactorGPU(std::vector<Key>{target_key, source_key}), is_binary_(true) {
    fixed_target_pose_ = Mat4{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  }
  IntegratedMatchingCostFactorB200(const Mat4& fixed_target_pose, Key source_key) : NonlinearFactorGPU(std::vector<Key>{source_key}), is_binary_(false), fixed_target_pose_(fixed_target_pose) {}

  void print_keys(const std::string& s, const char* name, const KeyFormatter& keyFormatter) const {
    std::cout << s << name;
    if (is_binary_)
      std::cout << "(" << keyFormatter(this->keys()[0]) << ", " << keyFormatter(this->keys()[1]) << ")" << std::endl;
    else
      std::cout << "(fixed, " << keyFormatter(this->keys()[0]) << ")" << std::endl;
  }

  LinearFactorPtr make_hessian(const b2_linearized& l) const {
#ifdef B2_HAVE_GTSAM
    auto M = [](const double* p) {
      gtsam::Matrix m(6, 6);
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) m(i, j) = p[i * 6 + j];
      return m;
    };
    auto V = [](const double* p, double s) {
      gtsam::Vector v(6);
      for (int i = 0; i < 6; i++) v(i) = s * p[i];
      return v;
    };
    // integrated_matching_cost_factor.cpp:46-52
    if (is_binary_) return LinearFactorPtr(new gtsam::HessianFactor(this->keys()[0], this->keys()[1], M(l.H_target), M(l.H_target_source), V(l.b_target, -1.0), M(l.H_source), V(l.b_source, -1.0), l.error));
    return LinearFactorPtr(new gtsam::HessianFactor(this->keys()[0], M(l.H_source), V(l.b_source, -1.0), l.error));
#else
    auto h = std::make_shared<HessianFactor>();
    h->keys = this->keys();
    for (int i = 0; i < 36; i++) {
      h->G11[i] = l.H_target[i];
      h->G12[i] = l.H_target_source[i];
      h->G22[i] = l.H_source[i];
    }
    for (int i = 0; i < 6; i++) {
      h->g1[i] = -l.b_target[i];
      h->g2[i] = -l.b_source[i];
    }
    h->f = l.error;
    return h;
#endif
  }

  friend class NonlinearFactorSetGPU;
  bool is_binary_;
  Mat4 fixed_target_pose_;
  b2_factor* factor_ = nullptr;
  mutable b2_linearized linearized_{};
  mutable bool has_linearization_ = false;
  mutable double evaluation_ = 0.0;
  mutable bool has_evaluation_ = false;
};

#ifdef B2_HAVE_GTSAM
using FixedPose = gtsam::Pose3;  // unary constructors take the reference's `const gtsam::Pose3& fixed_target_pose`
#else
using FixedPose = Mat4;
#endif

class IntegratedVGICPFactor : public IntegratedMatchingCostFactorB200 {
public:
  using shared_ptr = std::shared_ptr<IntegratedVGICPFactor>;
  // factors/integrated_vgicp_factor.hpp:37-54 / integrated_vgicp_factor_gpu.hpp:43-66: (target_key, source_key, target_voxels, source
  // [, stream, temp_buffer]).  The stream / temp-buffer arguments of the GPU factor are accepted and ignored: the stream is the
  // context's, and the kernel needs no temporary storage.
  IntegratedVGICPFactor(Key target_key, Key source_key, const GaussianVoxelMap::ConstPtr& target_voxels, const PointCloudGPU::ConstPtr& source, void* /*stream*/ = nullptr,
                        std::shared_ptr<void> /*temp_buffer*/ = nullptr)
  : IntegratedMatchingCostFactorB200(target_key, source_key), target_voxels_(std::dynamic_pointer_cast<const GaussianVoxelMapGPU>(target_voxels)), source_(source) {
    init();
  }
  IntegratedVGICPFactor(const FixedPose& fixed_target_pose, Key source_key, const GaussianVoxelMap::ConstPtr& target_voxels, const PointCloudGPU::ConstPtr& source, void* = nullptr,
                        std::shared_ptr<void> = nullptr)
  : IntegratedMatchingCostFactorB200(pose_matrix(fixed_target_pose), source_key), target_voxels_(std::dynamic_pointer_cast<const GaussianVoxelMapGPU>(target_voxels)), source_(source) {
    init();
  }
  // the reference's frame types (gtsam_points::PointCloud et al.): uploaded once, shared between factors
  template <typename Frame, typename = decltype(std::declval<const Frame&>().points)>
  IntegratedVGICPFactor(Key target_key, Key source_key, const GaussianVoxelMap::ConstPtr& target_voxels, const std::shared_ptr<const Frame>& source)
  : IntegratedVGICPFactor(target_key, source_key, target_voxels, PointCloudGPU::from_frame(source, map_context(target_voxels))) {}
  template <typename Frame, typename = decltype(std::declval<const Frame&>().points)>
  IntegratedVGICPFactor(const FixedPose& fixed_target_pose, Key source_key, const GaussianVoxelMap::ConstPtr& target_voxels, const std::shared_ptr<const Frame>& source)
  : IntegratedVGICPFactor(fixed_target_pose, source_key, target_voxels, PointCloudGPU::from_frame(source, map_context(target_voxels))) {}

  void print(const std::string& s = "", const KeyFormatter& keyFormatter = &default_key_format) const override {
    print_keys(s, "IntegratedVGICPFactor", keyFormatter);
    std::cout << "|source|=" << source_->size() << "pts, target resolution=" << target_voxels_->voxel_resolution() << std::endl;
  }
  GaussianVoxelMapGPU::ConstPtr get_target() const { return target_voxels_; }
  // integrated_vgicp_factor.hpp:71-90: the device path has no thread knob (no-op, like the reference's GPU factor) and no cache
  void set_num_threads(int) {}
  void set_fused_cov_cache_mode(FusedCovCacheMode) {}
  // integrated_vgicp_factor_gpu.hpp:83-95: inlier-list maintenance of the reference's two-pass GPU path; the fused kernel
  // compacts hits in-flight, so there is no list to maintain (accepted for source compatibility)
  void set_inlier_update_thresh(double, double) {}
  void set_enable_surface_validation(bool enable) {
    if (enable) std::cerr << "warning: surface-normal validation is not part of the B200 VGICP path (ignored)" << std::endl;
  }
  // NonlinearFactor::clone (integrated_vgicp_factor.hpp:90): a new factor over the same (shared) target map and source cloud
  FactorBasePtr clone() const override {
    if (is_binary_) return FactorBasePtr(new IntegratedVGICPFactor(this->keys()[0], this->keys()[1], target_voxels_, source_));
    IntegratedVGICPFactor* f = new IntegratedVGICPFactor(this->keys()[0], this->keys()[0], target_voxels_, source_);
    f->make_unary(fixed_target_pose_);
    return FactorBasePtr(f);
  }
  // device bytes this factor owns itself (correspondence array + linearization point); clouds / maps are shared
  std::size_t memory_usage() const { return b2_factor_num_points(factor_) * sizeof(std::int32_t) + 16 * sizeof(double) + sizeof(*this); }

private:
  static Context::Ptr map_context(const GaussianVoxelMap::ConstPtr& m) {
    auto g = std::dynamic_pointer_cast<const GaussianVoxelMapGPU>(m);
    return g ? g->context() : Context::default_context();
  }
  void make_unary(const Mat4& fixed) {
    is_binary_ = false;
    fixed_target_pose_ = fixed;
    keys_.erase(keys_.begin());
  }
  void init() {
    // same precondition messages as factors/impl/integrated_vgicp_factor_impl.hpp:32-45 and integrated_vgicp_factor_gpu.cpp:33-46
    if (!source_ || !source_->has_points()) {
      std::cerr << "error: source points have not been allocated!!" << std::endl;
      abort();
    }
    if (!source_->has_covs()) {
      std::cerr << "error: source don't have covs!!" << std::endl;
      abort();
    }
    if (!target_voxels_ || !target_voxels_->handle()) {
      std::cerr << "error: target voxelmap has not been created!!" << std::endl;
      abort();
    }
    check(b2_vgicp_factor_create(source_->context()->get(), target_voxels_->handle(), source_->handle(), &factor_), "b2_vgicp_factor_create");
  }
  GaussianVoxelMapGPU::ConstPtr target_voxels_;
  PointCloudGPU::ConstPtr source_;
};
using IntegratedVGICPFactorGPU = IntegratedVGICPFactor;

class IntegratedGICPFactor : public IntegratedMatchingCostFactorB200 {
public:
  using shared_ptr = std::shared_ptr<IntegratedGICPFactor>;
  // factors/integrated_gicp_factor.hpp:44-78: four constructors (binary / unary) x (with / without a search tree)
  IntegratedGICPFactor(Key target_key, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const KdTreeGPU::ConstPtr& target_tree)
  : IntegratedMatchingCostFactorB200(target_key, source_key), target_(target), source_(source), tree_(target_tree) {
    init();
  }
  IntegratedGICPFactor(const FixedPose& fixed_target_pose, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const KdTreeGPU::ConstPtr& target_tree)
  : IntegratedMatchingCostFactorB200(pose_matrix(fixed_target_pose), source_key), target_(target), source_(source), tree_(target_tree) {
    init();
  }
  // frames of the reference's shape; the reference builds a KdTree2 over the target when no tree is given
  // (integrated_gicp_factor_impl.hpp:47-51) -- here that tree is built on the device
  template <typename TargetFrame, typename SourceFrame, typename = decltype(std::declval<const TargetFrame&>().points), typename = decltype(std::declval<const SourceFrame&>().points)>
  IntegratedGICPFactor(Key target_key, Key source_key, const std::shared_ptr<const TargetFrame>& target, const std::shared_ptr<const SourceFrame>& source, Context::Ptr ctx = Context::default_context())
  : IntegratedGICPFactor(target_key, source_key, PointCloudGPU::from_frame(target, ctx), PointCloudGPU::from_frame(source, ctx),
                         std::make_shared<KdTreeGPU>(reinterpret_cast<const double*>(target->points), 4, target->size(), ctx)) {}
  template <typename TargetFrame, typename SourceFrame, typename = decltype(std::declval<const TargetFrame&>().points), typename = decltype(std::declval<const SourceFrame&>().points)>
  IntegratedGICPFactor(const FixedPose& fixed_target_pose, Key source_key, const std::shared_ptr<const TargetFrame>& target, const std::shared_ptr<const SourceFrame>& source,
                       Context::Ptr ctx = Context::default_context())
  : IntegratedGICPFactor(fixed_target_pose, source_key, PointCloudGPU::from_frame(target, ctx), PointCloudGPU::from_frame(source, ctx),
                         std::make_shared<KdTreeGPU>(reinterpret_cast<const double*>(target->points), 4, target->size(), ctx)) {}
  IntegratedGICPFactor(Key target_key, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const double* target_points, int target_point_stride)
  : IntegratedGICPFactor(target_key, source_key, target, source, std::make_shared<KdTreeGPU>(target_points, target_point_stride, target->size(), target->context())) {}

  void print(const std::string& s = "", const KeyFormatter& keyFormatter = &default_key_format) const override {
    print_keys(s, "IntegratedGICPFactor", keyFormatter);
    std::cout << "|target|=" << target_->size() << "pts, |source|=" << source_->size() << "pts" << std::endl;
  }
  void set_num_threads(int) {}  // kept for source compatibility; the device path has no thread knob
  void set_fused_cov_cache_mode(FusedCovCacheMode) {}
  void set_max_correspondence_distance(double dist) { check(b2_factor_set_max_correspondence_distance(factor_, dist), "b2_factor_set_max_correspondence_distance"); }
  // integrated_gicp_factor.hpp:103-109, impl:135-147: re-association is SKIPPED while the pose stays within (angle, trans) of
  // the pose of the last correspondence update; the frozen correspondences are then linearized at the new pose
  void set_correspondence_update_tolerance(double angle, double trans) {
    check(b2_factor_set_correspondence_update_tolerance(factor_, angle, trans), "b2_factor_set_correspondence_update_tolerance");
  }
  FactorBasePtr clone() const override {
    if (is_binary_) return FactorBasePtr(new IntegratedGICPFactor(this->keys()[0], this->keys()[1], target_, source_, tree_));
    IntegratedGICPFactor* f = new IntegratedGICPFactor(this->keys()[0], this->keys()[0], target_, source_, tree_);
    f->is_binary_ = false;
    f->fixed_target_pose_ = fixed_target_pose_;
    f->keys_.erase(f->keys_.begin());
    return FactorBasePtr(f);
  }
  std::size_t memory_usage() const {
    return b2_factor_num_points(factor_) * sizeof(std::int32_t) + 16 * sizeof(double) + target_->size() * 10 * sizeof(double) + sizeof(*this);
  }

private:
  void init() {
    if (!source_ || !source_->has_covs() || !target_ || !target_->has_covs() || !tree_) {
      std::cerr << "error: target or source points / covs / search tree have not been allocated!!" << std::endl;
      abort();
    }
    check(b2_gicp_factor_create(source_->context()->get(), target_->handle(), tree_->handle(), source_->handle(), &factor_), "b2_gicp_factor_create");
  }
  PointCloudGPU::ConstPtr target_, source_;
  KdTreeGPU::ConstPtr tree_;
};

// IntegratedICPFactor_ / IntegratedPointToPlaneICPFactor_ (include/gtsam_points/factors/integrated_icp_factor.hpp:27-145): the
// kd-tree kernel with M = I (point-to-point) or, with use_point_to_plane, the residual scaled row-wise by the target normal
// (target_normals: host array, n x 3, caller order).  Neither cloud needs covariances.
class IntegratedICPFactor : public IntegratedMatchingCostFactorB200 {
public:
  using shared_ptr = std::shared_ptr<IntegratedICPFactor>;
  IntegratedICPFactor(Key target_key, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const KdTreeGPU::ConstPtr& target_tree,
                      bool use_point_to_plane = false, const double* target_normals = nullptr)
  : IntegratedMatchingCostFactorB200(target_key, source_key), target_(target), source_(source), tree_(target_tree), plane_(use_point_to_plane) {
    init(target_normals);
  }
  IntegratedICPFactor(const FixedPose& fixed_target_pose, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source,
                      const KdTreeGPU::ConstPtr& target_tree, bool use_point_to_plane = false, const double* target_normals = nullptr)
  : IntegratedMatchingCostFactorB200(pose_matrix(fixed_target_pose), source_key), target_(target), source_(source), tree_(target_tree), plane_(use_point_to_plane) {
    init(target_normals);
  }
  // frames of the reference's shape (points as Vector4d arrays); the search tree is built on the device
  template <typename TargetFrame, typename SourceFrame, typename = decltype(std::declval<const TargetFrame&>().points), typename = decltype(std::declval<const SourceFrame&>().points)>
  IntegratedICPFactor(Key target_key, Key source_key, const std::shared_ptr<const TargetFrame>& target, const std::shared_ptr<const SourceFrame>& source, Context::Ptr ctx = Context::default_context())
  : IntegratedICPFactor(target_key, source_key, PointCloudGPU::from_frame(target, ctx), PointCloudGPU::from_frame(source, ctx),
                        std::make_shared<KdTreeGPU>(reinterpret_cast<const double*>(target->points), 4, target->size(), ctx)) {}

  void print(const std::string& s = "", const KeyFormatter& keyFormatter = &default_key_format) const override {
    print_keys(s, plane_ ? "IntegratedPointToPlaneICPFactor" : "IntegratedICPFactor", keyFormatter);
    std::cout << "|target|=" << target_->size() << "pts, |source|=" << source_->size() << "pts" << std::endl;
  }
  void set_num_threads(int) {}
  void set_max_correspondence_distance(double dist) { check(b2_factor_set_max_correspondence_distance(factor_, dist), "b2_factor_set_max_correspondence_distance"); }
  void set_correspondence_update_tolerance(double angle, double trans) {
    check(b2_factor_set_correspondence_update_tolerance(factor_, angle, trans), "b2_factor_set_correspondence_update_tolerance");
  }
  // NonlinearFactor::clone: a new factor over the same (shared) target, source and search tree
  FactorBasePtr clone() const override {
    const double* nrm = normals_.empty() ? nullptr : normals_.data();
    if (is_binary_) return FactorBasePtr(new IntegratedICPFactor(this->keys()[0], this->keys()[1], target_, source_, tree_, plane_, nrm));
    IntegratedICPFactor* f = new IntegratedICPFactor(this->keys()[0], this->keys()[0], target_, source_, tree_, plane_, nrm);
    f->is_binary_ = false;
    f->fixed_target_pose_ = fixed_target_pose_;
    f->keys_.erase(f->keys_.begin());
    return FactorBasePtr(f);
  }

private:
  void init(const double* target_normals) {
    if (!source_ || !target_ || !tree_ || (plane_ && target_normals == nullptr)) {
      std::cerr << "error: target frame doesn't have required attributes for icp" << std::endl;  // integrated_icp_factor_impl.hpp:36-50
      abort();
    }
    if (plane_) normals_.assign(target_normals, target_normals + 3 * target_->size());  // kept for clone()
    check(b2_icp_factor_create(source_->context()->get(), target_->handle(), tree_->handle(), source_->handle(), plane_ ? 1 : 0, target_normals, &factor_), "b2_icp_factor_create");
  }
  PointCloudGPU::ConstPtr target_, source_;
  KdTreeGPU::ConstPtr tree_;
  bool plane_;
  std::vector<double> normals_;
};

// estimate_covariances (include/gtsam_points/features/covariance_estimation.hpp:41-66) on the device: k nearest neighbours over a
// kd-tree built on the spot, neighbourhood covariance, eigenvalues replaced by (1e-3, 1, 1).  Returns n row-major 3x3 matrices.
inline std::vector<std::array<double, 9>> estimate_covariances(const double* points, int point_stride, std::size_t n, int k_neighbors = 10,
                                                               Context::Ptr ctx = Context::default_context()) {
  std::vector<std::array<double, 9>> covs(n);
  if (n) check(b2_estimate_covariances(ctx->get(), points, point_stride, n, k_neighbors, nullptr, covs[0].data()), "b2_estimate_covariances");
  return covs;
}

// merge_frames_gpu (include/gtsam_points/types/gaussian_voxelmap_gpu.hpp:127-150): posed frames (world <- frame) -> one cloud
// downsampled on a voxel grid laid out in the first frame's coordinates; bit-identical to the CPU merge_frames.
inline PointCloudGPU::Ptr merge_frames_gpu(const std::vector<Mat4>& poses, const std::vector<PointCloudGPU::ConstPtr>& frames, double downsample_resolution,
                                           Context::Ptr ctx = Context::default_context()) {
  if (poses.size() != frames.size() || frames.empty()) throw std::invalid_argument("merge_frames_gpu: one pose per frame, at least one frame");
  std::vector<double> P;
  std::vector<const b2_cloud*> handles;
  std::size_t total = 0;
  for (std::size_t i = 0; i < frames.size(); i++) {
    P.insert(P.end(), poses[i].begin(), poses[i].end());
    handles.push_back(frames[i]->handle());
    total += frames[i]->size();
  }
  std::vector<double> xyz(3 * std::max<std::size_t>(total, 1)), cov(9 * std::max<std::size_t>(total, 1));
  std::size_t m = 0;
  check(b2_merge_frames(ctx->get(), P.data(), handles.data(), frames.size(), downsample_resolution, xyz.data(), cov.data(), &m), "b2_merge_frames");
  return PointCloudGPU::from_packed(xyz.data(), cov.data(), m, ctx);
}

// ---------------------------------------------------------------------------------------------------------------------
// NonlinearFactorSet implementation: every device factor of the graph in ONE batched launch
// (replaces src/gtsam_points/cuda/nonlinear_factor_set_gpu.cpp:48-228; interface optimizers/linearization_hook.hpp:11-29)
// ---------------------------------------------------------------------------------------------------------------------
#ifdef B2_HAVE_LINEARIZATION_HOOK
using FactorSetBase = gtsam_points::NonlinearFactorSet;
#else
struct FactorSetBase {  // the same nine virtuals (add(graph) exists with GTSAM only)
  virtual ~FactorSetBase() {}
  virtual int size() const = 0;
  virtual void clear() = 0;
  virtual void clear_counts() = 0;
  virtual int linearization_count() const = 0;
  virtual int evaluation_count() const = 0;
  virtual bool add(FactorBasePtr factor) = 0;
  virtual void linearize(const Values& values) = 0;
  virtual void error(const Values& values) = 0;
  virtual std::vector<LinearFactorPtr> calc_linear_factors(const Values& linearization_point) = 0;
};
#endif

class NonlinearFactorSetGPU : public FactorSetBase {
public:
  using FactorPtr = decltype(factor_pointer_cast<IntegratedMatchingCostFactorB200>(std::declval<FactorBasePtr>()));

  explicit NonlinearFactorSetGPU(Context::Ptr ctx = Context::default_context()) : ctx_(ctx) {}
  ~NonlinearFactorSetGPU() override { b2_factor_set_destroy(set_); }

  int size() const override { return static_cast<int>(factors_.size()); }
  void clear() override {
    b2_factor_set_destroy(set_);
    set_ = nullptr;
    factors_.clear();
  }
  void clear_counts() override { num_linearizations_ = num_evaluations_ = 0; }
  int linearization_count() const override { return num_linearizations_; }
  int evaluation_count() const override { return num_evaluations_; }

  // returns false for factors this set cannot batch (the optimizer then linearizes them itself): nonlinear_factor_set_gpu.cpp:48-57
  bool add(FactorBasePtr factor) override {
    auto f = factor_pointer_cast<IntegratedMatchingCostFactorB200>(factor);
    if (!f) return false;
    b2_factor_set_destroy(set_);
    set_ = nullptr;
    factors_.push_back(f);
    return true;
  }
#ifdef B2_HAVE_GTSAM
  void add(const gtsam::NonlinearFactorGraph& factors)
#ifdef B2_HAVE_LINEARIZATION_HOOK
    override
#endif
  {
    for (const auto& f : factors) add(f);  // nonlinear_factor_set_gpu.cpp:59-63
  }
#endif

  void linearize(const Values& values) override {
    if (factors_.empty()) return;
    ensure();
    num_linearizations_ += size();
    pack(values);
    results_.resize(factors_.size());
    check(b2_factor_set_linearize(set_, deltas_.data(), results_.data()), "b2_factor_set_linearize");
    for (std::size_t i = 0; i < factors_.size(); i++) factors_[i]->store_linearized(&results_[i]);
  }

  void error(const Values& values) override {
    if (factors_.empty()) return;
    ensure();
    num_evaluations_ += size();
    pack(values);
    errors_.resize(factors_.size());
    check(b2_factor_set_error(set_, deltas_.data(), errors_.data()), "b2_factor_set_error");
    for (std::size_t i = 0; i < factors_.size(); i++) factors_[i]->store_computed_error(&errors_[i]);
  }

  std::vector<LinearFactorPtr> calc_linear_factors(const Values& linearization_point) override {
    linearize(linearization_point);
    std::vector<LinearFactorPtr> out(factors_.size());
    for (std::size_t i = 0; i < factors_.size(); i++) out[i] = factors_[i]->linearize(linearization_point);
    return out;
  }

private:
  void ensure() {
    if (set_) return;
    std::vector<b2_factor*> hs(factors_.size());
    for (std::size_t i = 0; i < factors_.size(); i++) hs[i] = factors_[i]->handle();
    check(b2_factor_set_create(ctx_->get(), hs.data(), hs.size(), &set_), "b2_factor_set_create");
  }
  void pack(const Values& values) {
    deltas_.resize(factors_.size() * 16);
    for (std::size_t i = 0; i < factors_.size(); i++) {
      const Mat4 d = factors_[i]->calc_delta(values);
      std::copy(d.begin(), d.end(), deltas_.begin() + i * 16);
    }
  }
  Context::Ptr ctx_;
  std::vector<FactorPtr> factors_;
  b2_factor_set* set_ = nullptr;
  std::vector<double> deltas_, errors_;
  std::vector<b2_linearized> results_;
  int num_linearizations_ = 0, num_evaluations_ = 0;
};

// what user code hands to LinearizationHook::register_hook (src/demo/demo_matching_cost_factors.cpp:52)
inline std::shared_ptr<FactorSetBase> create_nonlinear_factor_set_gpu() { return std::make_shared<NonlinearFactorSetGPU>(); }

}  // namespace gtsam_points_b200
