// gtsam_points_b200.hpp -- header-only C++ adapters that keep gtsam_points' C++ surface for the scan-matching path on
// top of the C ABI (include/b2points.h, libb2points.so).
//
//   gtsam_points_b200::PointCloudGPU            <- gtsam_points::PointCloudGPU / PointCloud   (types/point_cloud.hpp:19-119)
//   gtsam_points_b200::GaussianVoxelMapGPU      <- gtsam_points::GaussianVoxelMapGPU          (types/gaussian_voxelmap_gpu.hpp:39-108)
//   gtsam_points_b200::KdTreeGPU                <- gtsam_points::KdTree / NearestNeighborSearch (ann/nearest_neighbor_search.hpp:16-57)
//   gtsam_points_b200::IntegratedVGICPFactor    <- IntegratedVGICPFactor_ / IntegratedVGICPFactorGPU (factors/integrated_vgicp_factor.hpp:25-113)
//   gtsam_points_b200::IntegratedGICPFactor     <- IntegratedGICPFactor_                       (factors/integrated_gicp_factor.hpp:32-152)
//   gtsam_points_b200::NonlinearFactorSetGPU    <- gtsam_points::NonlinearFactorSetGPU / NonlinearFactorSet (optimizers/linearization_hook.hpp:11-29)
//
// With GTSAM on the include path the factors derive from gtsam::NonlinearFactor and linearize() returns a
// gtsam::HessianFactor, so they drop into LevenbergMarquardtOptimizerExt / ISAM2Ext unchanged, and
// NonlinearFactorSetGPU can be registered through LinearizationHook::register_hook.
// Without GTSAM (this image has none) the same classes are built on a 30-line Pose3 / Values stand-in so that the call
// sequence can be compiled and tested (tests/cpp/test_adapters.cpp).
//
// Error behaviour mirrors the reference: constructor preconditions print the reference's message and abort()
// (factors/impl/integrated_vgicp_factor_impl.hpp:32-45); runtime CUDA failures throw std::runtime_error with b2_last_error().
#pragma once

#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
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
inline Mat4 pose_matrix(const Values& values, Key key) {
  const gtsam::Matrix4 m = values.at<gtsam::Pose3>(key).matrix();
  Mat4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r[i * 4 + j] = m(i, j);
  return r;
}
#else
using Key = std::uint64_t;
struct Values {
  std::map<Key, Mat4> poses;
  void insert(Key k, const Mat4& T) { poses[k] = T; }
  const Mat4& at(Key k) const { return poses.at(k); }
};
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
  b2_ctx* ctx_ = nullptr;
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

class GaussianVoxelMapGPU {
public:
  using Ptr = std::shared_ptr<GaussianVoxelMapGPU>;
  using ConstPtr = std::shared_ptr<const GaussianVoxelMapGPU>;

  explicit GaussianVoxelMapGPU(double resolution, Context::Ptr ctx = Context::default_context()) : ctx_(ctx), resolution_(resolution) {}
  ~GaussianVoxelMapGPU() { b2_voxelmap_destroy(vm_); }
  GaussianVoxelMapGPU(const GaussianVoxelMapGPU&) = delete;

  double voxel_resolution() const { return resolution_; }

  // GaussianVoxelMap::insert(const PointCloud&): one-shot on the GPU (types/gaussian_voxelmap_gpu.hpp:63)
  void insert(const double* points, int point_stride, const double* covs, int cov_stride, std::size_t n) {
    if (vm_) {
      std::cerr << "error: incremental insertion is not supported for GPU voxelmaps" << std::endl;
      abort();
    }
    check(b2_voxelmap_create_from_points(ctx_->get(), resolution_, points, point_stride, covs, cov_stride, n, &vm_), "b2_voxelmap_create_from_points");
  }
  std::size_t num_voxels() const {
    b2_voxelmap_info info{};
    if (vm_) check(b2_voxelmap_get_info(vm_, &info), "b2_voxelmap_get_info");
    return info.num_voxels;
  }
  b2_voxelmap* handle() const { return vm_; }

private:
  Context::Ptr ctx_;
  double resolution_;
  b2_voxelmap* vm_ = nullptr;
};

// NearestNeighborSearch::knn_search for k = 1 (exact); batch entry point preferred on the device
class KdTreeGPU {
public:
  using Ptr = std::shared_ptr<KdTreeGPU>;
  using ConstPtr = std::shared_ptr<const KdTreeGPU>;
  KdTreeGPU(const double* points, int point_stride, std::size_t n, Context::Ptr ctx = Context::default_context()) : ctx_(ctx) {
    check(b2_kdtree_create(ctx_->get(), points, point_stride, n, &tree_), "b2_kdtree_create");
  }
  ~KdTreeGPU() { b2_kdtree_destroy(tree_); }
  KdTreeGPU(const KdTreeGPU&) = delete;

  // same signature as NearestNeighborSearch::knn_search (ann/nearest_neighbor_search.hpp:31-35); only k == 1 is accelerated
  std::size_t knn_search(const double* pt, std::size_t k, std::size_t* k_indices, double* k_sq_dists, double max_sq_dist = 1e300) const {
    if (k != 1) throw std::invalid_argument("KdTreeGPU::knn_search: k must be 1");
    std::int64_t idx = -1;
    check(b2_kdtree_knn1(tree_, pt, 3, 1, max_sq_dist, &idx, k_sq_dists), "b2_kdtree_knn1");
    k_indices[0] = idx < 0 ? static_cast<std::size_t>(-1) : static_cast<std::size_t>(idx);
    return idx < 0 ? 0 : 1;
  }
  void knn_search_batch(const double* queries, int stride, std::size_t n, double max_sq_dist, std::int64_t* indices, double* sq_dists) const {
    check(b2_kdtree_knn1(tree_, queries, stride, n, max_sq_dist, indices, sq_dists), "b2_kdtree_knn1");
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
#else
struct HessianFactor {  // stand-in for gtsam::HessianFactor(k_t, k_s, G11, G12, g1, G22, g2, f) / (k_s, G22, g2, f)
  std::vector<Key> keys;
  double G11[36], G12[36], G22[36], g1[6], g2[6], f;
};
struct FactorBase {
  explicit FactorBase(std::vector<Key> keys) : keys_(std::move(keys)) {}
  virtual ~FactorBase() {}
  const std::vector<Key>& keys() const { return keys_; }
  std::vector<Key> keys_;
};
using LinearFactorPtr = std::shared_ptr<HessianFactor>;
#endif

class NonlinearFactorSetGPU;

// factors/integrated_gicp_factor.hpp:20-24.  The device path recomputes M = (C_B + R C_A R^T)^-1 from the rotation of the
// linearization point inside the kernel (the reference's NONE behaviour, no cache to keep in HBM), so the mode is accepted
// for source compatibility and has no effect on results beyond rounding.
enum class FusedCovCacheMode { FULL, COMPACT, NONE };

class IntegratedMatchingCostFactorB200 : public FactorBase {
public:
  ~IntegratedMatchingCostFactorB200() override { b2_factor_destroy(factor_); }
  std::size_t dim() const
#ifdef B2_HAVE_GTSAM
    override
#endif
  {
    return 6;
  }

  Mat4 calc_delta(const Values& values) const {
    if (is_binary_) return gtsam_points_b200::calc_delta(pose_matrix(values, this->keys()[0]), pose_matrix(values, this->keys()[1]));
    return gtsam_points_b200::calc_delta(fixed_target_pose_, pose_matrix(values, this->keys()[0]));
  }

  // NonlinearFactor::error -- re-uses the correspondences / fused covariances frozen at the last linearize()
  double error(const Values& values) const
#ifdef B2_HAVE_GTSAM
    override
#endif
  {
    if (has_evaluation_) {  // filled by NonlinearFactorSetGPU::error, like IntegratedVGICPFactorGPU::evaluation_result
      has_evaluation_ = false;
      return evaluation_;
    }
    const Mat4 d = calc_delta(values);
    double e = 0.0;
    check(b2_factor_error(factor_, d.data(), &e), "b2_factor_error");
    return e;
  }

  LinearFactorPtr linearize(const Values& values) const
#ifdef B2_HAVE_GTSAM
    override
#endif
  {
    if (!has_linearization_) {  // not batched through a NonlinearFactorSetGPU: evaluate now
      const Mat4 d = calc_delta(values);
      check(b2_factor_linearize(factor_, d.data(), &linearized_), "b2_factor_linearize");
    }
    has_linearization_ = false;
    return make_hessian(linearized_);
  }

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
  IntegratedMatchingCostFactorB200(Key target_key, Key source_key) : FactorBase(std::vector<Key>{target_key, source_key}), is_binary_(true) { fixed_target_pose_ = Mat4{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
  IntegratedMatchingCostFactorB200(const Mat4& fixed_target_pose, Key source_key) : FactorBase(std::vector<Key>{source_key}), is_binary_(false), fixed_target_pose_(fixed_target_pose) {}

  LinearFactorPtr make_hessian(const b2_linearized& l) const {
#ifdef B2_HAVE_GTSAM
    auto M = [](const double* p) { return gtsam::Matrix(Eigen::Map<const Eigen::Matrix<double, 6, 6, Eigen::RowMajor>>(p)); };
    auto V = [](const double* p, double s) { return gtsam::Vector(s * Eigen::Map<const Eigen::Matrix<double, 6, 1>>(p)); };
    // integrated_matching_cost_factor.cpp:46-52
    if (is_binary_) return LinearFactorPtr(new gtsam::HessianFactor(keys()[0], keys()[1], M(l.H_target), M(l.H_target_source), V(l.b_target, -1.0), M(l.H_source), V(l.b_source, -1.0), l.error));
    return LinearFactorPtr(new gtsam::HessianFactor(keys()[0], M(l.H_source), V(l.b_source, -1.0), l.error));
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

class IntegratedVGICPFactor : public IntegratedMatchingCostFactorB200 {
public:
  using shared_ptr = std::shared_ptr<IntegratedVGICPFactor>;
  IntegratedVGICPFactor(Key target_key, Key source_key, const GaussianVoxelMapGPU::ConstPtr& target_voxels, const PointCloudGPU::ConstPtr& source)
  : IntegratedMatchingCostFactorB200(target_key, source_key), target_voxels_(target_voxels), source_(source) {
    init();
  }
  IntegratedVGICPFactor(const Mat4& fixed_target_pose, Key source_key, const GaussianVoxelMapGPU::ConstPtr& target_voxels, const PointCloudGPU::ConstPtr& source)
  : IntegratedMatchingCostFactorB200(fixed_target_pose, source_key), target_voxels_(target_voxels), source_(source) {
    init();
  }
  GaussianVoxelMapGPU::ConstPtr get_target() const { return target_voxels_; }
  // integrated_vgicp_factor.hpp:71-90 -- kept for source compatibility: the device path has no thread knob and no cache
  void set_num_threads(int) {}
  void set_fused_cov_cache_mode(FusedCovCacheMode) {}
  // NonlinearFactor::clone (integrated_vgicp_factor.hpp:90): a new factor over the same (shared) target map and source cloud
  shared_ptr clone() const {
    return is_binary_ ? std::make_shared<IntegratedVGICPFactor>(this->keys()[0], this->keys()[1], target_voxels_, source_)
                      : std::make_shared<IntegratedVGICPFactor>(fixed_target_pose_, this->keys()[0], target_voxels_, source_);
  }
  // device bytes this factor owns itself (correspondence array + linearization point); clouds / maps are shared
  std::size_t memory_usage() const { return b2_factor_num_points(factor_) * sizeof(std::int32_t) + 16 * sizeof(double) + sizeof(*this); }

private:
  void init() {
    // same precondition messages as factors/impl/integrated_vgicp_factor_impl.hpp:32-45
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
  IntegratedGICPFactor(Key target_key, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const KdTreeGPU::ConstPtr& target_tree)
  : IntegratedMatchingCostFactorB200(target_key, source_key), target_(target), source_(source), tree_(target_tree) {
    init();
  }
  IntegratedGICPFactor(const Mat4& fixed_target_pose, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const KdTreeGPU::ConstPtr& target_tree)
  : IntegratedMatchingCostFactorB200(fixed_target_pose, source_key), target_(target), source_(source), tree_(target_tree) {
    init();
  }
  // the reference builds a KdTree2 over the target when no search tree is given (integrated_gicp_factor_impl.hpp:47-51);
  // here that tree is built on the device
  IntegratedGICPFactor(Key target_key, Key source_key, const PointCloudGPU::ConstPtr& target, const PointCloudGPU::ConstPtr& source, const double* target_points,
                       int target_point_stride)
  : IntegratedGICPFactor(target_key, source_key, target, source, std::make_shared<KdTreeGPU>(target_points, target_point_stride, target->size(), target->context())) {}
  void set_num_threads(int) {}  // kept for source compatibility; the device path has no thread knob
  void set_fused_cov_cache_mode(FusedCovCacheMode) {}
  void set_max_correspondence_distance(double dist) { check(b2_factor_set_max_correspondence_distance(factor_, dist), "b2_factor_set_max_correspondence_distance"); }
  // integrated_gicp_factor.hpp:103-109: the reference may SKIP re-association when the pose moved less than these
  // tolerances since the last update (default 0 = always update).  The device path always re-associates (the search is
  // fused into the linearization kernel and costs less than the skip would save), i.e. it behaves like tolerance 0; the
  // values are kept so that callers can read them back.
  void set_correspondence_update_tolerance(double angle, double trans) {
    correspondence_update_tolerance_rot_ = angle;
    correspondence_update_tolerance_trans_ = trans;
  }
  shared_ptr clone() const {
    return is_binary_ ? std::make_shared<IntegratedGICPFactor>(this->keys()[0], this->keys()[1], target_, source_, tree_)
                      : std::make_shared<IntegratedGICPFactor>(fixed_target_pose_, this->keys()[0], target_, source_, tree_);
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
  double correspondence_update_tolerance_rot_ = 0.0, correspondence_update_tolerance_trans_ = 0.0;
};

// ---------------------------------------------------------------------------------------------------------------------
// NonlinearFactorSet implementation: every device factor of the graph in ONE batched launch
// (replaces src/gtsam_points/cuda/nonlinear_factor_set_gpu.cpp:48-228; interface optimizers/linearization_hook.hpp:11-29)
// ---------------------------------------------------------------------------------------------------------------------
class NonlinearFactorSetGPU {
public:
  explicit NonlinearFactorSetGPU(Context::Ptr ctx = Context::default_context()) : ctx_(ctx) {}
  ~NonlinearFactorSetGPU() { b2_factor_set_destroy(set_); }

  int size() const { return static_cast<int>(factors_.size()); }
  void clear() {
    b2_factor_set_destroy(set_);
    set_ = nullptr;
    factors_.clear();
  }
  void clear_counts() { num_linearizations_ = num_evaluations_ = 0; }
  int linearization_count() const { return num_linearizations_; }
  int evaluation_count() const { return num_evaluations_; }

  // returns false for factors this set cannot batch (the optimizer then linearizes them itself)
  template <typename FactorPtr>
  bool add(const FactorPtr& factor) {
    auto f = std::dynamic_pointer_cast<IntegratedMatchingCostFactorB200>(factor);
    if (!f) return false;
    b2_factor_set_destroy(set_);
    set_ = nullptr;
    factors_.push_back(f);
    return true;
  }

  void linearize(const Values& values) {
    if (factors_.empty()) return;
    ensure();
    num_linearizations_ += size();
    pack(values);
    results_.resize(factors_.size());
    check(b2_factor_set_linearize(set_, deltas_.data(), results_.data()), "b2_factor_set_linearize");
    for (std::size_t i = 0; i < factors_.size(); i++) {
      factors_[i]->linearized_ = results_[i];
      factors_[i]->has_linearization_ = true;
      factors_[i]->evaluation_ = results_[i].error;  // store_linearized also primes the evaluation result (integrated_vgicp_factor_gpu.cpp:239-245)
      factors_[i]->has_evaluation_ = true;
    }
  }

  void error(const Values& values) {
    if (factors_.empty()) return;
    ensure();
    num_evaluations_ += size();
    pack(values);
    errors_.resize(factors_.size());
    check(b2_factor_set_error(set_, deltas_.data(), errors_.data()), "b2_factor_set_error");
    for (std::size_t i = 0; i < factors_.size(); i++) {
      factors_[i]->evaluation_ = errors_[i];
      factors_[i]->has_evaluation_ = true;
    }
  }

  std::vector<LinearFactorPtr> calc_linear_factors(const Values& linearization_point) {
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
  std::vector<std::shared_ptr<IntegratedMatchingCostFactorB200>> factors_;
  b2_factor_set* set_ = nullptr;
  std::vector<double> deltas_, errors_;
  std::vector<b2_linearized> results_;
  int num_linearizations_ = 0, num_evaluations_ = 0;
};

}  // namespace gtsam_points_b200
