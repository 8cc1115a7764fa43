#pragma once  // MOCK of include/gtsam_points/optimizers/linearization_hook.hpp:11-58 (interface only, same signatures)
#include <functional>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
namespace gtsam_points {
class NonlinearFactorSet {
public:
  virtual ~NonlinearFactorSet() {}
  virtual int size() const = 0;
  virtual void clear() = 0;
  virtual void clear_counts() = 0;
  virtual int linearization_count() const = 0;
  virtual int evaluation_count() const = 0;
  virtual bool add(gtsam::NonlinearFactor::shared_ptr factor) = 0;
  virtual void add(const gtsam::NonlinearFactorGraph& factors) = 0;
  virtual void linearize(const gtsam::Values& values) = 0;
  virtual void error(const gtsam::Values& values) = 0;
  virtual std::vector<gtsam::GaussianFactor::shared_ptr> calc_linear_factors(const gtsam::Values& linearization_point) = 0;
};
class LinearizationHook {
public:
  static void register_hook(const std::function<std::shared_ptr<NonlinearFactorSet>()>& hook) { hook_constructors().push_back(hook); }
  static std::vector<std::function<std::shared_ptr<NonlinearFactorSet>()>>& hook_constructors() {
    static std::vector<std::function<std::shared_ptr<NonlinearFactorSet>()>> v;
    return v;
  }
};
}  // namespace gtsam_points
