#pragma once  // MOCK: the virtual interface of gtsam::NonlinearFactor the reference's factors override
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/nonlinear/Values.h>
namespace gtsam {
class NonlinearFactor {
public:
  using shared_ptr = std::shared_ptr<NonlinearFactor>;
  NonlinearFactor() {}
  template <typename CONTAINER>
  explicit NonlinearFactor(const CONTAINER& keys) : keys_(keys.begin(), keys.end()) {}
  virtual ~NonlinearFactor() {}
  virtual void print(const std::string& s = "", const KeyFormatter& keyFormatter = DefaultKeyFormatter) const {}
  virtual double error(const Values& c) const = 0;
  virtual size_t dim() const = 0;
  virtual std::shared_ptr<GaussianFactor> linearize(const Values& c) const = 0;
  virtual shared_ptr clone() const = 0;
  const KeyVector& keys() const { return keys_; }
protected:
  KeyVector keys_;
};
}  // namespace gtsam
