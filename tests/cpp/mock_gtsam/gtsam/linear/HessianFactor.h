#pragma once  // MOCK: the two constructors src/gtsam_points/factors/integrated_matching_cost_factor.cpp:46-52 uses
#include <gtsam/base/types.h>
namespace gtsam {
class GaussianFactor {
public:
  using shared_ptr = std::shared_ptr<GaussianFactor>;
  virtual ~GaussianFactor() {}
  KeyVector keys_;
  const KeyVector& keys() const { return keys_; }
};
class HessianFactor : public GaussianFactor {
public:
  HessianFactor(Key j, const Matrix& G, const Vector& g, double f) : G22(G), g2(g), f(f) { keys_ = {j}; }
  HessianFactor(Key j1, Key j2, const Matrix& G11, const Matrix& G12, const Vector& g1, const Matrix& G22, const Vector& g2, double f)
  : G11(G11), G12(G12), G22(G22), g1(g1), g2(g2), f(f) {
    keys_ = {j1, j2};
  }
  Matrix G11, G12, G22;
  Vector g1, g2;
  double f;
};
}  // namespace gtsam
