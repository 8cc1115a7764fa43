#pragma once  // MOCK
#include <gtsam/base/types.h>
namespace gtsam {
class Pose3 {
public:
  Pose3() = default;
  explicit Pose3(const Matrix4& T) : T_(T) {}
  Matrix4 matrix() const { return T_; }
private:
  Matrix4 T_;
};
}  // namespace gtsam
