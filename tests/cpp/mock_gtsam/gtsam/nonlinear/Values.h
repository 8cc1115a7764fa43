#pragma once  // MOCK
#include <map>
#include <gtsam/geometry/Pose3.h>
namespace gtsam {
class Values {
public:
  void insert(Key k, const Pose3& p) { poses_[k] = p; }
  template <typename T>
  const T& at(Key k) const { return poses_.at(k); }
  bool exists(Key k) const { return poses_.count(k) != 0; }
private:
  std::map<Key, Pose3> poses_;
};
}  // namespace gtsam
