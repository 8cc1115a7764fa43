// MOCK (see ../../Eigen/Core).  GTSAM 4.3 flavour: std::shared_ptr.  Compile with -DB2_MOCK_GTSAM_BOOST_PTR to get a distinct
// smart-pointer type in namespace boost_like, as GTSAM 4.2 (boost::shared_ptr) would give the adapters.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <Eigen/Core>
namespace gtsam {
using Key = std::uint64_t;
using KeyVector = std::vector<Key>;
using KeyFormatter = std::function<std::string(Key)>;
inline std::string _defaultKeyFormatter(Key k) { return std::to_string(k); }
static const KeyFormatter DefaultKeyFormatter = &_defaultKeyFormatter;
using Matrix = Eigen::MatrixXd;
using Vector = Eigen::VectorXd;
using Matrix4 = Eigen::Matrix4d;
}  // namespace gtsam
