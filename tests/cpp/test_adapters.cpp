// test_adapters.cpp -- exercises the C++ adapter classes (gtsam_points_b200.hpp) exactly in the call order the reference's
// optimizers use (LinearizationHook: set.add(f) -> set.linearize(values) -> f->linearize(values) -> set.error(values) -> f->error(values)).
// Inputs and expected values come from a binary file written by tests/test_cpp_adapters.py (oracle results).
//   usage: test_adapters <case.bin>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "gtsam_points_b200/gtsam_points_b200.hpp"

using namespace gtsam_points_b200;

static std::vector<double> read_vec(FILE* f) {
  std::uint64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) throw std::runtime_error("short read");
  std::vector<double> v(n);
  if (n && fread(v.data(), 8, n, f) != n) throw std::runtime_error("short read");
  return v;
}

static double relerr(const double* a, const double* b, int n) {
  double num = 0, den = 0;
  for (int i = 0; i < n; i++) {
    num = std::fmax(num, std::fabs(a[i] - b[i]));
    den = std::fmax(den, std::fabs(b[i]));
  }
  return num / std::fmax(den, 1e-300);
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  const auto tp = read_vec(f), tc = read_vec(f), sp = read_vec(f), sc = read_vec(f);
  const auto Tt = read_vec(f), Ts = read_vec(f), Ts2 = read_vec(f);
  const auto exp_vgicp = read_vec(f), exp_gicp = read_vec(f), exp_err = read_vec(f);
  const auto res = read_vec(f);
  fclose(f);
  const std::size_t nt = tp.size() / 4, ns = sp.size() / 4;

  auto target = std::make_shared<PointCloudGPU>(tp.data(), tc.data(), nt);
  auto source = std::make_shared<PointCloudGPU>(sp.data(), sc.data(), ns);
  auto voxels = std::make_shared<GaussianVoxelMapGPU>(res[0]);
  voxels->insert(tp.data(), 4, tc.data(), 16, nt);
  auto tree = std::make_shared<KdTreeGPU>(tp.data(), 4, nt);

  auto vgicp = std::make_shared<IntegratedVGICPFactor>(0, 1, voxels, source);
  auto gicp = std::make_shared<IntegratedGICPFactor>(0, 1, target, source, tree);
  // the reference's tuning setters are accepted (source compatibility) and do not change results
  vgicp->set_num_threads(4);
  vgicp->set_fused_cov_cache_mode(FusedCovCacheMode::COMPACT);
  gicp->set_num_threads(4);
  gicp->set_fused_cov_cache_mode(FusedCovCacheMode::FULL);
  gicp->set_correspondence_update_tolerance(0.0, 0.0);
  if (vgicp->memory_usage() == 0 || gicp->memory_usage() == 0 || vgicp->get_target() != voxels) return 3;

  Values values, values2;
  Mat4 a, b, c;
  std::copy(Tt.begin(), Tt.end(), a.begin());
  std::copy(Ts.begin(), Ts.end(), b.begin());
  std::copy(Ts2.begin(), Ts2.end(), c.begin());
  values.insert(0, a);
  values.insert(1, b);
  values2.insert(0, a);
  values2.insert(1, c);

  NonlinearFactorSetGPU set;
  if (!set.add(vgicp) || !set.add(gicp)) return 3;
  auto lin = set.calc_linear_factors(values);
  if (set.linearization_count() != 2) return 4;
  const double tol = 1e-9;
  int fails = 0;
  const std::vector<double>* expd[2] = {&exp_vgicp, &exp_gicp};
  for (int k = 0; k < 2; k++) {
    const auto& e = *expd[k];
    fails += relerr(lin[k]->G11, e.data(), 36) > tol;
    fails += relerr(lin[k]->G22, e.data() + 36, 36) > tol;
    fails += relerr(lin[k]->G12, e.data() + 72, 36) > tol;
    double g1[6], g2[6];
    for (int i = 0; i < 6; i++) {
      g1[i] = -e[108 + i];
      g2[i] = -e[114 + i];
    }
    fails += relerr(lin[k]->g1, g1, 6) > tol;
    fails += relerr(lin[k]->g2, g2, 6) > tol;
    fails += std::fabs(lin[k]->f - e[120]) > tol * e[120];
  }
  if (vgicp->num_inliers() != static_cast<int>(exp_vgicp[121])) fails++;
  // error() after a batched evaluation returns the stored result (IntegratedVGICPFactorGPU::error semantics)
  set.error(values2);
  fails += std::fabs(vgicp->error(values2) - exp_err[0]) > tol * exp_err[0];
  fails += std::fabs(gicp->error(values2) - exp_err[1]) > tol * exp_err[1];
  // ... and without a set the factor evaluates on its own (sync path)
  fails += std::fabs(vgicp->error(values2) - exp_err[0]) > tol * exp_err[0];
  // single-query NearestNeighborSearch signature
  std::size_t idx;
  double sq;
  const double q[3] = {tp[0] + 0.01, tp[1], tp[2]};
  if (tree->knn_search(q, 1, &idx, &sq) != 1 || sq > 1e-3) fails++;
  printf("%s (%d failed checks, vgicp inliers %d)\n", fails ? "FAIL" : "OK", fails, vgicp->num_inliers());
  return fails ? 1 : 0;
}
