// b2_kdtree.cu -- kd-tree construction (host) and batched exact 1-NN queries (device).
//
// Replaces KdTree / KdTree2 behind NearestNeighborSearch::knn_search for k = 1
// (reference: include/gtsam_points/ann/nearest_neighbor_search.hpp:31-35, ann/kdtree2.hpp:26-61,
//  builders ann/small_kdtree.hpp:124-274).  The tree shape is our own (balanced median split on the axis of largest
// extent, <= 8 points per leaf, children adjacent, points re-ordered into leaf order as 16- / 32-byte records); since the search is exact the
// neighbours are the same as the reference's.
#include <algorithm>
#include <numeric>
#include <thread>
#include <vector>

#include "b2_kdtree.cuh"

namespace b2 {
namespace {

#ifndef B2_KD_LEAF
#define B2_KD_LEAF 16
#endif
constexpr int kMaxLeaf = B2_KD_LEAF;

struct Builder {
  const double* pts;
  int stride;
  std::vector<uint32_t>& order;  // permutation being partitioned; final = leaf order
  std::vector<KdNodeGPU>& nodes;

  double coord(uint32_t i, int axis) const { return pts[static_cast<size_t>(i) * stride + axis]; }

  // iterative build with an explicit work list; children of a node are allocated as an adjacent pair
  void build() {
    struct Work {
      uint32_t node, first, last;
    };
    nodes.clear();
    nodes.push_back(KdNodeGPU{0.0, 0u, 4u});
    std::vector<Work> work;
    work.push_back(Work{0u, 0u, static_cast<uint32_t>(order.size())});
    while (!work.empty()) {
      const Work w = work.back();
      work.pop_back();
      const uint32_t n = w.last - w.first;
      if (n <= static_cast<uint32_t>(kMaxLeaf)) {
        nodes[w.node] = KdNodeGPU{0.0, w.first, 4u + n};
        continue;
      }
      double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
      for (uint32_t k = w.first; k < w.last; k++) {
        for (int a = 0; a < 3; a++) {
          const double v = coord(order[k], a);
          mn[a] = std::min(mn[a], v);
          mx[a] = std::max(mx[a], v);
        }
      }
      int axis = 0;
      if (mx[1] - mn[1] > mx[axis] - mn[axis]) axis = 1;
      if (mx[2] - mn[2] > mx[axis] - mn[axis]) axis = 2;
      const uint32_t mid = w.first + n / 2;
      std::nth_element(order.begin() + w.first, order.begin() + mid, order.begin() + w.last,
                       [&](uint32_t i, uint32_t j) { return coord(i, axis) < coord(j, axis); });
      const uint32_t left = static_cast<uint32_t>(nodes.size());
      nodes.push_back(KdNodeGPU{0.0, 0u, 4u});
      nodes.push_back(KdNodeGPU{0.0, 0u, 4u});
      nodes[w.node] = KdNodeGPU{coord(order[mid], axis), left, static_cast<uint32_t>(axis)};
      work.push_back(Work{left + 1, mid, w.last});
      work.push_back(Work{left, w.first, mid});
    }
  }
};

__global__ void knn1_kernel(KdTreeView tree, const double* __restrict__ q, int qstride, size_t nq, double max_sq, const uint32_t* __restrict__ leaf_index,
                            long long* __restrict__ out_idx, double* __restrict__ out_sq) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const bool active = i < nq;
  const size_t k = active ? i : 0;
  double sq;
  const int j = kdtree_nn1_warp(tree, q[k * qstride], q[k * qstride + 1], q[k * qstride + 2], active, max_sq, &sq);
  if (!active) return;
  if (out_idx) out_idx[i] = j < 0 ? -1ll : static_cast<long long>(leaf_index[j]);
  if (out_sq) out_sq[i] = sq;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" {

b2_status b2_kdtree_create(b2_ctx* ctx, const double* points, int point_stride, size_t n, b2_kdtree** out) {
  B2_REQUIRE(out != nullptr, "b2_kdtree_create: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_kdtree_create: ctx is NULL");
  B2_REQUIRE(points != nullptr || n == 0, "b2_kdtree_create: points is NULL");
  B2_REQUIRE(point_stride == 3 || point_stride == 4, "b2_kdtree_create: point_stride must be 3 or 4");
  B2_REQUIRE(n < (1ull << 31), "b2_kdtree_create: at most 2^31-1 points");
  B2_CUDA(cudaSetDevice(ctx->device));

  b2_kdtree* t = new b2_kdtree;
  t->ctx = ctx;
  t->n = n;
  t->n_pad = round_up(std::max<size_t>(n, 1), 32);
  t->h_leaf_index.resize(n);
  std::iota(t->h_leaf_index.begin(), t->h_leaf_index.end(), 0u);
  std::vector<KdNodeGPU> nodes;
  Builder b{points, point_stride, t->h_leaf_index, nodes};
  b.build();
  t->num_nodes = nodes.size();

  // leaf-order point records: float32 when that is lossless (checked here), else float64
  bool f32 = true;
  for (size_t j = 0; j < n && f32; j++) {
    const double* p = points + j * point_stride;
    for (int a = 0; a < 3; a++) f32 = f32 && static_cast<double>(static_cast<float>(p[a])) == p[a];
  }
  t->leaf_f32 = f32;
  const size_t rec_bytes = f32 ? 4 * sizeof(float) : 4 * sizeof(double);
  std::vector<unsigned char> recs(std::max<size_t>(n, 1) * rec_bytes, 0);
  for (size_t j = 0; j < n; j++) {
    const double* p = points + static_cast<size_t>(t->h_leaf_index[j]) * point_stride;
    if (f32) {
      float* r = reinterpret_cast<float*>(recs.data()) + 4 * j;
      r[0] = static_cast<float>(p[0]), r[1] = static_cast<float>(p[1]), r[2] = static_cast<float>(p[2]);
    } else {
      double* r = reinterpret_cast<double*>(recs.data()) + 4 * j;
      r[0] = p[0], r[1] = p[1], r[2] = p[2];
    }
  }

  cudaStream_t st = ctx->stream;
  cudaError_t e;
  if ((e = cudaMalloc(reinterpret_cast<void**>(&t->d_nodes), nodes.size() * sizeof(KdNodeGPU))) != cudaSuccess ||
      (e = cudaMalloc(&t->d_leaf_points, recs.size())) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&t->d_leaf_index), std::max<size_t>(n, 1) * sizeof(uint32_t))) != cudaSuccess) {
    b2_kdtree_destroy(t);
    return fail(B2_ERR_OUT_OF_MEMORY, "b2_kdtree_create: %s", cudaGetErrorString(e));
  }
  t->device_bytes = nodes.size() * sizeof(KdNodeGPU) + recs.size() + n * sizeof(uint32_t);
  if ((e = cudaMemcpyAsync(t->d_nodes, nodes.data(), nodes.size() * sizeof(KdNodeGPU), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(t->d_leaf_points, recs.data(), recs.size(), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (n > 0 && (e = cudaMemcpyAsync(t->d_leaf_index, t->h_leaf_index.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) ||
      (e = cudaStreamSynchronize(st)) != cudaSuccess) {
    b2_kdtree_destroy(t);
    return fail(B2_ERR_CUDA, "b2_kdtree_create: %s", cudaGetErrorString(e));
  }
  *out = t;
  return B2_OK;
}

b2_status b2_kdtree_destroy(b2_kdtree* t) {
  if (!t) return B2_OK;
  cudaSetDevice(t->ctx->device);
  if (t->d_nodes) cudaFree(t->d_nodes);
  if (t->d_leaf_points) cudaFree(t->d_leaf_points);
  if (t->d_leaf_index) cudaFree(t->d_leaf_index);
  delete t;
  return B2_OK;
}

b2_status b2_kdtree_knn1(const b2_kdtree* t, const double* queries, int query_stride, size_t nq, double max_sq_dist, int64_t* out_index, double* out_sq_dist) {
  B2_REQUIRE(t != nullptr, "b2_kdtree_knn1: tree is NULL");
  B2_REQUIRE(nq == 0 || queries != nullptr, "b2_kdtree_knn1: queries is NULL");
  B2_REQUIRE(query_stride == 3 || query_stride == 4, "b2_kdtree_knn1: query_stride must be 3 or 4");
  B2_REQUIRE(max_sq_dist >= 0.0, "b2_kdtree_knn1: max_sq_dist must be >= 0");
  if (nq == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(t->ctx->device));
  cudaStream_t st = t->ctx->stream;
  DevBuf dq, di, ds;
  B2_CUDA(cudaMalloc(&dq.p, nq * query_stride * sizeof(double)));
  B2_CUDA(cudaMalloc(&di.p, nq * sizeof(long long)));
  B2_CUDA(cudaMalloc(&ds.p, nq * sizeof(double)));
  B2_CUDA(cudaMemcpyAsync(dq.p, queries, nq * query_stride * sizeof(double), cudaMemcpyHostToDevice, st));
  KdTreeView view{t->d_nodes, t->d_leaf_points, t->leaf_f32 ? 1 : 0};
  knn1_kernel<<<static_cast<unsigned>((nq + 127) / 128), 128, 0, st>>>(view, static_cast<const double*>(dq.p), query_stride, nq, max_sq_dist, t->d_leaf_index,
                                                                      static_cast<long long*>(di.p), static_cast<double*>(ds.p));
  B2_CUDA(cudaGetLastError());
  if (out_index) B2_CUDA(cudaMemcpyAsync(out_index, di.p, nq * sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (out_sq_dist) B2_CUDA(cudaMemcpyAsync(out_sq_dist, ds.p, nq * sizeof(double), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  return B2_OK;
}

}  // extern "C"
