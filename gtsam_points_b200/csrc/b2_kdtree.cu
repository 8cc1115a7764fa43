// b2_kdtree.cu -- kd-tree construction (device) and batched exact 1-NN queries (device).
//
// Replaces KdTree / KdTree2 behind NearestNeighborSearch::knn_search for k = 1
// (reference: include/gtsam_points/ann/nearest_neighbor_search.hpp:31-35, ann/kdtree2.hpp:26-61,
//  builders ann/small_kdtree.hpp:124-274).  The tree shape is our own; since the search is exact the neighbours are the
// same as the reference's for any valid tree.
//
// Build (all on the device, deterministic, ~1 ms for 500k points instead of 0.36 s on the host):
//   1. bounding box; every coordinate is quantised to 16 bits per axis against the threshold grid
//      T_a(Q) = min_a + Q * cell_a, with the quantum CORRECTED so that T_a(q) <= x < T_a(q + 1) holds exactly in floating point;
//   2. 48-bit Morton keys (x, y, z bits interleaved, x most significant), stable radix sort of (key, index): the sorted
//      order is the leaf order;
//   3. top-down, one level per pass: a node is a range of the sorted keys; it becomes a leaf if it holds <= 16 points or
//      all its keys are equal, else it is split at the most significant key bit in which its first and last key differ --
//      the children are the sub-ranges with that bit clear / set (binary search), allocated as an adjacent pair by an
//      exclusive scan over the level (deterministic numbering), and the split plane is x_axis = T_axis(Q) with Q the
//      quantum prefix of the upper half: by (1) every point of the lower child is strictly below it and every point of the
//      upper child is on or above it, which is all the exact search needs.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>

#include "b2_kdtree.cuh"

namespace b2 {
namespace {

#ifndef B2_KD_LEAF
#define B2_KD_LEAF 16
#endif
constexpr int kMaxLeaf = B2_KD_LEAF;
constexpr int kQuantBits = 16;
constexpr uint32_t kQuantMax = (1u << kQuantBits) - 1u;

struct Grid {
  double mn[3];
  double cell[3];
  double inv_cell[3];
};

__device__ __forceinline__ double grid_threshold(const Grid& g, int axis, uint32_t Q) { return __dadd_rn(g.mn[axis], __dmul_rn(static_cast<double>(Q), g.cell[axis])); }

// order-preserving map double -> uint64 (for atomicMin / atomicMax on coordinates)
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
inline double from_ordered_bits(unsigned long long o) {
  const unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
  double v;
  std::memcpy(&v, &b, sizeof(v));
  return v;
}

__global__ void bbox_kernel(const double* __restrict__ pts, int stride, size_t n, unsigned long long* __restrict__ mnmx /* min xyz | max xyz, ordered bits */,
                            unsigned int* __restrict__ not_f32) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
  bool lossy = false;
  if (i < n) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double v = pts[i * stride + a];
      lo[a] = hi[a] = ordered_bits(v);
      lossy |= static_cast<double>(static_cast<float>(v)) != v;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
  }
  lossy = __any_sync(0xffffffffu, lossy);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(&mnmx[a], lo[a]);
      atomicMax(&mnmx[3 + a], hi[a]);
    }
    if (lossy) atomicOr(not_f32, 1u);
  }
}

__device__ __forceinline__ unsigned long long spread3(uint32_t v) {  // 16 bits -> every third bit
  unsigned long long x = v & 0xffffull;
  x = (x | (x << 32)) & 0x001f00000000ffffull;
  x = (x | (x << 16)) & 0x001f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

__global__ void morton_kernel(const double* __restrict__ pts, int stride, size_t n, Grid g, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t q[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const double v = pts[i * stride + a];
    double f = floor((v - g.mn[a]) * g.inv_cell[a]);
    f = f < 0.0 ? 0.0 : (f > static_cast<double>(kQuantMax) ? static_cast<double>(kQuantMax) : f);
    uint32_t Q = static_cast<uint32_t>(f);
    // make the quantum consistent with the threshold grid: T(Q) <= v < T(Q + 1), exactly
    while (Q > 0u && v < grid_threshold(g, a, Q)) Q--;
    while (Q < kQuantMax && v >= grid_threshold(g, a, Q + 1u)) Q++;
    q[a] = Q;
  }
  keys[i] = (spread3(q[0]) << 2) | (spread3(q[1]) << 1) | spread3(q[2]);
  idx[i] = static_cast<uint32_t>(i);
}

struct Range {
  uint32_t first, last;
};

// One level of the top-down build: decide leaf / internal for every node of the level, find the split of internal nodes.
__global__ void classify_level_kernel(const unsigned long long* __restrict__ keys, const Range* __restrict__ level, uint32_t count, uint32_t* __restrict__ flag,
                                      uint32_t* __restrict__ split, int* __restrict__ bit) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const Range r = level[k];
  const unsigned long long kf = keys[r.first], kl = keys[r.last - 1u];
  if (r.last - r.first <= static_cast<uint32_t>(kMaxLeaf) || kf == kl) {
    flag[k] = 0u;
    return;
  }
  const int b = 63 - __clzll(static_cast<long long>(kf ^ kl));  // most significant differing bit; higher bits are common to the range
  const unsigned long long pivot = ((kl >> b) << b);            // smallest key of the upper half
  uint32_t lo = r.first, hi = r.last - 1u;                      // keys[lo] < pivot <= keys[hi]
  while (hi - lo > 1u) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    if (keys[mid] < pivot)
      lo = mid;
    else
      hi = mid;
  }
  flag[k] = 1u;
  split[k] = hi;
  bit[k] = b;
}

__global__ void emit_level_kernel(const unsigned long long* __restrict__ keys, const Range* __restrict__ level, uint32_t count, uint32_t level_base,
                                  const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ split, const int* __restrict__ bit,
                                  uint32_t child_base, Grid g, KdNodeGPU* __restrict__ nodes, Range* __restrict__ next) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const Range r = level[k];
  KdNodeGPU nd;
  if (!flag[k]) {
    nd.thresh = 0.0;
    nd.a = r.first;
    nd.b = 4u + (r.last - r.first);
  } else {
    const int b = bit[k];
    const int axis = 2 - (b % 3);   // key bit 3c+2 is x, 3c+1 is y, 3c is z
    const int cb = b / 3;           // coordinate bit
    // quantum prefix of the upper half along `axis`: the coordinate bits above and including cb of its smallest key
    const unsigned long long ks = keys[split[k]];
    uint32_t q = 0u;
    for (int c = kQuantBits - 1; c >= cb; c--) q |= static_cast<uint32_t>((ks >> (3 * c + (2 - axis))) & 1ull) << c;
    nd.thresh = grid_threshold(g, axis, q);
    nd.a = child_base + 2u * rank[k];
    nd.b = static_cast<uint32_t>(axis);
    next[2u * rank[k]] = Range{r.first, split[k]};
    next[2u * rank[k] + 1u] = Range{split[k], r.last};
  }
  nodes[level_base + k] = nd;
}

__global__ void leaf_records_kernel(const double* __restrict__ pts, int stride, size_t n, const uint32_t* __restrict__ order, int f32, void* __restrict__ recs) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  const double* p = pts + static_cast<size_t>(order[j]) * stride;
  if (f32) {
    static_cast<float4*>(recs)[j] = make_float4(static_cast<float>(p[0]), static_cast<float>(p[1]), static_cast<float>(p[2]), 0.0f);
  } else {
    double2* r = static_cast<double2*>(recs) + 2 * j;
    r[0] = make_double2(p[0], p[1]);
    r[1] = make_double2(p[2], 0.0);
  }
}

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
  cudaStream_t st = ctx->stream;

  b2_kdtree* t = new b2_kdtree;
  t->ctx = ctx;
  t->n = n;
  t->n_pad = round_up(std::max<size_t>(n, 1), 32);
  auto bail = [&](b2_status stt) {
    b2_kdtree_destroy(t);
    return stt;
  };
#define KD_CUDA(expr)                                                                                                     \
  do {                                                                                                                    \
    cudaError_t _e = (expr);                                                                                              \
    if (_e != cudaSuccess) return bail(fail(_e == cudaErrorMemoryAllocation ? B2_ERR_OUT_OF_MEMORY : B2_ERR_CUDA, "b2_kdtree_create: %s -> %s", #expr, cudaGetErrorString(_e))); \
  } while (0)

  if (n == 0) {  // an empty tree: one empty leaf
    const KdNodeGPU root{0.0, 0u, 4u};
    KD_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_nodes), sizeof(KdNodeGPU)));
    KD_CUDA(cudaMalloc(&t->d_leaf_points, 32));
    KD_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_leaf_index), sizeof(uint32_t)));
    KD_CUDA(cudaMemcpyAsync(t->d_nodes, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    KD_CUDA(cudaStreamSynchronize(st));
    t->num_nodes = 1;
    t->leaf_f32 = true;
    *out = t;
    return B2_OK;
  }

  const unsigned grid_n = static_cast<unsigned>((n + 255) / 256);
  DevBuf d_pts, d_mnmx, d_flag32, d_keys, d_keys_sorted, d_idx, d_tmp, d_levelA, d_levelB, d_flag, d_rank, d_split, d_bit, d_scan_tmp;
  KD_CUDA(cudaMalloc(&d_pts.p, n * point_stride * sizeof(double)));
  KD_CUDA(cudaMemcpyAsync(d_pts.p, points, n * point_stride * sizeof(double), cudaMemcpyHostToDevice, st));
  const double* dp = static_cast<const double*>(d_pts.p);

  // 1. bounding box + "is every coordinate float32-representable"
  unsigned long long h_mnmx[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  unsigned int h_not_f32 = 0u;
  KD_CUDA(cudaMalloc(&d_mnmx.p, sizeof(h_mnmx)));
  KD_CUDA(cudaMalloc(&d_flag32.p, sizeof(unsigned int)));
  KD_CUDA(cudaMemcpyAsync(d_mnmx.p, h_mnmx, sizeof(h_mnmx), cudaMemcpyHostToDevice, st));
  KD_CUDA(cudaMemsetAsync(d_flag32.p, 0, sizeof(unsigned int), st));
  bbox_kernel<<<grid_n, 256, 0, st>>>(dp, point_stride, n, static_cast<unsigned long long*>(d_mnmx.p), static_cast<unsigned int*>(d_flag32.p));
  KD_CUDA(cudaGetLastError());
  KD_CUDA(cudaMemcpyAsync(h_mnmx, d_mnmx.p, sizeof(h_mnmx), cudaMemcpyDeviceToHost, st));
  KD_CUDA(cudaMemcpyAsync(&h_not_f32, d_flag32.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  KD_CUDA(cudaStreamSynchronize(st));
  t->leaf_f32 = h_not_f32 == 0u;
  Grid g;
  for (int a = 0; a < 3; a++) {
    const double mn = from_ordered_bits(h_mnmx[a]), mx = from_ordered_bits(h_mnmx[3 + a]);
    if (!(std::isfinite(mn) && std::isfinite(mx))) return bail(fail(B2_ERR_INVALID_ARGUMENT, "b2_kdtree_create: non-finite coordinate"));
    g.mn[a] = mn;
    g.cell[a] = mx > mn ? (mx - mn) / static_cast<double>(1u << kQuantBits) : 1.0;
    g.inv_cell[a] = 1.0 / g.cell[a];
  }

  // 2. Morton keys, stable sort: the sorted order is the leaf order
  KD_CUDA(cudaMalloc(&d_keys.p, n * sizeof(unsigned long long)));
  KD_CUDA(cudaMalloc(&d_keys_sorted.p, n * sizeof(unsigned long long)));
  KD_CUDA(cudaMalloc(&d_idx.p, n * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_leaf_index), n * sizeof(uint32_t)));
  morton_kernel<<<grid_n, 256, 0, st>>>(dp, point_stride, n, g, static_cast<unsigned long long*>(d_keys.p), static_cast<uint32_t*>(d_idx.p));
  KD_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  KD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, static_cast<const unsigned long long*>(d_keys.p), static_cast<unsigned long long*>(d_keys_sorted.p),
                                          static_cast<const uint32_t*>(d_idx.p), t->d_leaf_index, static_cast<int>(n), 0, 3 * kQuantBits, st));
  KD_CUDA(cudaMalloc(&d_tmp.p, std::max<size_t>(tmp_bytes, 16)));
  KD_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, static_cast<const unsigned long long*>(d_keys.p), static_cast<unsigned long long*>(d_keys_sorted.p),
                                          static_cast<const uint32_t*>(d_idx.p), t->d_leaf_index, static_cast<int>(n), 0, 3 * kQuantBits, st));
  const unsigned long long* keys = static_cast<const unsigned long long*>(d_keys_sorted.p);

  // 3. top-down build, one level per pass (at most 3 * kQuantBits + 1 levels)
  const size_t max_nodes = 2 * n + 1;
  KD_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_nodes), max_nodes * sizeof(KdNodeGPU)));
  KD_CUDA(cudaMalloc(&d_levelA.p, (n + 1) * sizeof(Range)));
  KD_CUDA(cudaMalloc(&d_levelB.p, (n + 1) * sizeof(Range)));
  KD_CUDA(cudaMalloc(&d_flag.p, (n + 1) * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_rank.p, (n + 1) * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_split.p, (n + 1) * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_bit.p, (n + 1) * sizeof(int)));
  size_t scan_bytes = 0;
  KD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<const uint32_t*>(d_flag.p), static_cast<uint32_t*>(d_rank.p), static_cast<int>(n + 1), st));
  KD_CUDA(cudaMalloc(&d_scan_tmp.p, std::max<size_t>(scan_bytes, 16)));
  const Range root{0u, static_cast<uint32_t>(n)};
  KD_CUDA(cudaMemcpyAsync(d_levelA.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
  Range* cur = static_cast<Range*>(d_levelA.p);
  Range* nxt = static_cast<Range*>(d_levelB.p);
  uint32_t count = 1u, level_base = 0u, total = 1u;
  int depth = 0;
  while (count > 0u) {
    if (++depth > kKdStackDepth) return bail(fail(B2_ERR_INVALID_STATE, "b2_kdtree_create: tree deeper than the traversal stack (%d levels)", kKdStackDepth));
    const unsigned gl = (count + 255u) / 256u;
    classify_level_kernel<<<gl, 256, 0, st>>>(keys, cur, count, static_cast<uint32_t*>(d_flag.p), static_cast<uint32_t*>(d_split.p), static_cast<int*>(d_bit.p));
    KD_CUDA(cudaGetLastError());
    KD_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, scan_bytes, static_cast<const uint32_t*>(d_flag.p), static_cast<uint32_t*>(d_rank.p), static_cast<int>(count), st));
    emit_level_kernel<<<gl, 256, 0, st>>>(keys, cur, count, level_base, static_cast<const uint32_t*>(d_flag.p), static_cast<const uint32_t*>(d_rank.p),
                                          static_cast<const uint32_t*>(d_split.p), static_cast<const int*>(d_bit.p), total, g, t->d_nodes, nxt);
    KD_CUDA(cudaGetLastError());
    uint32_t last_rank = 0u, last_flag = 0u;  // number of internal nodes of this level = rank[count-1] + flag[count-1]
    KD_CUDA(cudaMemcpyAsync(&last_rank, static_cast<uint32_t*>(d_rank.p) + (count - 1u), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    KD_CUDA(cudaMemcpyAsync(&last_flag, static_cast<uint32_t*>(d_flag.p) + (count - 1u), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    KD_CUDA(cudaStreamSynchronize(st));
    const uint32_t internal = last_rank + last_flag;
    level_base = total;
    total += 2u * internal;
    count = 2u * internal;
    std::swap(cur, nxt);
  }
  t->num_nodes = total;

  // leaf-order point records: float32 when that is lossless, else float64
  const size_t rec_bytes = t->leaf_f32 ? 4 * sizeof(float) : 4 * sizeof(double);
  KD_CUDA(cudaMalloc(&t->d_leaf_points, n * rec_bytes));
  leaf_records_kernel<<<grid_n, 256, 0, st>>>(dp, point_stride, n, t->d_leaf_index, t->leaf_f32 ? 1 : 0, t->d_leaf_points);
  KD_CUDA(cudaGetLastError());
  KD_CUDA(cudaStreamSynchronize(st));
  t->device_bytes = max_nodes * sizeof(KdNodeGPU) + n * rec_bytes + n * sizeof(uint32_t);
#undef KD_CUDA
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
