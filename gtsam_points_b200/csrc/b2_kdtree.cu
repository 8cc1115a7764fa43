// b2_kdtree.cu -- kd-tree construction (device) and batched exact 1-NN queries (device).
//
// Replaces KdTree / KdTree2 behind NearestNeighborSearch::knn_search for k = 1
// (reference: include/gtsam_points/ann/nearest_neighbor_search.hpp:31-35, ann/kdtree2.hpp:26-61,
//  builders ann/small_kdtree.hpp:124-274).  The tree shape is our own (balanced: median split on the axis of largest
// extent, <= 16 points per leaf, children adjacent, points re-ordered into leaf order); since the search is exact the
// neighbours are the same as the reference's for any valid tree.
//
// Build, all on the device, one tree level per pass (~16 passes for 500k points), deterministic:
//   a node is a contiguous range of the point permutation `order`.  Per level: (1) every position finds its node (binary
//   search in the level's sorted range list) and folds its coordinates into the node's bounding box (atomic min / max on
//   order-preserving integer images of the doubles); (2) every node bigger than a leaf picks its axis of largest extent;
//   (3) every position of such a node emits the order-preserving image of its coordinate along that axis as a 64-bit key;
//   (4) ONE stable segmented radix sort (CUB) sorts all those nodes' ranges by key at once; (5) the node is split at the
//   median position: threshold = the median element's coordinate (exact data value), so every point of the lower child is
//   <= threshold and every point of the upper child >= threshold -- all the exact search needs; children are allocated as
//   an adjacent pair by an exclusive scan over the level (deterministic numbering).
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

// order-preserving map double -> uint64 (sort keys, atomicMin / atomicMax on coordinates)
__host__ __device__ __forceinline__ unsigned long long ordered_bits(double v) {
  unsigned long long b;
#ifdef __CUDA_ARCH__
  b = static_cast<unsigned long long>(__double_as_longlong(v));
#else
  std::memcpy(&b, &v, sizeof(b));
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

struct Range {
  uint32_t first, last;
};

// the range of the level that contains position `pos` (ranges are sorted by `first` and disjoint), or -1
__device__ __forceinline__ int find_range(const Range* __restrict__ level, uint32_t count, uint32_t pos) {
  uint32_t lo = 0u, hi = count;  // first index with level[idx].first > pos
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    if (level[mid].first <= pos)
      lo = mid + 1u;
    else
      hi = mid;
  }
  if (lo == 0u) return -1;
  const uint32_t k = lo - 1u;
  return pos < level[k].last ? static_cast<int>(k) : -1;
}

__global__ void scan_points_kernel(const double* __restrict__ pts, int stride, size_t n, unsigned int* __restrict__ not_f32, unsigned int* __restrict__ not_finite,
                                   uint32_t* __restrict__ order) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  bool lossy = false, bad = false;
  if (i < n) {
    order[i] = static_cast<uint32_t>(i);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double v = pts[i * stride + a];
      lossy |= static_cast<double>(static_cast<float>(v)) != v;
      bad |= !isfinite(v);
    }
  }
  if (__any_sync(0xffffffffu, lossy) && (threadIdx.x & 31) == 0) atomicOr(not_f32, 1u);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(not_finite, 1u);
}

__global__ void classify_level_kernel(const Range* __restrict__ level, uint32_t count, uint32_t* __restrict__ flag) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) flag[k] = (level[k].last - level[k].first > static_cast<uint32_t>(kMaxLeaf)) ? 1u : 0u;
}

__global__ void init_bbox_kernel(unsigned long long* __restrict__ bbox, uint32_t internal) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < internal * 6u) bbox[k] = (k % 6u) < 3u ? ~0ull : 0ull;  // min x y z | max x y z
}

__global__ void bbox_level_kernel(const double* __restrict__ pts, int stride, size_t n, const uint32_t* __restrict__ order, const Range* __restrict__ level,
                                  uint32_t count, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank, unsigned long long* __restrict__ bbox) {
  const size_t pos = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (pos >= n) return;
  const int k = find_range(level, count, static_cast<uint32_t>(pos));
  if (k < 0 || !flag[k]) return;
  const double* p = pts + static_cast<size_t>(order[pos]) * stride;
  unsigned long long* b = bbox + static_cast<size_t>(rank[k]) * 6;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const unsigned long long o = ordered_bits(p[a]);
    atomicMin(&b[a], o);
    atomicMax(&b[3 + a], o);
  }
}

__device__ __forceinline__ double from_ordered_bits_dev(unsigned long long o) {
  const unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double(static_cast<long long>(b));
}

// per internal node (indexed by rank): axis of largest extent + the segment it occupies in `order`
__global__ void axis_level_kernel(const Range* __restrict__ level, uint32_t count, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank,
                                  const unsigned long long* __restrict__ bbox, int* __restrict__ axis, uint32_t* __restrict__ seg_begin, uint32_t* __restrict__ seg_end) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count || !flag[k]) return;
  const uint32_t r = rank[k];
  const unsigned long long* b = bbox + static_cast<size_t>(r) * 6;
  double ext[3];
#pragma unroll
  for (int a = 0; a < 3; a++) ext[a] = from_ordered_bits_dev(b[3 + a]) - from_ordered_bits_dev(b[a]);
  int ax = 0;
  if (ext[1] > ext[ax]) ax = 1;
  if (ext[2] > ext[ax]) ax = 2;
  axis[r] = ax;
  seg_begin[r] = level[k].first;
  seg_end[r] = level[k].last;
}

__global__ void keys_level_kernel(const double* __restrict__ pts, int stride, size_t n, const uint32_t* __restrict__ order, const Range* __restrict__ level,
                                  uint32_t count, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank, const int* __restrict__ axis,
                                  unsigned long long* __restrict__ keys) {
  const size_t pos = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (pos >= n) return;
  const int k = find_range(level, count, static_cast<uint32_t>(pos));
  keys[pos] = (k >= 0 && flag[k]) ? ordered_bits(pts[static_cast<size_t>(order[pos]) * stride + axis[rank[k]]]) : 0ull;
}

__global__ void emit_level_kernel(const double* __restrict__ pts, int stride, const uint32_t* __restrict__ order_sorted, const Range* __restrict__ level, uint32_t count,
                                  uint32_t level_base, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank, const int* __restrict__ axis,
                                  uint32_t child_base, KdNodeGPU* __restrict__ nodes, Range* __restrict__ next) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const Range r = level[k];
  KdNodeGPU nd;
  if (!flag[k]) {
    nd.thresh = 0.0;
    nd.a = r.first;
    nd.b = 4u + (r.last - r.first);
  } else {
    const uint32_t rk = rank[k];
    const uint32_t mid = r.first + (r.last - r.first) / 2u;
    const int ax = axis[rk];
    nd.thresh = pts[static_cast<size_t>(order_sorted[mid]) * stride + ax];  // lower child <= thresh <= upper child
    nd.a = child_base + 2u * rk;
    nd.b = static_cast<uint32_t>(ax);
    next[2u * rk] = Range{r.first, mid};
    next[2u * rk + 1u] = Range{mid, r.last};
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
  DevBuf d_pts, d_flags2, d_orderB, d_keysA, d_keysB, d_tmp, d_levelA, d_levelB, d_flag, d_rank, d_bbox, d_axis, d_segb, d_sege, d_scan_tmp;
  KD_CUDA(cudaMalloc(&d_pts.p, n * point_stride * sizeof(double)));
  KD_CUDA(cudaMemcpyAsync(d_pts.p, points, n * point_stride * sizeof(double), cudaMemcpyHostToDevice, st));
  const double* dp = static_cast<const double*>(d_pts.p);

  // point scan: identity permutation, "is every coordinate float32-representable", "is every coordinate finite"
  KD_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_leaf_index), n * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_orderB.p, n * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_flags2.p, 2 * sizeof(unsigned int)));
  KD_CUDA(cudaMemsetAsync(d_flags2.p, 0, 2 * sizeof(unsigned int), st));
  uint32_t* order = t->d_leaf_index;                     // the two permutation buffers swap roles level by level;
  uint32_t* order_alt = static_cast<uint32_t*>(d_orderB.p);  // the final one is copied into t->d_leaf_index if needed
  scan_points_kernel<<<grid_n, 256, 0, st>>>(dp, point_stride, n, static_cast<unsigned int*>(d_flags2.p), static_cast<unsigned int*>(d_flags2.p) + 1, order);
  KD_CUDA(cudaGetLastError());
  unsigned int h_flags2[2] = {0u, 0u};
  KD_CUDA(cudaMemcpyAsync(h_flags2, d_flags2.p, sizeof(h_flags2), cudaMemcpyDeviceToHost, st));
  KD_CUDA(cudaStreamSynchronize(st));
  if (h_flags2[1]) return bail(fail(B2_ERR_INVALID_ARGUMENT, "b2_kdtree_create: non-finite coordinate"));
  t->leaf_f32 = h_flags2[0] == 0u;

  const size_t max_nodes = 2 * n + 1;
  const size_t max_level = n / 2 + 2;  // nodes per level (every node of a level below the root holds >= kMaxLeaf / 2 >= 2 points)
  KD_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->d_nodes), max_nodes * sizeof(KdNodeGPU)));
  KD_CUDA(cudaMalloc(&d_keysA.p, n * sizeof(unsigned long long)));
  KD_CUDA(cudaMalloc(&d_keysB.p, n * sizeof(unsigned long long)));
  KD_CUDA(cudaMalloc(&d_levelA.p, max_level * sizeof(Range)));
  KD_CUDA(cudaMalloc(&d_levelB.p, max_level * sizeof(Range)));
  KD_CUDA(cudaMalloc(&d_flag.p, max_level * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_rank.p, max_level * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_bbox.p, max_level * 6 * sizeof(unsigned long long)));
  KD_CUDA(cudaMalloc(&d_axis.p, max_level * sizeof(int)));
  KD_CUDA(cudaMalloc(&d_segb.p, max_level * sizeof(uint32_t)));
  KD_CUDA(cudaMalloc(&d_sege.p, max_level * sizeof(uint32_t)));
  size_t scan_bytes = 0;
  KD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<const uint32_t*>(d_flag.p), static_cast<uint32_t*>(d_rank.p), static_cast<int>(max_level), st));
  KD_CUDA(cudaMalloc(&d_scan_tmp.p, std::max<size_t>(scan_bytes, 16)));
  size_t sort_bytes = 0;
  const Range root{0u, static_cast<uint32_t>(n)};
  KD_CUDA(cudaMemcpyAsync(d_levelA.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
  Range* cur = static_cast<Range*>(d_levelA.p);
  Range* nxt = static_cast<Range*>(d_levelB.p);
  uint32_t* flag = static_cast<uint32_t*>(d_flag.p);
  uint32_t* rank = static_cast<uint32_t*>(d_rank.p);
  uint32_t count = 1u, level_base = 0u, total = 1u;
  int depth = 0;
  while (count > 0u) {
    if (++depth > kKdStackDepth) return bail(fail(B2_ERR_INVALID_STATE, "b2_kdtree_create: tree deeper than the traversal stack (%d levels)", kKdStackDepth));
    const unsigned gl = (count + 255u) / 256u;
    classify_level_kernel<<<gl, 256, 0, st>>>(cur, count, flag);
    KD_CUDA(cudaGetLastError());
    KD_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, scan_bytes, flag, rank, static_cast<int>(count), st));
    uint32_t last_rank = 0u, last_flag = 0u;  // number of nodes to split on this level = rank[count-1] + flag[count-1]
    KD_CUDA(cudaMemcpyAsync(&last_rank, rank + (count - 1u), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    KD_CUDA(cudaMemcpyAsync(&last_flag, flag + (count - 1u), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    KD_CUDA(cudaStreamSynchronize(st));
    const uint32_t internal = last_rank + last_flag;
    if (internal > 0u) {
      init_bbox_kernel<<<(internal * 6u + 255u) / 256u, 256, 0, st>>>(static_cast<unsigned long long*>(d_bbox.p), internal);
      bbox_level_kernel<<<grid_n, 256, 0, st>>>(dp, point_stride, n, order, cur, count, flag, rank, static_cast<unsigned long long*>(d_bbox.p));
      axis_level_kernel<<<gl, 256, 0, st>>>(cur, count, flag, rank, static_cast<const unsigned long long*>(d_bbox.p), static_cast<int*>(d_axis.p),
                                            static_cast<uint32_t*>(d_segb.p), static_cast<uint32_t*>(d_sege.p));
      keys_level_kernel<<<grid_n, 256, 0, st>>>(dp, point_stride, n, order, cur, count, flag, rank, static_cast<const int*>(d_axis.p),
                                                static_cast<unsigned long long*>(d_keysA.p));
      KD_CUDA(cudaGetLastError());
      // positions outside the sorted segments (finished leaves) keep their element
      KD_CUDA(cudaMemcpyAsync(order_alt, order, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
      size_t need = 0;
      KD_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr, need, static_cast<const unsigned long long*>(d_keysA.p), static_cast<unsigned long long*>(d_keysB.p),
                                                        static_cast<const uint32_t*>(order), order_alt, static_cast<int>(n), static_cast<int>(internal),
                                                        static_cast<const uint32_t*>(d_segb.p), static_cast<const uint32_t*>(d_sege.p), st));
      if (need > sort_bytes) {
        KD_CUDA(cudaStreamSynchronize(st));
        if (d_tmp.p) cudaFree(d_tmp.p);
        d_tmp.p = nullptr;
        KD_CUDA(cudaMalloc(&d_tmp.p, need));
        sort_bytes = need;
      }
      KD_CUDA(cub::DeviceSegmentedSort::StableSortPairs(d_tmp.p, need, static_cast<const unsigned long long*>(d_keysA.p), static_cast<unsigned long long*>(d_keysB.p),
                                                        static_cast<const uint32_t*>(order), order_alt, static_cast<int>(n), static_cast<int>(internal),
                                                        static_cast<const uint32_t*>(d_segb.p), static_cast<const uint32_t*>(d_sege.p), st));
      std::swap(order, order_alt);
    }
    emit_level_kernel<<<gl, 256, 0, st>>>(dp, point_stride, order, cur, count, level_base, flag, rank, static_cast<const int*>(d_axis.p), total, t->d_nodes, nxt);
    KD_CUDA(cudaGetLastError());
    level_base = total;
    total += 2u * internal;
    count = 2u * internal;
    std::swap(cur, nxt);
  }
  t->num_nodes = total;
  if (order != t->d_leaf_index) KD_CUDA(cudaMemcpyAsync(t->d_leaf_index, order, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));

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
