// b2_voxelmap.cu -- Gaussian voxel map on the device.
//
// Replaces GaussianVoxelMapGPU (reference: include/gtsam_points/types/gaussian_voxelmap_gpu.hpp:39-108,
// src/gtsam_points/types/gaussian_voxelmap_gpu.cu:178-307) while keeping the INDEX SEMANTICS of GaussianVoxelMapCPU
// (src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:23-77, ann/impl/incremental_voxelmap_impl.hpp:31-68):
//   * voxel id = first-touch order of the voxel in the inserted cloud (the reference GPU builder hands out ids from an
//     atomicAdd race, gaussian_voxelmap_gpu.cu:57-67);
//   * no point is dropped (the reference GPU builder tolerates target_points_drop_rate = 1e-3);
//   * voxel mean / covariance are float64 sums taken in point order, divided by the count -- bit-identical to the CPU map;
//   * lookups probe until an empty bucket: a voxel that exists is always found (no max_bucket_scan_count).
// Build = two stable LSD radix sorts on the integer voxel coordinate (segments come out with their points in insertion
// order), a scan over segment heads, one more sort of the segments by first-touch index, then one thread per voxel
// accumulates its points sequentially.  Everything is deterministic.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstring>
#include <vector>

#include "b2_device.cuh"

namespace b2 {
namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  template <typename T>
  T* as() {
    return static_cast<T*>(p);
  }
};

__device__ __forceinline__ uint32_t bias32(int v) { return static_cast<uint32_t>(v) ^ 0x80000000u; }

__global__ void point_coord_keys_kernel(const double* __restrict__ pts, int pstride, size_t n, double inv_leaf, uint32_t* __restrict__ key_z,
                                        uint32_t* __restrict__ idx) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  key_z[i] = bias32(voxel_coord1(pts[i * pstride + 2], inv_leaf));
  idx[i] = static_cast<uint32_t>(i);
}

__global__ void gather_xy_keys_kernel(const double* __restrict__ pts, int pstride, size_t n, double inv_leaf, const uint32_t* __restrict__ idx,
                                      unsigned long long* __restrict__ key_xy) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  const size_t i = idx[j];
  const uint32_t x = bias32(voxel_coord1(pts[i * pstride + 0], inv_leaf));
  const uint32_t y = bias32(voxel_coord1(pts[i * pstride + 1], inv_leaf));
  key_xy[j] = (static_cast<unsigned long long>(x) << 32) | y;
}

__global__ void head_flags_kernel(const double* __restrict__ pts, int pstride, size_t n, double inv_leaf, const uint32_t* __restrict__ idx,
                                  uint32_t* __restrict__ flags) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  if (j == 0) {
    flags[0] = 1;
    return;
  }
  const size_t a = idx[j], b = idx[j - 1];
  bool differ = false;
#pragma unroll
  for (int k = 0; k < 3; k++) differ |= voxel_coord1(pts[a * pstride + k], inv_leaf) != voxel_coord1(pts[b * pstride + k], inv_leaf);
  flags[j] = differ ? 1u : 0u;
}

// seg_of[j] = inclusive scan of flags - 1.  Writes, for every segment, its start position and first-touch point index.
__global__ void segment_heads_kernel(size_t n, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ seg_incl, const uint32_t* __restrict__ idx,
                                     uint32_t* __restrict__ seg_start, uint32_t* __restrict__ seg_first, uint32_t* __restrict__ seg_iota) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  if (flags[j]) {
    const uint32_t s = seg_incl[j] - 1;
    seg_start[s] = static_cast<uint32_t>(j);
    seg_first[s] = idx[j];  // stable sorts keep insertion order inside a segment => its head is the first-touch point
    seg_iota[s] = s;
  }
}

// One thread per voxel (id order): sequential float64 accumulation in insertion order, then finalize (divide by count).
__global__ void accumulate_voxels_kernel(const double* __restrict__ pts, int pstride, const double* __restrict__ covs, int cstride, size_t n, size_t num_voxels,
                                         double inv_leaf, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ seg_start,
                                         const uint32_t* __restrict__ seg_of_rank, double* __restrict__ records, int32_t* __restrict__ coords) {
  const size_t r = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (r >= num_voxels) return;
  const uint32_t s = seg_of_rank[r];
  const size_t begin = seg_start[s];
  const size_t end = (s + 1 < num_voxels) ? seg_start[s + 1] : n;
  const int ld = cstride == 16 ? 4 : 3;
  double m[3] = {0, 0, 0}, c[6] = {0, 0, 0, 0, 0, 0};
  for (size_t j = begin; j < end; j++) {
    const size_t i = idx[j];
    const double* p = pts + i * pstride;
    const double* cv = covs + i * cstride;
    m[0] = __dadd_rn(m[0], p[0]);
    m[1] = __dadd_rn(m[1], p[1]);
    m[2] = __dadd_rn(m[2], p[2]);
    c[0] = __dadd_rn(c[0], cv[0 * ld + 0]);
    c[1] = __dadd_rn(c[1], cv[0 * ld + 1]);
    c[2] = __dadd_rn(c[2], cv[0 * ld + 2]);
    c[3] = __dadd_rn(c[3], cv[1 * ld + 1]);
    c[4] = __dadd_rn(c[4], cv[1 * ld + 2]);
    c[5] = __dadd_rn(c[5], cv[2 * ld + 2]);
  }
  const double cnt = static_cast<double>(end - begin);
  double* rec = records + r * kRecordDoubles;
#pragma unroll
  for (int k = 0; k < 3; k++) rec[k] = __ddiv_rn(m[k], cnt);
#pragma unroll
  for (int k = 0; k < 6; k++) rec[3 + k] = __ddiv_rn(c[k], cnt);
  rec[9] = cnt;
  const double* p0 = pts + static_cast<size_t>(idx[begin]) * pstride;
#pragma unroll
  for (int k = 0; k < 3; k++) coords[r * 3 + k] = voxel_coord1(p0[k], inv_leaf);
}

__global__ void insert_buckets_kernel(const int32_t* __restrict__ coords, size_t num_voxels, VoxelBucket* __restrict__ buckets, uint32_t mask) {
  const size_t r = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (r >= num_voxels) return;
  const int x = coords[r * 3], y = coords[r * 3 + 1], z = coords[r * 3 + 2];
  uint32_t g = voxel_hash(x, y, z) & mask;  // mask = number of groups - 1
  while (true) {
    // all keys are distinct: claim the first empty slot of the first non-full group; nobody compares against a half-written key
    for (int k = 0; k < kGroup; k++) {
      VoxelBucket* b = buckets + static_cast<size_t>(g) * kGroup + k;
      if (atomicCAS(&b->id, -1, static_cast<int>(r)) == -1) {
        b->x = x;
        b->y = y;
        b->z = z;
        return;
      }
    }
    g = (g + 1) & mask;
  }
}

__global__ void lookup_points_kernel(const double* __restrict__ pts, int pstride, size_t n, double inv_leaf, const VoxelBucket* __restrict__ buckets,
                                     uint32_t mask, int32_t* __restrict__ out) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int x = voxel_coord1(pts[i * pstride + 0], inv_leaf);
  const int y = voxel_coord1(pts[i * pstride + 1], inv_leaf);
  const int z = voxel_coord1(pts[i * pstride + 2], inv_leaf);
  out[i] = lookup_voxel(buckets, mask, x, y, z);
}

// load factor <= 0.25 (see b2_device.cuh: one-round-trip lookups)
size_t bucket_count_for(size_t num_voxels) { return static_cast<size_t>(next_pow2(std::max<uint64_t>(64, 4 * static_cast<uint64_t>(num_voxels)))); }

b2_status alloc_map(b2_ctx* ctx, double resolution, size_t V, b2_voxelmap** out) {
  b2_voxelmap* vm = new b2_voxelmap;
  vm->ctx = ctx;
  vm->resolution = resolution;
  vm->inv_resolution = 1.0 / resolution;  // ann/impl/incremental_voxelmap_impl.hpp:14
  vm->num_voxels = V;
  vm->num_buckets = bucket_count_for(V);
  cudaError_t e;
  const size_t vr = std::max<size_t>(V, 1);
  if ((e = cudaMalloc(reinterpret_cast<void**>(&vm->d_buckets), vm->num_buckets * sizeof(VoxelBucket))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&vm->d_records), vr * kRecordDoubles * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&vm->d_coords), vr * 3 * sizeof(int32_t))) != cudaSuccess) {
    b2_voxelmap_destroy(vm);
    return fail(B2_ERR_OUT_OF_MEMORY, "voxelmap allocation: %s", cudaGetErrorString(e));
  }
  vm->device_bytes = vm->num_buckets * sizeof(VoxelBucket) + vr * kRecordDoubles * sizeof(double) + vr * 3 * sizeof(int32_t);
  *out = vm;
  return B2_OK;
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" {

b2_status b2_voxelmap_create_from_voxels(b2_ctx* ctx, double resolution, const int32_t* coords, const double* means, const double* covs,
                                         const int32_t* num_points, size_t V, b2_voxelmap** out) {
  B2_REQUIRE(out != nullptr, "b2_voxelmap_create_from_voxels: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_voxelmap_create_from_voxels: ctx is NULL");
  B2_REQUIRE(resolution > 0.0, "b2_voxelmap_create_from_voxels: resolution must be positive");
  B2_REQUIRE(V == 0 || (coords && means && covs), "b2_voxelmap_create_from_voxels: NULL voxel arrays");
  B2_REQUIRE(V < (1ull << 30), "b2_voxelmap_create_from_voxels: too many voxels");
  B2_CUDA(cudaSetDevice(ctx->device));

  // host-side table build in id order (deterministic placement), duplicate coordinates rejected
  const size_t nb = bucket_count_for(V);
  const uint32_t gmask = static_cast<uint32_t>(nb / kGroup - 1);
  std::vector<VoxelBucket> buckets(nb, VoxelBucket{-1, -1, -1, -1});
  std::vector<double> records(std::max<size_t>(V, 1) * kRecordDoubles, 0.0);
  for (size_t r = 0; r < V; r++) {
    const int x = coords[r * 3], y = coords[r * 3 + 1], z = coords[r * 3 + 2];
    uint32_t g = voxel_hash(x, y, z) & gmask;
    bool placed = false;
    while (!placed) {
      for (int k = 0; k < kGroup && !placed; k++) {
        VoxelBucket& b = buckets[static_cast<size_t>(g) * kGroup + k];
        if (b.id < 0) {
          b = VoxelBucket{x, y, z, static_cast<int32_t>(r)};
          placed = true;
        } else if (b.x == x && b.y == y && b.z == z) {
          return fail(B2_ERR_INVALID_ARGUMENT, "b2_voxelmap_create_from_voxels: duplicate voxel coordinate (%d, %d, %d) at ids %d and %zu", x, y, z, b.id, r);
        }
      }
      g = (g + 1) & gmask;
    }
    double* rec = &records[r * kRecordDoubles];
    for (int k = 0; k < 3; k++) rec[k] = means[r * 3 + k];
    const double* c = covs + r * 9;
    rec[3] = c[0];
    rec[4] = c[1];
    rec[5] = c[2];
    rec[6] = c[4];
    rec[7] = c[5];
    rec[8] = c[8];
    rec[9] = num_points ? static_cast<double>(num_points[r]) : 1.0;
  }

  b2_voxelmap* vm = nullptr;
  B2_TRY(alloc_map(ctx, resolution, V, &vm));
  cudaStream_t st = ctx->stream;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(vm->d_buckets, buckets.data(), nb * sizeof(VoxelBucket), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(vm->d_records, records.data(), records.size() * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
      (V > 0 && (e = cudaMemcpyAsync(vm->d_coords, coords, V * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) ||
      (e = cudaStreamSynchronize(st)) != cudaSuccess) {
    b2_voxelmap_destroy(vm);
    return fail(B2_ERR_CUDA, "b2_voxelmap_create_from_voxels: %s", cudaGetErrorString(e));
  }
  *out = vm;
  return B2_OK;
}

b2_status b2_voxelmap_create_from_points(b2_ctx* ctx, double resolution, const double* points, int point_stride, const double* covs, int cov_stride,
                                         size_t n, b2_voxelmap** out) {
  B2_REQUIRE(out != nullptr, "b2_voxelmap_create_from_points: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_voxelmap_create_from_points: ctx is NULL");
  B2_REQUIRE(resolution > 0.0, "b2_voxelmap_create_from_points: resolution must be positive");
  B2_REQUIRE(point_stride == 3 || point_stride == 4, "b2_voxelmap_create_from_points: point_stride must be 3 or 4");
  B2_REQUIRE(cov_stride == 9 || cov_stride == 16, "b2_voxelmap_create_from_points: cov_stride must be 9 or 16");
  B2_REQUIRE(n == 0 || (points && covs), "b2_voxelmap_create_from_points: points and covs are required");  // reference: GaussianVoxel::add reads covs
  B2_REQUIRE(n < (1ull << 31), "b2_voxelmap_create_from_points: at most 2^31-1 points");
  B2_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const double inv_leaf = 1.0 / resolution;

  if (n == 0) {
    return b2_voxelmap_create_from_voxels(ctx, resolution, nullptr, nullptr, nullptr, nullptr, 0, out);
  }

  DevBuf raw_p, raw_c, key_z, key_z2, idx_a, idx_b, key_xy, key_xy2, flags, seg_incl, tmp;
  B2_CUDA(cudaMalloc(&raw_p.p, n * point_stride * sizeof(double)));
  B2_CUDA(cudaMalloc(&raw_c.p, n * cov_stride * sizeof(double)));
  B2_CUDA(cudaMemcpyAsync(raw_p.p, points, n * point_stride * sizeof(double), cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMemcpyAsync(raw_c.p, covs, n * cov_stride * sizeof(double), cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMalloc(&key_z.p, n * 4));
  B2_CUDA(cudaMalloc(&key_z2.p, n * 4));
  B2_CUDA(cudaMalloc(&idx_a.p, n * 4));
  B2_CUDA(cudaMalloc(&idx_b.p, n * 4));
  B2_CUDA(cudaMalloc(&key_xy.p, n * 8));
  B2_CUDA(cudaMalloc(&key_xy2.p, n * 8));
  B2_CUDA(cudaMalloc(&flags.p, n * 4));
  B2_CUDA(cudaMalloc(&seg_incl.p, n * 4));

  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  const double* rp = raw_p.as<double>();
  const double* rc = raw_c.as<double>();
  const int ni = static_cast<int>(n);

  size_t t1 = 0, t2 = 0, t3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, key_z.as<uint32_t>(), key_z2.as<uint32_t>(), idx_a.as<uint32_t>(), idx_b.as<uint32_t>(), ni, 0, 32, st);
  cub::DeviceRadixSort::SortPairs(nullptr, t2, key_xy.as<unsigned long long>(), key_xy2.as<unsigned long long>(), idx_b.as<uint32_t>(), idx_a.as<uint32_t>(), ni, 0, 64, st);
  cub::DeviceScan::InclusiveSum(nullptr, t3, flags.as<uint32_t>(), seg_incl.as<uint32_t>(), ni, st);
  const size_t tmp_bytes = std::max(std::max(t1, t2), std::max(t3, static_cast<size_t>(16)));
  B2_CUDA(cudaMalloc(&tmp.p, tmp_bytes));

  // pass 1: stable sort by z; pass 2: stable sort by (x, y)  => lexicographic (x, y, z), insertion order inside a voxel
  point_coord_keys_kernel<<<grid, 256, 0, st>>>(rp, point_stride, n, inv_leaf, key_z.as<uint32_t>(), idx_a.as<uint32_t>());
  size_t tb = tmp_bytes;
  B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key_z.as<uint32_t>(), key_z2.as<uint32_t>(), idx_a.as<uint32_t>(), idx_b.as<uint32_t>(), ni, 0, 32, st));
  gather_xy_keys_kernel<<<grid, 256, 0, st>>>(rp, point_stride, n, inv_leaf, idx_b.as<uint32_t>(), key_xy.as<unsigned long long>());
  tb = tmp_bytes;
  B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key_xy.as<unsigned long long>(), key_xy2.as<unsigned long long>(), idx_b.as<uint32_t>(), idx_a.as<uint32_t>(), ni, 0, 64, st));
  const uint32_t* idx = idx_a.as<uint32_t>();

  head_flags_kernel<<<grid, 256, 0, st>>>(rp, point_stride, n, inv_leaf, idx, flags.as<uint32_t>());
  tb = tmp_bytes;
  B2_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, flags.as<uint32_t>(), seg_incl.as<uint32_t>(), ni, st));
  uint32_t V32 = 0;
  B2_CUDA(cudaMemcpyAsync(&V32, seg_incl.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const size_t V = V32;

  DevBuf seg_start, seg_first, seg_first2, seg_iota, seg_of_rank, tmp2;
  B2_CUDA(cudaMalloc(&seg_start.p, V * 4));
  B2_CUDA(cudaMalloc(&seg_first.p, V * 4));
  B2_CUDA(cudaMalloc(&seg_first2.p, V * 4));
  B2_CUDA(cudaMalloc(&seg_iota.p, V * 4));
  B2_CUDA(cudaMalloc(&seg_of_rank.p, V * 4));
  segment_heads_kernel<<<grid, 256, 0, st>>>(n, flags.as<uint32_t>(), seg_incl.as<uint32_t>(), idx, seg_start.as<uint32_t>(), seg_first.as<uint32_t>(),
                                             seg_iota.as<uint32_t>());
  size_t t4 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t4, seg_first.as<uint32_t>(), seg_first2.as<uint32_t>(), seg_iota.as<uint32_t>(), seg_of_rank.as<uint32_t>(), static_cast<int>(V), 0, 32, st);
  B2_CUDA(cudaMalloc(&tmp2.p, std::max<size_t>(t4, 16)));
  B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, t4, seg_first.as<uint32_t>(), seg_first2.as<uint32_t>(), seg_iota.as<uint32_t>(), seg_of_rank.as<uint32_t>(), static_cast<int>(V), 0, 32, st));

  b2_voxelmap* vm = nullptr;
  B2_TRY(alloc_map(ctx, resolution, V, &vm));
  cudaError_t e;
  const unsigned vgrid = static_cast<unsigned>((V + 127) / 128);
  accumulate_voxels_kernel<<<vgrid, 128, 0, st>>>(rp, point_stride, rc, cov_stride, n, V, inv_leaf, idx, seg_start.as<uint32_t>(), seg_of_rank.as<uint32_t>(),
                                                  vm->d_records, vm->d_coords);
  if ((e = cudaMemsetAsync(vm->d_buckets, 0xFF, vm->num_buckets * sizeof(VoxelBucket), st)) != cudaSuccess) {
    b2_voxelmap_destroy(vm);
    return fail(B2_ERR_CUDA, "b2_voxelmap_create_from_points: %s", cudaGetErrorString(e));
  }
  insert_buckets_kernel<<<vgrid, 128, 0, st>>>(vm->d_coords, V, vm->d_buckets, static_cast<uint32_t>(vm->num_buckets / kGroup - 1));
  if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(st)) != cudaSuccess) {
    b2_voxelmap_destroy(vm);
    return fail(B2_ERR_CUDA, "b2_voxelmap_create_from_points: %s", cudaGetErrorString(e));
  }
  *out = vm;
  return B2_OK;
}

b2_status b2_voxelmap_destroy(b2_voxelmap* vm) {
  if (!vm) return B2_OK;
  cudaSetDevice(vm->ctx->device);
  if (vm->d_buckets) cudaFree(vm->d_buckets);
  if (vm->d_records) cudaFree(vm->d_records);
  if (vm->d_coords) cudaFree(vm->d_coords);
  delete vm;
  return B2_OK;
}

b2_status b2_voxelmap_get_info(const b2_voxelmap* vm, b2_voxelmap_info* info) {
  B2_REQUIRE(vm && info, "b2_voxelmap_get_info: NULL argument");
  info->num_voxels = vm->num_voxels;
  info->num_buckets = vm->num_buckets;
  info->resolution = vm->resolution;
  info->device_bytes = vm->device_bytes;
  return B2_OK;
}

b2_status b2_voxelmap_download(const b2_voxelmap* vm, int32_t* coords, double* means, double* covs, int32_t* num_points) {
  B2_REQUIRE(vm != nullptr, "b2_voxelmap_download: vm is NULL");
  const size_t V = vm->num_voxels;
  if (V == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(vm->ctx->device));
  cudaStream_t st = vm->ctx->stream;
  std::vector<double> rec(V * kRecordDoubles);
  B2_CUDA(cudaMemcpyAsync(rec.data(), vm->d_records, rec.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (coords) B2_CUDA(cudaMemcpyAsync(coords, vm->d_coords, V * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  for (size_t r = 0; r < V; r++) {
    const double* q = &rec[r * kRecordDoubles];
    if (means)
      for (int k = 0; k < 3; k++) means[r * 3 + k] = q[k];
    if (covs) {
      double* c = covs + r * 9;
      c[0] = q[3];
      c[1] = c[3] = q[4];
      c[2] = c[6] = q[5];
      c[4] = q[6];
      c[5] = c[7] = q[7];
      c[8] = q[8];
    }
    if (num_points) num_points[r] = static_cast<int32_t>(q[9]);
  }
  return B2_OK;
}

b2_status b2_voxelmap_lookup(const b2_voxelmap* vm, const double* points, int point_stride, size_t n, int32_t* out_index) {
  B2_REQUIRE(vm && (n == 0 || (points && out_index)), "b2_voxelmap_lookup: NULL argument");
  B2_REQUIRE(point_stride == 3 || point_stride == 4, "b2_voxelmap_lookup: point_stride must be 3 or 4");
  if (n == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(vm->ctx->device));
  cudaStream_t st = vm->ctx->stream;
  DevBuf dp, di;
  B2_CUDA(cudaMalloc(&dp.p, n * point_stride * sizeof(double)));
  B2_CUDA(cudaMalloc(&di.p, n * sizeof(int32_t)));
  B2_CUDA(cudaMemcpyAsync(dp.p, points, n * point_stride * sizeof(double), cudaMemcpyHostToDevice, st));
  lookup_points_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(dp.as<double>(), point_stride, n, vm->inv_resolution, vm->d_buckets,
                                                                              static_cast<uint32_t>(vm->num_buckets / kGroup - 1), di.as<int32_t>());
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaMemcpyAsync(out_index, di.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  return B2_OK;
}

}  // extern "C"
