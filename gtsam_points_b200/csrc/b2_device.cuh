// b2_device.cuh -- device helpers shared by the voxel-map and factor kernels.
#pragma once

#include "b2_internal.hpp"

namespace b2 {

// Voxel coordinate of a transformed point: fast_floor(x * inv_leaf_size)
// (reference: util/fast_floor.hpp:12-15, src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:59-61).
// Explicitly rounded multiply (no FMA contraction) so that the result is bit-identical to the CPU float64 path.
__device__ __forceinline__ int voxel_coord1(double q, double inv_leaf) {
  const double v = __dmul_rn(q, inv_leaf);
  const int n = __double2int_rz(v);
  return n - (v < static_cast<double>(n) ? 1 : 0);
}

// Bucket index of an integer voxel coordinate.  Our own mixing function (the table layout is private to this
// library; the reference's XORVector3iHash / boost hash_combine only matter for ITS containers).
__host__ __device__ __forceinline__ uint32_t voxel_hash(int x, int y, int z) {
  uint32_t h = static_cast<uint32_t>(x) * 0x8da6b343u ^ static_cast<uint32_t>(y) * 0xd8163841u ^ static_cast<uint32_t>(z) * 0xcb1ab31fu;
  h ^= h >> 15;
  h *= 0x2c1b3c6du;
  h ^= h >> 12;
  return h;
}

__device__ __forceinline__ VoxelBucket load_bucket(const VoxelBucket* p) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(p));
  VoxelBucket b;
  b.x = v.x;
  b.y = v.y;
  b.z = v.z;
  b.id = v.w;
  return b;
}

// Exact lookup: linear probing until the key or an empty bucket is found (the table is never full: load <= 0.5).
// Unlike the reference's GPU map there is no max_bucket_scan_count cut-off, so a present voxel is always found.
__device__ __forceinline__ int lookup_voxel(const VoxelBucket* __restrict__ buckets, uint32_t mask, int x, int y, int z) {
  uint32_t h = voxel_hash(x, y, z) & mask;
  while (true) {
    const VoxelBucket b = load_bucket(buckets + h);
    if (b.id < 0) return -1;
    if (b.x == x && b.y == y && b.z == z) return b.id;
    h = (h + 1) & mask;
  }
}

}  // namespace b2
