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

// Bucket groups: the table is an array of 64-byte groups of kGroup = 4 buckets; a key hashes to a group and is stored
// in the first group of its (linear, over groups) probe sequence that had a free slot at insertion time.  Buckets are
// never deleted, so a lookup may stop at the first group that still has an empty slot.  The table is sized for a load
// factor <= 0.25, which makes >98% of all lookups (hits AND misses) resolve with ONE round trip of four independent
// 16-byte loads out of two 32-byte sectors -- on the GPU the probe chain is pure latency, so its length is what counts.
constexpr int kGroup = 4;

struct BucketGroup {
  int4 b[kGroup];
};

__device__ __forceinline__ BucketGroup load_group(const VoxelBucket* __restrict__ buckets, uint32_t g) {
  const int4* p = reinterpret_cast<const int4*>(buckets) + static_cast<size_t>(g) * kGroup;
  BucketGroup r;
#pragma unroll
  for (int k = 0; k < kGroup; k++) r.b[k] = __ldg(p + k);
  return r;
}

// returns id >= 0 (found), -1 (absent), -2 (group full and key not in it: continue with the next group)
__device__ __forceinline__ int match_group(const BucketGroup& g, int x, int y, int z) {
  int id = -2;
  bool has_empty = false;
#pragma unroll
  for (int k = 0; k < kGroup; k++) {
    if (g.b[k].w >= 0 && g.b[k].x == x && g.b[k].y == y && g.b[k].z == z) id = g.b[k].w;
    has_empty |= g.b[k].w < 0;
  }
  return id >= 0 ? id : (has_empty ? -1 : -2);
}

// Exact lookup (no max_bucket_scan_count cut-off as in the reference's GPU map: a present voxel is always found).
// group_mask = number of groups - 1.
__device__ __forceinline__ int lookup_voxel(const VoxelBucket* __restrict__ buckets, uint32_t group_mask, int x, int y, int z) {
  uint32_t g = voxel_hash(x, y, z) & group_mask;
  while (true) {
    const int r = match_group(load_group(buckets, g), x, y, z);
    if (r != -2) return r;
    g = (g + 1) & group_mask;
  }
}

}  // namespace b2
