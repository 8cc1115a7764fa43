// b2_device.cuh -- device helpers shared by the voxel-map and factor kernels.
#pragma once

#include "b2_internal.hpp"

namespace b2 {

// Voxel coordinate of a transformed point: fast_floor(x * inv_leaf_size)
// (reference: util/fast_floor.hpp:12-15, src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:59-61).
// Explicitly rounded multiply (no FMA contraction) so that the result is bit-identical to the CPU float64 path.
// fast_floor is "truncate, then subtract 1 if the value is below its truncation", i.e. floor(v) for every finite v in int
// range: one round-toward-minus-infinity conversion gives the same integer.
__device__ __forceinline__ int voxel_coord1(double q, double inv_leaf) { return __double2int_rd(__dmul_rn(q, inv_leaf)); }

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

// Bucket groups: the table is an array of groups of kGroup 16-byte buckets (32 bytes = one sector for kGroup = 2); a key hashes
// to a group and is stored in the first group of its (linear, over groups) probe sequence that had a free slot at insertion
// time.  Buckets are never deleted, so a lookup may stop at the first group that still has an empty slot.  The table is sized
// for a load factor <= 0.25: ~95 % of all lookups (hits AND misses) resolve inside the home group, i.e. with ONE access -- on
// the GPU the probe chain is pure latency, so its length is what counts -- and the factor kernel can stage the home group of
// every point asynchronously (32 bytes per point in flight); the few keys whose home group is full continue with ordinary
// loads.  (Round 1 used 64-byte groups of 4: 98 % one-access lookups, but twice the bytes per probe in flight.)
#ifndef B2_BUCKET_GROUP
#define B2_BUCKET_GROUP 2
#endif
constexpr int kGroup = B2_BUCKET_GROUP;

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

// returns id >= 0 (found), -1 (absent), -2 (group full and key not in it: continue with the next group).
// Branch-free: a bucket matches iff ((bx ^ x) | (by ^ y) | (bz ^ z)) == 0; ids are >= 0 and an empty bucket carries id -1
// (whatever its coordinate bytes), so "max over buckets of (match ? id : -1)" is the id of the match or -1, and the group
// has an empty slot iff the minimum stored id is negative.
__device__ __forceinline__ int match_group(const BucketGroup& g, int x, int y, int z) {
  int id = -1, lo = 0;
#pragma unroll
  for (int k = 0; k < kGroup; k++) {
    const int diff = (g.b[k].x ^ x) | (g.b[k].y ^ y) | (g.b[k].z ^ z);
    id = max(id, diff == 0 ? g.b[k].w : -1);
    lo = min(lo, g.b[k].w);
  }
  return id >= 0 ? id : (lo < 0 ? -1 : -2);
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
