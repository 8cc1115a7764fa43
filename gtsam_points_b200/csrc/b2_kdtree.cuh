// b2_kdtree.cuh -- device-side exact 1-NN traversal of the library's kd-tree (shared by b2_kdtree.cu and b2_factors.cu).
//
// Replaces the recursive UnsafeKdTree::knn_search (reference: include/gtsam_points/ann/small_kdtree.hpp:436-476) for
// k = 1 with KnnResult semantics of include/gtsam_points/ann/knn_result.hpp:69-71,89-109: the best distance starts at
// max_sq_dist, a candidate replaces the best iff its squared distance is strictly smaller, the far child is visited iff
// best > cut^2.  The search is exact, so the result equals the reference's for any tree shape (barring exact ties).
//
// Layout (private to this library): 16-byte nodes (one LDG.128 per visit), children of an internal node adjacent,
// points stored in LEAF ORDER as three float64 planes so that a leaf scan reads contiguous memory and neighbouring
// queries (the source cloud is Morton-ordered) read the same lines.
#pragma once

#include "b2_internal.hpp"

namespace b2 {

constexpr int kKdStackDepth = 40;  // > log2(2^31 / leaf) with margin

struct KdTreeView {
  const KdNodeGPU* nodes;
  const double* px;
  const double* py;
  const double* pz;
};

__device__ __forceinline__ KdNodeGPU load_node(const KdNodeGPU* p) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(p));
  KdNodeGPU n;
  n.thresh = __hiloint2double(v.y, v.x);
  n.a = static_cast<uint32_t>(v.z);
  n.b = static_cast<uint32_t>(v.w);
  return n;
}

// Returns the leaf-order position of the nearest point with squared distance < max_sq (else -1); *out_sq = that distance.
// Squared distance is evaluated as (dx*dx + dy*dy) + dz*dz with individually rounded operations, the same order as the
// CPU oracle, so distances (and therefore arg-min decisions) are bit-identical to the float64 CPU path.
__device__ __forceinline__ int kdtree_nn1(const KdTreeView& t, double qx, double qy, double qz, double max_sq, double* out_sq) {
  uint32_t stack_node[kKdStackDepth];
  double stack_cut[kKdStackDepth];
  int sp = 0;
  double best = max_sq;
  int best_j = -1;
  uint32_t node_idx = 0;
  double cut = -1.0;  // root is always visited
  while (true) {
    if (best > cut) {
      // descend to a leaf, pushing far children
      KdNodeGPU n = load_node(t.nodes + node_idx);
      while (n.b < 4u) {
        const double qa = n.b == 0u ? qx : (n.b == 1u ? qy : qz);
        const double diff = __dsub_rn(qa, n.thresh);
        const uint32_t near_c = diff < 0.0 ? n.a : n.a + 1u;
        const uint32_t far_c = diff < 0.0 ? n.a + 1u : n.a;
        stack_node[sp] = far_c;
        stack_cut[sp] = __dmul_rn(diff, diff);
        sp++;
        n = load_node(t.nodes + near_c);
      }
      const uint32_t first = n.a, cnt = n.b - 4u;
      for (uint32_t j = first; j < first + cnt; j++) {
        const double dx = __dsub_rn(__ldg(t.px + j), qx);
        const double dy = __dsub_rn(__ldg(t.py + j), qy);
        const double dz = __dsub_rn(__ldg(t.pz + j), qz);
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d < best) {
          best = d;
          best_j = static_cast<int>(j);
        }
      }
    }
    if (sp == 0) break;
    sp--;
    node_idx = stack_node[sp];
    cut = stack_cut[sp];
  }
  *out_sq = best;
  return best_j;
}

}  // namespace b2
