// b2_factors.cu -- the hot path: fused correspondence search + linearization of VGICP / GICP factors.
//
// Replaces, in ONE persistent kernel per (factor kind, storage type) group of a factor set:
//   * IntegratedVGICPFactor_::update_correspondences + ::evaluate  (reference: include/gtsam_points/factors/impl/
//     integrated_vgicp_factor_impl.hpp:99-172, :175-257), IntegratedGICPFactor_ likewise (impl/integrated_gicp_factor_impl.hpp:132-296),
//   * scan_matching_reduce_omp (impl/scan_matching_reduction.hpp:16-68),
//   * the reference's GPU twin: lookup_voxels_kernel / vgicp_derivatives_kernel / cub::DeviceReduce of LinearizedSystem6
//     (include/gtsam_points/cuda/kernels/*.cuh, src/gtsam_points/factors/integrated_vgicp_derivatives_linearize.cu:23-55),
//   * NonlinearFactorSetGPU's per-factor issue/sync loop (src/gtsam_points/cuda/nonlinear_factor_set_gpu.cpp:64-218).
//
// Design (B200-first):
//   * float64 arithmetic end to end (pose * point and the voxel floor use individually rounded operations so that
//     correspondence indices are bit-identical to the CPU float64 path); storage may be float32 where that is lossless.
//   * per correspondence only A' = J'^T M J' (21 unique), c' = J'^T M r (6) and e are accumulated, with
//     J' = [-hat(R p) | I]; since J_target = J' X and J_source = -J' D with X = [[I,0],[-hat(t),I]], D = diag(R,R),
//     the five reference blocks are recovered once per factor: H_t = X^T A' X, H_s = D^T A' D, H_ts = -X^T A' D,
//     b_t = X^T c', b_s = -D^T c'.  (~170 DFMA-class instructions per correspondence instead of ~1000.)
//   * the kernel (b2_factor_kernel_ws.cuh) is warp-specialised: probe warps search correspondences and feed the hits,
//     compacted, through shared-memory rings to accumulate warps that own the float64 accumulators.
//   * one transposing butterfly reduction (31 shuffles for 32 values) per CTA and factor; partial sums go to fixed
//     slots and the last CTA of a factor reduces the slots in slot order => results are bit-reproducible run to run.
//   * all factors of a set are covered by one launch: the grid walks a tile list (tile -> factor), CTA c takes the
//     contiguous, balanced tile range [c T / G, (c + 1) T / G), so a CTA meets few factors and each in one run.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "b2_device.cuh"
#include "b2_kdtree.cuh"

namespace b2 {

constexpr int kAcc = 32;  // accumulator slots per partial record (29 used)

enum { MODE_LINEARIZE = 0, MODE_ERROR = 1 };
constexpr int kMaxPeers = 8;  // GPUs of one NVSwitch node

// Optional completion signal of a host call: the CTA that finishes the LAST factor of the call (device counter == total)
// stores `seq` to a word in pinned, mapped host memory after the results (also mapped) have been fenced at system scope;
// the host spins on that word instead of paying cudaStreamSynchronize's completion latency.  flag == nullptr: unused.
struct DoneSignal {
  unsigned int* counter;
  volatile unsigned int* flag;
  unsigned int total;
  unsigned int seq;
  // Multi-GPU exchange fused into the epilogue (b2_factor_set_linearize_exchange): besides its own buffer, the CTA that
  // finishes a factor stores the 1 KiB record into the same slot of every peer GPU's buffer (NVLink peer stores), and the
  // CTA that finishes the LAST local factor then raises flag[my_rank] = seq in every GPU's flag array (its own included).
  // n_peers == 0: unused.  `flag` doubles as this GPU's own flag array in that mode.
  int n_peers;
  int my_rank;
  int wait_in_kernel;  // the signalling thread also waits until every rank's flag in THIS GPU's array shows seq: the launch completes = the exchange completed
  double* peer_out[kMaxPeers];          // same offset as `out` of the launch, in each peer's buffer (entry my_rank unused)
  unsigned int* peer_flag[kMaxPeers];   // each GPU's flag array [n_peers]
  // Host delivery of an exchange step (b2_exchange_linearize_host): once every rank's flag has arrived, the CTA that waited copies
  // ALL records of the step from this GPU's block into pinned mapped host memory and then raises a mapped completion word: the
  // host gets every rank's records without a copy operation or a stream synchronisation.  mirror == nullptr: unused.
  const double* mirror_src;
  double* mirror;
  unsigned int mirror_doubles;
  unsigned int mirror_seq;              // value the host waits for in *mirror_flag
  volatile unsigned int* mirror_flag;
};

// stride S ~ n / golden ratio with gcd(S, n) == 1: v -> (v * S) mod n is a permutation that spreads any run of v evenly
inline uint32_t golden_stride(uint32_t n) {
  if (n <= 2u) return 1u;
  uint32_t s = static_cast<uint32_t>(static_cast<double>(n) * 0.6180339887498949);
  if (s < 1u) s = 1u;
  auto gcd = [](uint32_t a, uint32_t b) {
    while (b) {
      const uint32_t r = a % b;
      a = b;
      b = r;
    }
    return a;
  };
  while (gcd(s, n) != 1u) s++;
  return s % n == 0u ? 1u : s % n;
}

// called by ONE thread of the CTA that finished a factor, after that factor's results were fenced at system scope
// returns true iff this caller completed the whole call (and, with wait_in_kernel, has seen every rank's flag)
__device__ __forceinline__ bool signal_done(const DoneSignal& sig) {
  if (sig.flag == nullptr) return false;
  // a call of ONE factor needs no counter: its only finisher publishes directly (saves a device atomic's round trip on the
  // latency path of the headline call)
  const unsigned int prev = sig.total == 1u ? 0u : atomicAdd(sig.counter, 1u);
  if (prev == sig.total - 1u) {  // every factor of this call is done: re-arm the counter, publish the sequence number
    if (sig.total != 1u) *sig.counter = 0u;
    // acquire side of the counter chain: the other factors' CTAs fenced their records before their atomicAdd; with a single
    // factor the caller's own fence (just before this call) already ordered the record before the flag
    if (sig.total > 1u) __threadfence_system();
    if (sig.n_peers > 0) {
      for (int p = 0; p < sig.n_peers; p++) *reinterpret_cast<volatile unsigned int*>(sig.peer_flag[p] + sig.my_rank) = sig.seq;  // every GPU, own included
      if (sig.wait_in_kernel) {
        // fold the flag wait into this launch (no second kernel): every other CTA of this GPU is done or draining, so one
        // spinning thread costs nothing; bounded (a rank may still be loading its modules at the first step: ~30 s)
        const volatile unsigned int* mine = sig.peer_flag[sig.my_rank];
        for (int r = 0; r < sig.n_peers; r++) {
          unsigned polls = 0;
          while (mine[r] != sig.seq) {
            __nanosleep(64);
            if (++polls > (1u << 28)) __trap();
          }
        }
        __threadfence_system();
      }
    } else {
      *sig.flag = sig.seq;
    }
    return true;
  }
  return false;
}

// all `nthreads` threads of the CTA whose thread 0 got `true` from signal_done: copy the step's records to the host mirror
__device__ __forceinline__ void mirror_to_host(const DoneSignal& sig, int tid, int nthreads) {
  for (unsigned int i = tid; i < sig.mirror_doubles; i += nthreads) sig.mirror[i] = __ldcg(sig.mirror_src + i);
  __threadfence_system();
}

struct FactorDesc {
  const void* pts;    // 3 planes of n_pad
  const void* covs;   // 6 planes of n_pad
  uint32_t n;
  uint32_t n_pad;
  // VGICP target
  const VoxelBucket* buckets;
  uint32_t bucket_mask;
  uint32_t num_records;  // voxels in the map (length of `records` for the voxel path)
  double inv_leaf;
  // GICP target
  const KdNodeGPU* nodes;
  const void* leaf_pts;  // leaf-order point records (float4 or 4 doubles each)
  uint32_t leaf_f32;
  uint32_t pad1;
  double max_sq;
  // mean(3) | cov(6) | count records: voxels (id order) or target points (leaf order)
  const double* records;
  int32_t* corr;
  double* lin_pose;  // per-factor linearization point (16 doubles, row-major 4x4), owned by the factor
  uint32_t tile_begin;
  uint32_t num_tiles;
  uint32_t slot_begin[2];  // per mode
  uint32_t num_slots[2];   // per mode
  uint32_t out_index;      // index of the factor in its set (pose / result / counter)
  uint32_t cta_first[2];   // per mode: first CTA whose (contiguous) tile range touches this factor; slot = blockIdx.x - cta_first
  uint32_t perm_stride;    // virtual -> physical tile permutation within the factor: (v * perm_stride) mod num_tiles
};

// Single-factor launches pass the pose (row-major 4x4) AND the factor's descriptor BY VALUE as a kernel parameter: the arithmetic
// takes the pose as constant-bank operands, the descriptor costs no dependent (cold) global load at kernel start, and the kernel
// reads nothing from host memory on its way in.
struct PoseArg {
  double m[16];
  FactorDesc desc;  // valid for single-factor launches (copy of the set's only descriptor)
};

__device__ __forceinline__ double ldv(const float* p, size_t i) { return static_cast<double>(__ldg(p + i)); }
__device__ __forceinline__ double ldv(const double* p, size_t i) { return __ldg(p + i); }

template <int OFF>
__device__ __forceinline__ void bfly_step(double (&v)[kAcc], int lane) {
  const bool upper = (lane & OFF) != 0;
#pragma unroll
  for (int k = 0; k < OFF; k++) {
    const double send = upper ? v[k] : v[k + OFF];
    const double keep = upper ? v[k + OFF] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

// Transposing butterfly: 32 values per lane in, lane l ends up holding the warp total of value l (31 shuffles, fixed order).
__device__ __forceinline__ double warp_reduce32(double (&v)[kAcc], int lane) {
  bfly_step<16>(v, lane);
  bfly_step<8>(v, lane);
  bfly_step<4>(v, lane);
  bfly_step<2>(v, lane);
  bfly_step<1>(v, lane);
  return v[0];
}

__device__ __forceinline__ int sym_idx(int i, int j) {
  if (i > j) {
    const int t = i;
    i = j;
    j = t;
  }
  return i * 3 - (i * (i - 1)) / 2 + (j - i);
}

// accumulator layout: 0..5 A_rr (upper), 6..14 A_rt (row-major, rows = rotation), 15..20 A_tt (upper), 21..23 c_r, 24..26 c_t, 27 error, 28 count

// Per-factor epilogue, step 1 (threads 0..35): unpack A' (symmetric 6x6, tangent order [rot, trans]) from the reduced
// accumulators and build X = [[I, 0], [-hat(t), I]] and D = diag(R, R).
__device__ __forceinline__ void epilogue_build(double* __restrict__ A, double* __restrict__ X, double* __restrict__ D, const double* __restrict__ tot,
                                               const double* __restrict__ R, const double* __restrict__ t, int tid) {
  if (tid >= 36) return;
  const int i = tid / 6, j = tid % 6;
  double a;
  if (i < 3 && j < 3)
    a = tot[sym_idx(i, j)];
  else if (i < 3)
    a = tot[6 + i * 3 + (j - 3)];
  else if (j < 3)
    a = tot[6 + j * 3 + (i - 3)];
  else
    a = tot[15 + sym_idx(i - 3, j - 3)];
  A[tid] = a;
  double x = (i == j) ? 1.0 : 0.0;
  if (i >= 3 && j < 3) {
    const int r = i - 3, cc = j;
    // -hat(t) = [[0, t2, -t1], [-t2, 0, t0], [t1, -t0, 0]]
    if (r == 0 && cc == 1) x = t[2];
    if (r == 0 && cc == 2) x = -t[1];
    if (r == 1 && cc == 0) x = -t[2];
    if (r == 1 && cc == 2) x = t[0];
    if (r == 2 && cc == 0) x = t[1];
    if (r == 2 && cc == 1) x = -t[0];
  }
  X[tid] = x;
  D[tid] = ((i < 3) == (j < 3)) ? R[(i % 3) * 3 + (j % 3)] : 0.0;
}

// Step 2 (after a barrier): H_t = X^T A' X, H_s = D^T A' D, H_ts = -X^T A' D (threads 0..35), b_t = X^T c', b_s = -D^T c'
// (threads 64..69), error / inlier count / padding (thread 96) -> the factor's b2_linearized record.
__device__ __forceinline__ void epilogue_store(double* __restrict__ rec, const double* __restrict__ A, const double* __restrict__ X,
                                               const double* __restrict__ D, const double* __restrict__ tot, int tid) {
  if (tid < 36) {
    const int i = tid / 6, j = tid % 6;
    double ht = 0.0, hs = 0.0, hts = 0.0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
      double xa = 0.0, da = 0.0;  // (A X)[a][j], (A D)[a][j]
#pragma unroll
      for (int b = 0; b < 6; b++) {
        xa += A[a * 6 + b] * X[b * 6 + j];
        da += A[a * 6 + b] * D[b * 6 + j];
      }
      ht += X[a * 6 + i] * xa;
      hs += D[a * 6 + i] * da;
      hts -= X[a * 6 + i] * da;
    }
    rec[tid] = ht;
    rec[36 + tid] = hs;
    rec[72 + tid] = hts;
  } else if (tid >= 64 && tid < 70) {
    const int i = tid - 64;
    double bt = 0.0, bs = 0.0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
      const double ca = tot[21 + a];
      bt += X[a * 6 + i] * ca;
      bs -= D[a * 6 + i] * ca;
    }
    rec[108 + i] = bt;
    rec[114 + i] = bs;
  } else if (tid == 96) {
    rec[120] = tot[27];
    rec[121] = tot[28];
#pragma unroll
    for (int k = 122; k < B2_LINEARIZED_DOUBLES; k++) rec[k] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-correspondence arithmetic (shared by the VGICP and GICP kernels).
//   u = R p (rotated source point), t = translation, (mb, b**) = target mean / covariance, a** = source covariance,
//   RL = rotation of the linearization point (== R when linearizing).
// ---------------------------------------------------------------------------------------------------------------
struct TargetRec {
  double2 r01, r23, r45, r67, r89;  // mean(3) | cov upper(6) | count
};
struct SourceCov {
  double a00, a01, a02, a11, a12, a22;
};

__device__ __forceinline__ TargetRec load_record(const double* __restrict__ records, int id) {
  const double2* rec = reinterpret_cast<const double2*>(records + static_cast<size_t>(id < 0 ? 0 : id) * kRecordDoubles);
  TargetRec r;
  r.r01 = __ldg(rec);
  r.r23 = __ldg(rec + 1);
  r.r45 = __ldg(rec + 2);
  r.r67 = __ldg(rec + 3);
  r.r89 = __ldg(rec + 4);
  return r;
}

template <typename CT>
__device__ __forceinline__ SourceCov load_cov(const CT* __restrict__ cv, size_t n_pad, uint32_t i) {
  SourceCov c;
  c.a00 = ldv(cv, i);
  c.a01 = ldv(cv + n_pad, i);
  c.a02 = ldv(cv + 2 * n_pad, i);
  c.a11 = ldv(cv + 3 * n_pad, i);
  c.a12 = ldv(cv + 4 * n_pad, i);
  c.a22 = ldv(cv + 5 * n_pad, i);
  return c;
}

// 1 / x for a normal, finite x (here: the determinant of a sum of regularised covariances), branch-free: the hardware seed
// (MUFU.RCP64H, ~20 bits) refined by one cubic and one quadratic Newton step -- the sequence the compiler emits for a
// float64 division, minus its special-case branch, so that the per-correspondence arithmetic stays ONE basic block and
// the chains of two correspondences handled by the same lane can interleave.  Result within 1 ulp of 1.0 / x.
__device__ __forceinline__ double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// RLf(i) / tf(i): accessors of the linearization rotation (row-major) and the translation of the evaluation pose -- arrays in
// registers, or by-value kernel parameters (constant-bank operands) for single-factor launches.
// METRIC selects the weight matrix M of r^T M r:
//   0  GICP / VGICP: M = (C_B + RL C_A RL^T)^-1                              (integrated_gicp_factor_impl.hpp:177-183)
//   1  point-to-point ICP: M = I                                             (integrated_icp_factor_impl.hpp:205-248)
//   2  point-to-plane ICP: M = diag(n_x^2, n_y^2, n_z^2), n = target normal  (residual and Jacobian rows scaled by n: N^T N)
//      -- the normal travels in the record's first three covariance slots.
// MASKED: branch-free form for kernels that carry several points per lane through one basic block: `vf` is 1.0 for a real
// correspondence and 0.0 for a lane without one (whose operands are any finite stand-ins); M is scaled by vf, so every
// contribution of a masked lane is an exact zero, and the inlier count is kept by the caller (integer).
template <int MODE, int METRIC = 0, bool MASKED = false, class RLF, class TF>
__device__ __forceinline__ void accumulate_point_f(double (&acc)[kAcc], const RLF& RLf, const TF& tf, double v0, double v1, double v2, const TargetRec& T,
                                                   const SourceCov& A, double vf = 1.0) {
  const double mb0 = T.r01.x, mb1 = T.r01.y, mb2 = T.r23.x;
  const double b00 = T.r23.y, b01 = T.r45.x, b02 = T.r45.y, b11 = T.r67.x, b12 = T.r67.y, b22 = T.r89.x;
  double m00, m01, m02, m11, m12, m22;
  if (METRIC == 1) {
    m00 = m11 = m22 = 1.0;
    m01 = m02 = m12 = 0.0;
  } else if (METRIC == 2) {
    m00 = b00 * b00, m11 = b01 * b01, m22 = b02 * b02;
    m01 = m02 = m12 = 0.0;
  } else {
  // fused covariance S = C_B + RL C_A RL^T (symmetric), M = S^-1 (cofactors / det)
  const double t00 = RLf(0) * A.a00 + RLf(1) * A.a01 + RLf(2) * A.a02;
  const double t01 = RLf(0) * A.a01 + RLf(1) * A.a11 + RLf(2) * A.a12;
  const double t02 = RLf(0) * A.a02 + RLf(1) * A.a12 + RLf(2) * A.a22;
  const double s00 = b00 + (t00 * RLf(0) + t01 * RLf(1) + t02 * RLf(2));
  const double s01 = b01 + (t00 * RLf(3) + t01 * RLf(4) + t02 * RLf(5));
  const double s02 = b02 + (t00 * RLf(6) + t01 * RLf(7) + t02 * RLf(8));
  const double t10 = RLf(3) * A.a00 + RLf(4) * A.a01 + RLf(5) * A.a02;
  const double t11 = RLf(3) * A.a01 + RLf(4) * A.a11 + RLf(5) * A.a12;
  const double t12 = RLf(3) * A.a02 + RLf(4) * A.a12 + RLf(5) * A.a22;
  const double s11 = b11 + (t10 * RLf(3) + t11 * RLf(4) + t12 * RLf(5));
  const double s12 = b12 + (t10 * RLf(6) + t11 * RLf(7) + t12 * RLf(8));
  const double t20 = RLf(6) * A.a00 + RLf(7) * A.a01 + RLf(8) * A.a02;
  const double t21 = RLf(6) * A.a01 + RLf(7) * A.a11 + RLf(8) * A.a12;
  const double t22 = RLf(6) * A.a02 + RLf(7) * A.a12 + RLf(8) * A.a22;
  const double s22 = b22 + (t20 * RLf(6) + t21 * RLf(7) + t22 * RLf(8));

  const double c00 = s11 * s22 - s12 * s12;
  const double c01 = s02 * s12 - s01 * s22;
  const double c02 = s01 * s12 - s02 * s11;
  const double c11 = s00 * s22 - s02 * s02;
  const double c12 = s01 * s02 - s00 * s12;
  const double c22 = s00 * s11 - s01 * s01;
  double inv_det = rcp_nr(s00 * c00 + s01 * c01 + s02 * c02);
  if (MASKED) inv_det *= vf;
  m00 = c00 * inv_det, m01 = c01 * inv_det, m02 = c02 * inv_det;
  m11 = c11 * inv_det, m12 = c12 * inv_det, m22 = c22 * inv_det;
  }

  // residual r = mean_B - (u + t), Mahalanobis error
  const double e0 = mb0 - __dadd_rn(v0, tf(0)), e1 = mb1 - __dadd_rn(v1, tf(1)), e2 = mb2 - __dadd_rn(v2, tf(2));
  const double w0 = m00 * e0 + m01 * e1 + m02 * e2;
  const double w1 = m01 * e0 + m11 * e1 + m12 * e2;
  const double w2 = m02 * e0 + m12 * e1 + m22 * e2;
  acc[27] += e0 * w0 + e1 * w1 + e2 * w2;
  if (!MASKED) acc[28] += 1.0;

  if (MODE == MODE_LINEARIZE) {
    // K = hat(u) M  (column j = u x M[:,j])
    const double k00 = v1 * m02 - v2 * m01, k10 = v2 * m00 - v0 * m02, k20 = v0 * m01 - v1 * m00;
    const double k01 = v1 * m12 - v2 * m11, k11 = v2 * m01 - v0 * m12, k21 = v0 * m11 - v1 * m01;
    const double k02 = v1 * m22 - v2 * m12, k12 = v2 * m02 - v0 * m22, k22 = v0 * m12 - v1 * m02;
    // A_rr = K hat(u)^T
    acc[0] += k02 * v1 - k01 * v2;
    acc[1] += k00 * v2 - k02 * v0;
    acc[2] += k01 * v0 - k00 * v1;
    acc[3] += k10 * v2 - k12 * v0;
    acc[4] += k11 * v0 - k10 * v1;
    acc[5] += k21 * v0 - k20 * v1;
    acc[6] += k00;
    acc[7] += k01;
    acc[8] += k02;
    acc[9] += k10;
    acc[10] += k11;
    acc[11] += k12;
    acc[12] += k20;
    acc[13] += k21;
    acc[14] += k22;
    acc[15] += m00;
    acc[16] += m01;
    acc[17] += m02;
    acc[18] += m11;
    acc[19] += m12;
    acc[20] += m22;
    // c_r = u x (M r), c_t = M r
    acc[21] += v1 * w2 - v2 * w1;
    acc[22] += v2 * w0 - v0 * w2;
    acc[23] += v0 * w1 - v1 * w0;
    acc[24] += w0;
    acc[25] += w1;
    acc[26] += w2;
  }
}

template <int MODE, int METRIC = 0>
__device__ __forceinline__ void accumulate_point(double (&acc)[kAcc], const double (&RL)[9], const double (&t)[3], double v0, double v1, double v2,
                                                 const TargetRec& T, const SourceCov& A) {
  accumulate_point_f<MODE, METRIC>(acc, [&](int i) { return RL[i]; }, [&](int i) { return t[i]; }, v0, v1, v2, T, A);
}

// u = R p : coefficient sums in index order, each operation individually rounded (bit-parity with the CPU float64 path)
__device__ __forceinline__ void rotate_point(const double (&R)[9], double x, double y, double z, double& u0, double& u1, double& u2) {
  u0 = __dadd_rn(__dadd_rn(__dmul_rn(R[0], x), __dmul_rn(R[1], y)), __dmul_rn(R[2], z));
  u1 = __dadd_rn(__dadd_rn(__dmul_rn(R[3], x), __dmul_rn(R[4], y)), __dmul_rn(R[5], z));
  u2 = __dadd_rn(__dadd_rn(__dmul_rn(R[6], x), __dmul_rn(R[7], y)), __dmul_rn(R[8], z));
}

}  // namespace b2

// ---- kernel configurations -------------------------------------------------------------------------------------
// Which kernel serves the VGICP path: 1 = the warp-specialised register-pipelined kernel (b2_factor_kernel_ws.cuh), 2 = the
// TMA / cp.async staged kernel (b2_factor_kernel_v2.cuh).  The default is whichever measured faster on B200
// (profiles/r02_experiments.md); both pass the same parity tests.
#ifndef B2_VGICP_IMPL
#define B2_VGICP_IMPL 1
#endif
#define B2_WS_NAMESPACE ws
#ifndef B2_WS_PRODUCERS
#define B2_WS_PRODUCERS 8
#endif
#ifndef B2_WS_CONSUMERS
#define B2_WS_CONSUMERS 8
#endif
#ifndef B2_WS_REGS_PRODUCER
#define B2_WS_REGS_PRODUCER 88
#endif
#ifndef B2_WS_REGS_CONSUMER
#define B2_WS_REGS_CONSUMER 168
#endif
#ifndef B2_WS_RING
#define B2_WS_RING 256
#endif
#ifndef B2_WS_PPL
#define B2_WS_PPL 2
#endif
#ifndef B2_WS_IPL
#define B2_WS_IPL 2  // two correspondences per accumulate lane in one basic block (pays since the pose is a by-value parameter: 39.4 vs 41.5 us)
#endif
#ifndef B2_WS_COORDS_AHEAD
#define B2_WS_COORDS_AHEAD 0
#endif
#ifndef B2_WS_PREFETCH_OPERANDS
#define B2_WS_PREFETCH_OPERANDS 1
#endif
#include "b2_factor_kernel_ws.cuh"
#undef B2_WS_NAMESPACE
#undef B2_WS_PRODUCERS
#undef B2_WS_CONSUMERS
#undef B2_WS_REGS_PRODUCER
#undef B2_WS_REGS_CONSUMER
#undef B2_WS_RING
#undef B2_WS_PPL
#undef B2_WS_IPL

// VGICP (voxel hash probe): the v2 kernel (TMA-staged streams, shared-memory-fed accumulate warps).
#define B2_V2_NAMESPACE v2
#ifndef B2_V2_PRODUCERS
#define B2_V2_PRODUCERS 8
#endif
#ifndef B2_V2_CONSUMERS
#define B2_V2_CONSUMERS 8
#endif
#ifndef B2_V2_REGS_PRODUCER
#define B2_V2_REGS_PRODUCER 72
#endif
#ifndef B2_V2_REGS_CONSUMER
#define B2_V2_REGS_CONSUMER 128
#endif
#ifndef B2_V2_RING
#define B2_V2_RING 128
#endif
#ifndef B2_V2_PPL
#define B2_V2_PPL 2
#endif
#ifndef B2_V2_XYZ_STAGES
#define B2_V2_XYZ_STAGES 3
#endif
#ifndef B2_V2_BACKOFF_MIN
#define B2_V2_BACKOFF_MIN 32u  // ns: first sleep of a waiting warp
#endif
#ifndef B2_V2_BACKOFF_MAX
#define B2_V2_BACKOFF_MAX 128u
#endif
#ifndef B2_V2_BUCKET_STAGES
#define B2_V2_BUCKET_STAGES 3  // probe warps request bucket groups 2 tiles ahead
#endif
#ifndef B2_V2_PREFETCH_GROUPS
#define B2_V2_PREFETCH_GROUPS 0  // phase A starts every point's bucket group towards L2 (only useful with B2_V2_PIPELINE_A)
#endif
#ifndef B2_V2_PIPELINE_A
#define B2_V2_PIPELINE_A 0  // probe warps compute tile j + 1's hashes while tile j's bucket groups are in flight (costs registers)
#endif
#ifndef B2_V2_WARM_L2
#define B2_V2_WARM_L2 1  // idle warps of the producer warpgroup stream the bucket table into L2 at kernel start
#endif
#ifndef B2_V2_REGS_AUX
#define B2_V2_REGS_AUX 24
#endif
#include "b2_factor_kernel_v2.cuh"
#undef B2_V2_NAMESPACE

// VGICP as two launches: streaming correspondence search + persistent accumulate-only kernel (B2_VGICP_IMPL == 3)
#include "b2_factor_kernel_split.cuh"
// VGICP, single-role form: every warp runs the whole chain as a cp.async software pipeline (B2_VGICP_IMPL == 4)
#include "b2_factor_kernel_sr.cuh"

// GICP (kd-tree 1-NN): the tree walk is ~100 dependent loads per point, the accumulate work is unchanged -> many thin
// probe warps (the walk keeps its stack in local memory and needs few registers) feeding 4 fat accumulate warps.
#define B2_WS_NAMESPACE ws_gicp
#ifndef B2_WS_COORDS_AHEAD
#define B2_WS_COORDS_AHEAD 0
#endif
#ifndef B2_WS_PREFETCH_OPERANDS
#define B2_WS_PREFETCH_OPERANDS 1
#endif
#ifndef B2_WS_GICP_PRODUCERS
#define B2_WS_GICP_PRODUCERS 28  // (with B2_WS_GICP_SMEM_STACK at most 24: rings 96 KB + traversal stacks 99 KB)
#endif
#undef B2_WS_KD_SMEM_STACK
#ifndef B2_WS_GICP_SMEM_STACK
#define B2_WS_GICP_SMEM_STACK 0  // traversal stack in shared memory instead of local memory: measured slower (479 vs 412 us on cfg3, 24 x 56 vs 28 x 48 registers)
#endif
#define B2_WS_KD_SMEM_STACK B2_WS_GICP_SMEM_STACK
#ifndef B2_WS_GICP_CONSUMERS
#define B2_WS_GICP_CONSUMERS 4
#endif
#ifndef B2_WS_GICP_REGS_PRODUCER
#define B2_WS_GICP_REGS_PRODUCER 48
#endif
#ifndef B2_WS_GICP_REGS_CONSUMER
#define B2_WS_GICP_REGS_CONSUMER 168
#endif
#ifndef B2_WS_GICP_RING
#define B2_WS_GICP_RING 128
#endif
#define B2_WS_PRODUCERS B2_WS_GICP_PRODUCERS
#define B2_WS_CONSUMERS B2_WS_GICP_CONSUMERS
#define B2_WS_REGS_PRODUCER B2_WS_GICP_REGS_PRODUCER
#define B2_WS_REGS_CONSUMER B2_WS_GICP_REGS_CONSUMER
#define B2_WS_RING B2_WS_GICP_RING
#define B2_WS_PPL 1
#ifndef B2_WS_GICP_IPL
#define B2_WS_GICP_IPL 1
#endif
#define B2_WS_IPL B2_WS_GICP_IPL
#include "b2_factor_kernel_ws.cuh"
#undef B2_WS_NAMESPACE

namespace b2 {

// ---------------------------------------------------------------------------------------------------------------
// Host side: factor / factor-set objects
// ---------------------------------------------------------------------------------------------------------------
using KernelFn = void (*)(const FactorDesc*, const uint32_t*, uint32_t, const double*, const double*, double*, unsigned int*, double*, DoneSignal, PoseArg, const uint32_t*);

#if B2_VGICP_IMPL == 1
namespace vgicp = ws;
template <int MODE, bool SINGLE>
KernelFn pick_vgicp(int pb, int cb) {
  if (pb == 4 && cb == 4) return ws::factor_kernel<float, float, 0, MODE, SINGLE>;
  if (pb == 4 && cb == 8) return ws::factor_kernel<float, double, 0, MODE, SINGLE>;
  if (pb == 8 && cb == 4) return ws::factor_kernel<double, float, 0, MODE, SINGLE>;
  return ws::factor_kernel<double, double, 0, MODE, SINGLE>;
}
template <int MODE>
size_t vgicp_smem(int, int) {
  return ws::kDynSmemBytes;
}
#elif B2_VGICP_IMPL == 4
namespace vgicp = sr;
template <int MODE, bool SINGLE>
KernelFn pick_vgicp(int pb, int cb) {
  if (pb == 4 && cb == 4) return sr::factor_kernel<float, float, 0, MODE, SINGLE>;
  if (pb == 4 && cb == 8) return sr::factor_kernel<float, double, 0, MODE, SINGLE>;
  if (pb == 8 && cb == 4) return sr::factor_kernel<double, float, 0, MODE, SINGLE>;
  return ws::factor_kernel<double, double, 0, MODE, SINGLE>;
}
template <int MODE>
size_t vgicp_smem(int pb, int cb) {
  if (pb == 4 && cb == 4) return sr::Layout<float, float, MODE>::kTotal;
  if (pb == 4 && cb == 8) return sr::Layout<float, double, MODE>::kTotal;
  if (pb == 8 && cb == 4) return sr::Layout<double, float, MODE>::kTotal;
  return ws::kDynSmemBytes;
}
#elif B2_VGICP_IMPL == 3
namespace vgicp = sp;
template <int MODE, bool SINGLE>
KernelFn pick_vgicp(int pb, int cb) {
  if (pb == 4 && cb == 4) return sp::factor_kernel<float, float, 0, MODE, SINGLE>;
  if (pb == 4 && cb == 8) return sp::factor_kernel<float, double, 0, MODE, SINGLE>;
  if (pb == 8 && cb == 4) return sp::factor_kernel<double, float, 0, MODE, SINGLE>;
  return sp::factor_kernel<double, double, 0, MODE, SINGLE>;
}
template <int MODE>
size_t vgicp_smem(int pb, int cb) {
  if (pb == 4 && cb == 4) return sp::StageLayout<float, float>::kTotal;
  if (pb == 4 && cb == 8) return sp::StageLayout<float, double>::kTotal;
  if (pb == 8 && cb == 4) return sp::StageLayout<double, float>::kTotal;
  return sp::StageLayout<double, double>::kTotal;
}
#else
namespace vgicp = v2;
template <int MODE, bool SINGLE>
KernelFn pick_vgicp(int pb, int cb) {
  if (pb == 4 && cb == 4) return v2::factor_kernel<float, float, 0, MODE, SINGLE>;
  if (pb == 4 && cb == 8) return v2::factor_kernel<float, double, 0, MODE, SINGLE>;
  if (pb == 8 && cb == 4) return v2::factor_kernel<double, float, 0, MODE, SINGLE>;
  return v2::factor_kernel<double, double, 0, MODE, SINGLE>;
}
template <int MODE>
size_t vgicp_smem(int pb, int cb) {
  if (pb == 4 && cb == 4) return v2::Layout<float, float, MODE>::kTotal;
  if (pb == 4 && cb == 8) return v2::Layout<float, double, MODE>::kTotal;
  if (pb == 8 && cb == 4) return v2::Layout<double, float, MODE>::kTotal;
  return v2::Layout<double, double, MODE>::kTotal;
}
#endif
template <int MODE, bool SINGLE>
KernelFn pick_gicp(int pb, int cb) {
  if (pb == 4 && cb == 4) return ws_gicp::factor_kernel<float, float, 1, MODE, SINGLE>;
  if (pb == 4 && cb == 8) return ws_gicp::factor_kernel<float, double, 1, MODE, SINGLE>;
  if (pb == 8 && cb == 4) return ws_gicp::factor_kernel<double, float, 1, MODE, SINGLE>;
  return ws_gicp::factor_kernel<double, double, 1, MODE, SINGLE>;
}
// ICP / point-to-plane ICP: the same kd-tree kernel with M = I / diag(n^2); no source covariance is read (CT is a dummy)
template <int MODE>
KernelFn pick_icp(int kind, int pb) {
  if (kind == B2_FACTOR_ICP) return pb == 4 ? ws_gicp::factor_kernel<float, float, 2, MODE> : ws_gicp::factor_kernel<double, float, 2, MODE>;
  return pb == 4 ? ws_gicp::factor_kernel<float, float, 3, MODE> : ws_gicp::factor_kernel<double, float, 3, MODE>;
}

// correspondence-search kernel of the two-launch form (nullptr: the factor kernel searches itself)
using ProbeFn = void (*)(const FactorDesc*, const uint32_t*, const double*, PoseArg, const uint32_t*);
ProbeFn pick_probe(int kind, int pb, bool single) {
#if B2_VGICP_IMPL == 3
  if (kind == 0) {
    if (single) return pb == 4 ? sp::probe_kernel<float, true> : sp::probe_kernel<double, true>;
    return pb == 4 ? sp::probe_kernel<float, false> : sp::probe_kernel<double, false>;
  }
#endif
  (void)kind, (void)pb, (void)single;
  return nullptr;
}

// launch shape of a kernel configuration
struct KernelShape {
  int threads, tile;
};
#if B2_VGICP_IMPL == 4
// float64 point AND covariance storage does not fit the single-role kernel's shared-memory slots: that (rare) storage mode
// keeps the warp-specialised kernel
inline bool vgicp_uses_ws(int pb, int cb) { return pb == 8 && cb == 8; }
#else
inline bool vgicp_uses_ws(int, int) { return false; }
#endif
KernelShape kernel_shape(int kind, int pb, int cb) {
  if (kind == 0) return vgicp_uses_ws(pb, cb) ? KernelShape{ws::kThreads, ws::kTile} : KernelShape{vgicp::kThreads, vgicp::kTile};
  return {ws_gicp::kThreads, ws_gicp::kTile};
}
size_t kernel_smem(int kind, int mode, int pb, int cb) {
  if (kind == 0) return mode == MODE_LINEARIZE ? vgicp_smem<MODE_LINEARIZE>(pb, cb) : vgicp_smem<MODE_ERROR>(pb, cb);
  return ws_gicp::kDynSmemBytes;
}

// single == true: the by-value-pose instantiation for launches that cover exactly one factor (VGICP kernel only)
KernelFn pick_kernel(int kind, int mode, int pb, int cb, bool single) {
  if (kind == 0) {
    if (single) return mode == MODE_LINEARIZE ? pick_vgicp<MODE_LINEARIZE, true>(pb, cb) : pick_vgicp<MODE_ERROR, true>(pb, cb);
    return mode == MODE_LINEARIZE ? pick_vgicp<MODE_LINEARIZE, false>(pb, cb) : pick_vgicp<MODE_ERROR, false>(pb, cb);
  }
  if (kind >= B2_FACTOR_ICP) return single ? nullptr : (mode == MODE_LINEARIZE ? pick_icp<MODE_LINEARIZE>(kind, pb) : pick_icp<MODE_ERROR>(kind, pb));
  if (single) return mode == MODE_LINEARIZE ? pick_gicp<MODE_LINEARIZE, true>(pb, cb) : pick_gicp<MODE_ERROR, true>(pb, cb);
  return mode == MODE_LINEARIZE ? pick_gicp<MODE_LINEARIZE, false>(pb, cb) : pick_gicp<MODE_ERROR, false>(pb, cb);
}

// host-API sets up to this many factors use the zero-copy path (poses read from / results written to mapped pinned memory)
constexpr size_t kZeroCopyMaxFactors = 64;

struct Group {
  int kind, pb, cb;
  std::vector<size_t> members;  // indices into the set's factor list
  std::vector<uint64_t> gen;    // per member: the factor's params_gen its device descriptor was built from
  std::vector<uint64_t> vm_gen; // per member: generation of its voxel map (device pointers / ids change on insert())
  std::vector<FactorDesc> h_descs;  // host copy of the device descriptors (patched and re-uploaded when a generation moves)
  FactorDesc* d_descs = nullptr;
  uint32_t* d_tile_factor = nullptr;
  uint32_t num_tiles = 0;
  uint32_t grid[2] = {0, 0};
  KernelFn fn[2] = {nullptr, nullptr};
  KernelFn fn_single[2] = {nullptr, nullptr};  // by-value-pose instantiation (only for sets of exactly one factor)
  ProbeFn probe = nullptr, probe_single = nullptr;  // two-launch form: correspondence search before a linearize
  size_t dyn_smem[2] = {0, 0};
};

}  // namespace b2

using namespace b2;

struct b2_factor_set {
  b2_ctx* ctx = nullptr;
  std::vector<b2_factor*> factors;
  std::vector<Group> groups;
  double* d_partials = nullptr;
  unsigned int* d_counters = nullptr;
  double* d_poses_lin = nullptr;   // F x 16
  double* d_poses_eval = nullptr;  // F x 16
  double* d_out = nullptr;         // F x 128
  double* d_err = nullptr;         // F
  uint32_t* d_frozen = nullptr;    // F: per factor, 1 = keep the stored correspondences in this linearize (correspondence-update tolerance)
  bool any_tolerance = false;      // some factor of the set has a tolerance: the flags are maintained
  uint64_t launches = 0;
};

namespace {

__global__ void scatter_corr_kernel(const int32_t* __restrict__ corr, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ leaf_index, size_t n,
                                    long long* __restrict__ out) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int c = corr[i];
  const long long v = c < 0 ? -1ll : (leaf_index ? static_cast<long long>(leaf_index[c]) : static_cast<long long>(c));
  out[perm ? perm[i] : i] = v;
}

// GICP: target records (mean, cov, 1) in the tree's leaf order, gathered on the device from the target cloud's planes.
template <typename PT, typename CT>
__global__ void build_target_records_kernel(const PT* __restrict__ pts, const CT* __restrict__ covs, size_t n_pad, const uint32_t* __restrict__ cloud_inv_perm,
                                            const uint32_t* __restrict__ leaf_index, size_t n, double* __restrict__ records) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  const uint32_t caller = leaf_index[j];
  const size_t s = cloud_inv_perm ? cloud_inv_perm[caller] : caller;
  double* r = records + j * kRecordDoubles;
  r[0] = static_cast<double>(pts[s]);
  r[1] = static_cast<double>(pts[n_pad + s]);
  r[2] = static_cast<double>(pts[2 * n_pad + s]);
#pragma unroll
  for (int k = 0; k < 6; k++) r[3 + k] = static_cast<double>(covs[k * n_pad + s]);
  r[9] = 1.0;
}

// ICP: target records (mean | normal or zeros | 1) in the tree's leaf order; normals are a host array in caller order (n x 3)
template <typename PT>
__global__ void build_icp_records_kernel(const PT* __restrict__ pts, size_t n_pad, const uint32_t* __restrict__ cloud_inv_perm, const uint32_t* __restrict__ leaf_index,
                                         const double* __restrict__ normals, size_t n, double* __restrict__ records) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  const uint32_t caller = leaf_index[j];
  const size_t s = cloud_inv_perm ? cloud_inv_perm[caller] : caller;
  double* r = records + j * kRecordDoubles;
  r[0] = static_cast<double>(pts[s]);
  r[1] = static_cast<double>(pts[n_pad + s]);
  r[2] = static_cast<double>(pts[2 * n_pad + s]);
#pragma unroll
  for (int k = 0; k < 3; k++) r[3 + k] = normals ? normals[static_cast<size_t>(caller) * 3 + k] : 0.0;
  r[6] = r[7] = r[8] = 0.0;
  r[9] = 1.0;
}

__global__ void invert_perm_kernel(const uint32_t* __restrict__ perm, size_t n, uint32_t* __restrict__ inv) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) inv[perm[i]] = static_cast<uint32_t>(i);
}

// Wait for the completion word of a zero-copy call (see DoneSignal).  The stream is polled every 1024 spins so that a
// failed launch / faulting kernel is reported instead of spinning forever.
b2_status wait_done(b2_ctx* ctx, unsigned int seq) {
  const volatile unsigned int* flag = ctx->h_done;
  unsigned spins = 0;
  while (*flag != seq) {
    if ((++spins & 0x3FFu) == 0u) {
      const cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q == cudaSuccess) break;  // everything on the stream has finished: the results are in place
      if (q != cudaErrorNotReady) return fail(B2_ERR_CUDA, "kernel failed: %s", cudaGetErrorString(q));
    }
    __builtin_ia32_pause();
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return B2_OK;
}

// h_pose != nullptr: the set holds ONE factor and its pose (16 doubles, host memory) travels by value with the launch
// Eigen::AngleAxisd(R).angle() the way Eigen computes it (through the quaternion: 2 atan2(|vec|, |w|)), for the tolerance test
double rotation_angle_rm(const double* T) {
  const double m00 = T[0], m11 = T[5], m22 = T[10];
  double w, q[3];
  const double tr = m00 + m11 + m22;
  if (tr > 0.0) {
    double t = std::sqrt(tr + 1.0);
    w = 0.5 * t;
    t = 0.5 / t;
    q[0] = (T[9] - T[6]) * t, q[1] = (T[2] - T[8]) * t, q[2] = (T[4] - T[1]) * t;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > T[i * 5]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(T[i * 5] - T[j * 5] - T[k * 5] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    w = (T[k * 4 + j] - T[j * 4 + k]) * t;
    q[j] = (T[j * 4 + i] + T[i * 4 + j]) * t;
    q[k] = (T[k * 4 + i] + T[i * 4 + k]) * t;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  return n != 0.0 ? 2.0 * std::atan2(n, std::fabs(w)) : 0.0;
}

// integrated_gicp_factor_impl.hpp:135-147 for every factor of a host-pose linearize: decides which factors keep their
// correspondences (pose within the tolerances of the last association point), remembers the new association points, and
// uploads the flags (stream-ordered before the launch).  Returns the device flag array, or nullptr if no factor has a tolerance.
b2_status prepare_frozen_flags(b2_factor_set* s, const double* deltas, const uint32_t** out_flags) {
  *out_flags = nullptr;
  if (!s->any_tolerance) {
    for (size_t i = 0; i < s->factors.size(); i++) {
      std::memcpy(s->factors[i]->last_corr_delta, deltas + i * 16, sizeof(double) * 16);
      s->factors[i]->has_corr = true;
    }
    return B2_OK;
  }
  const size_t F = s->factors.size();
  std::vector<uint32_t> flags(F, 0u);
  for (size_t i = 0; i < F; i++) {
    b2_factor* f = s->factors[i];
    const double* d = deltas + i * 16;
    bool do_update = true;
    if (f->kind != B2_FACTOR_VGICP && f->has_corr && (f->corr_tol_trans > 0.0 || f->corr_tol_rot > 0.0)) {
      // diff = delta^-1 * last_correspondence_point
      double diff[16] = {0};
      const double* L = f->last_corr_delta;
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) diff[r * 4 + c] = d[0 * 4 + r] * L[0 * 4 + c] + d[1 * 4 + r] * L[1 * 4 + c] + d[2 * 4 + r] * L[2 * 4 + c];
        diff[r * 4 + 3] = d[0 * 4 + r] * (L[3] - d[3]) + d[1 * 4 + r] * (L[7] - d[7]) + d[2 * 4 + r] * (L[11] - d[11]);
      }
      const double diff_rot = rotation_angle_rm(diff);
      const double diff_trans = std::sqrt(diff[3] * diff[3] + diff[7] * diff[7] + diff[11] * diff[11]);
      if (diff_rot < f->corr_tol_rot && diff_trans < f->corr_tol_trans) do_update = false;
    }
    if (do_update) {
      std::memcpy(f->last_corr_delta, d, sizeof(double) * 16);
      f->has_corr = true;
    } else {
      flags[i] = 1u;
    }
  }
  B2_CUDA(cudaMemcpyAsync(s->d_frozen, flags.data(), F * sizeof(uint32_t), cudaMemcpyHostToDevice, s->ctx->stream));  // pageable source: staged before return
  *out_flags = s->d_frozen;
  return B2_OK;
}

b2_status launch_groups(b2_factor_set* s, int mode, const double* d_lin, const double* d_eval, double* d_out, DoneSignal sig = DoneSignal{}, const double* h_pose = nullptr,
                        const uint32_t* d_frozen = nullptr) {
  cudaStream_t st = s->ctx->stream;
  PoseArg pa{};
  const bool single = h_pose != nullptr && s->factors.size() == 1 && s->groups.size() == 1 && s->groups[0].fn_single[mode] != nullptr;
  if (single) {
    std::memcpy(pa.m, h_pose, sizeof(pa.m));
  }
  for (auto& g : s->groups) {
    // tuning setters called after the set was built (set_max_correspondence_distance) and voxel maps that grew since
    // (b2_voxelmap_insert): refresh the device descriptors, stream-ordered before the launch -- like the reference, a new
    // value takes effect at the next correspondence update
    for (size_t k = 0; k < g.members.size(); k++) {
      const b2_factor* f = s->factors[g.members[k]];
      const uint64_t vgen = f->voxelmap ? f->voxelmap->generation : 0;
      if (f->params_gen != g.gen[k] || vgen != g.vm_gen[k]) {
        FactorDesc& d = g.h_descs[k];
        d.max_sq = f->max_corr_sq;
        if (f->voxelmap) {
          d.buckets = f->voxelmap->d_buckets;
          d.bucket_mask = static_cast<uint32_t>(f->voxelmap->num_buckets / kGroup - 1);
          d.inv_leaf = f->voxelmap->inv_resolution;
          d.records = f->voxelmap->d_records;
          d.num_records = static_cast<uint32_t>(f->voxelmap->num_voxels);
        }
        B2_CUDA(cudaMemcpyAsync(&g.d_descs[k], &d, sizeof(FactorDesc), cudaMemcpyHostToDevice, st));
        g.gen[k] = f->params_gen;
        g.vm_gen[k] = vgen;
      }
    }
    if (single) pa.desc = g.h_descs[0];  // after the refresh above: the by-value copy the kernel works from
    if (mode == MODE_LINEARIZE && g.probe != nullptr) {
      (single ? g.probe_single : g.probe)<<<g.num_tiles, kernel_shape(g.kind, g.pb, g.cb).tile, 0, st>>>(g.d_descs, g.d_tile_factor, d_lin, pa, d_frozen);
      s->launches++;
    }
    (single ? g.fn_single[mode] : g.fn[mode])<<<g.grid[mode], kernel_shape(g.kind, g.pb, g.cb).threads, g.dyn_smem[mode], st>>>(g.d_descs, g.d_tile_factor, g.num_tiles, d_lin, d_eval, s->d_partials,
                                                                                                                  s->d_counters, d_out, sig, pa, d_frozen);
    s->launches++;
  }
  B2_CUDA(cudaGetLastError());
  return B2_OK;
}

}  // namespace

extern "C" {

b2_status b2_vgicp_factor_create(b2_ctx* ctx, const b2_voxelmap* target, const b2_cloud* source, b2_factor** out) {
  B2_REQUIRE(out != nullptr, "b2_vgicp_factor_create: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_vgicp_factor_create: ctx is NULL");
  // reference aborts on these (integrated_vgicp_factor_impl.hpp:32-45); the ABI reports them instead
  B2_REQUIRE(source != nullptr && source->d_points != nullptr, "error: source points have not been allocated!!");
  B2_REQUIRE(source->d_covs != nullptr, "error: source don't have covs!!");
  B2_REQUIRE(target != nullptr, "error: target voxelmap has not been created!!");
  B2_REQUIRE(target->ctx->device == ctx->device && source->ctx->device == ctx->device, "b2_vgicp_factor_create: handles live on different devices");
  B2_CUDA(cudaSetDevice(ctx->device));
  b2_factor* f = new b2_factor;
  f->ctx = ctx;
  f->kind = B2_FACTOR_VGICP;
  f->voxelmap = target;
  f->source = source;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&f->d_corr), std::max<size_t>(source->n_pad, 1) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&f->d_lin_pose), 16 * sizeof(double));
  if (e != cudaSuccess) {
    b2_factor_destroy(f);
    return fail(B2_ERR_OUT_OF_MEMORY, "b2_vgicp_factor_create: %s", cudaGetErrorString(e));
  }
  *out = f;
  return B2_OK;
}

b2_status b2_gicp_factor_create(b2_ctx* ctx, const b2_cloud* target_cloud, const b2_kdtree* tree, const b2_cloud* source, b2_factor** out) {
  B2_REQUIRE(out != nullptr, "b2_gicp_factor_create: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_gicp_factor_create: ctx is NULL");
  // reference aborts on these (integrated_gicp_factor_impl.hpp:37-65)
  B2_REQUIRE(source != nullptr && source->d_points != nullptr, "error: source points have not been allocated!!");
  B2_REQUIRE(source->d_covs != nullptr, "error: source don't have covs!!");
  B2_REQUIRE(target_cloud != nullptr && target_cloud->d_points != nullptr, "error: target points have not been allocated!!");
  B2_REQUIRE(target_cloud->d_covs != nullptr, "error: target don't have covs!!");
  B2_REQUIRE(tree != nullptr, "b2_gicp_factor_create: tree is NULL");
  B2_REQUIRE(tree->n == target_cloud->n, "b2_gicp_factor_create: tree (%zu points) was not built over the target cloud (%zu points)", tree->n, target_cloud->n);
  B2_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  b2_factor* f = new b2_factor;
  f->ctx = ctx;
  f->kind = B2_FACTOR_GICP;
  f->target = target_cloud;
  f->tree = tree;
  f->source = source;
  const size_t nt = tree->n;
  cudaError_t e;
  uint32_t* d_inv = nullptr;
  if ((e = cudaMalloc(reinterpret_cast<void**>(&f->d_corr), std::max<size_t>(source->n_pad, 1) * sizeof(int32_t))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&f->d_lin_pose), 16 * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&f->d_target_records), std::max<size_t>(nt, 1) * kRecordDoubles * sizeof(double))) != cudaSuccess) {
    b2_factor_destroy(f);
    return fail(B2_ERR_OUT_OF_MEMORY, "b2_gicp_factor_create: %s", cudaGetErrorString(e));
  }
  if (nt > 0) {
    const unsigned grid = static_cast<unsigned>((nt + 255) / 256);
    if (target_cloud->d_perm) {
      if ((e = cudaMalloc(reinterpret_cast<void**>(&d_inv), nt * sizeof(uint32_t))) != cudaSuccess) {
        b2_factor_destroy(f);
        return fail(B2_ERR_OUT_OF_MEMORY, "b2_gicp_factor_create: %s", cudaGetErrorString(e));
      }
      invert_perm_kernel<<<grid, 256, 0, st>>>(target_cloud->d_perm, nt, d_inv);
    }
    const size_t np = target_cloud->n_pad;
    if (target_cloud->point_bytes == 4 && target_cloud->cov_bytes == 4)
      build_target_records_kernel<float, float><<<grid, 256, 0, st>>>(static_cast<const float*>(target_cloud->d_points), static_cast<const float*>(target_cloud->d_covs), np, d_inv, tree->d_leaf_index, nt, f->d_target_records);
    else if (target_cloud->point_bytes == 4)
      build_target_records_kernel<float, double><<<grid, 256, 0, st>>>(static_cast<const float*>(target_cloud->d_points), static_cast<const double*>(target_cloud->d_covs), np, d_inv, tree->d_leaf_index, nt, f->d_target_records);
    else if (target_cloud->cov_bytes == 4)
      build_target_records_kernel<double, float><<<grid, 256, 0, st>>>(static_cast<const double*>(target_cloud->d_points), static_cast<const float*>(target_cloud->d_covs), np, d_inv, tree->d_leaf_index, nt, f->d_target_records);
    else
      build_target_records_kernel<double, double><<<grid, 256, 0, st>>>(static_cast<const double*>(target_cloud->d_points), static_cast<const double*>(target_cloud->d_covs), np, d_inv, tree->d_leaf_index, nt, f->d_target_records);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (d_inv) cudaFree(d_inv);
    if (e != cudaSuccess) {
      b2_factor_destroy(f);
      return fail(B2_ERR_CUDA, "b2_gicp_factor_create: %s", cudaGetErrorString(e));
    }
  }
  *out = f;
  return B2_OK;
}

b2_status b2_icp_factor_create(b2_ctx* ctx, const b2_cloud* target_cloud, const b2_kdtree* tree, const b2_cloud* source, int use_point_to_plane, const double* target_normals,
                               b2_factor** out) {
  B2_REQUIRE(out != nullptr, "b2_icp_factor_create: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_icp_factor_create: ctx is NULL");
  // reference aborts on these (integrated_icp_factor_impl.hpp:37-45)
  B2_REQUIRE(target_cloud != nullptr && target_cloud->d_points != nullptr && (!use_point_to_plane || target_normals != nullptr), "error: target frame doesn't have required attributes for icp");
  B2_REQUIRE(source != nullptr && source->d_points != nullptr, "error: source frame doesn't have required attributes for icp");
  B2_REQUIRE(tree != nullptr && tree->n == target_cloud->n, "b2_icp_factor_create: the tree was not built over the target cloud");
  B2_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  b2_factor* f = new b2_factor;
  f->ctx = ctx;
  f->kind = use_point_to_plane ? B2_FACTOR_ICP_PLANE : B2_FACTOR_ICP;
  f->target = target_cloud;
  f->tree = tree;
  f->source = source;
  const size_t nt = tree->n;
  cudaError_t e;
  uint32_t* d_inv = nullptr;
  double* d_normals = nullptr;
  auto bail = [&](b2_status stt) {
    if (d_inv) cudaFree(d_inv);
    if (d_normals) cudaFree(d_normals);
    b2_factor_destroy(f);
    return stt;
  };
  if ((e = cudaMalloc(reinterpret_cast<void**>(&f->d_corr), std::max<size_t>(source->n_pad, 1) * sizeof(int32_t))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&f->d_lin_pose), 16 * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&f->d_target_records), std::max<size_t>(nt, 1) * kRecordDoubles * sizeof(double))) != cudaSuccess) {
    return bail(fail(B2_ERR_OUT_OF_MEMORY, "b2_icp_factor_create: %s", cudaGetErrorString(e)));
  }
  if (nt > 0) {
    const unsigned grid = static_cast<unsigned>((nt + 255) / 256);
    if (target_cloud->d_perm) {
      if ((e = cudaMalloc(reinterpret_cast<void**>(&d_inv), nt * sizeof(uint32_t))) != cudaSuccess) return bail(fail(B2_ERR_OUT_OF_MEMORY, "b2_icp_factor_create: %s", cudaGetErrorString(e)));
      invert_perm_kernel<<<grid, 256, 0, st>>>(target_cloud->d_perm, nt, d_inv);
    }
    if (use_point_to_plane) {
      if ((e = cudaMalloc(reinterpret_cast<void**>(&d_normals), nt * 3 * sizeof(double))) != cudaSuccess ||
          (e = cudaMemcpyAsync(d_normals, target_normals, nt * 3 * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return bail(fail(B2_ERR_OUT_OF_MEMORY, "b2_icp_factor_create: %s", cudaGetErrorString(e)));
    }
    if (target_cloud->point_bytes == 4)
      build_icp_records_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(target_cloud->d_points), target_cloud->n_pad, d_inv, tree->d_leaf_index, d_normals, nt, f->d_target_records);
    else
      build_icp_records_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(target_cloud->d_points), target_cloud->n_pad, d_inv, tree->d_leaf_index, d_normals, nt, f->d_target_records);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(fail(B2_ERR_CUDA, "b2_icp_factor_create: %s", cudaGetErrorString(e)));
  }
  if (d_inv) cudaFree(d_inv);
  if (d_normals) cudaFree(d_normals);
  *out = f;
  return B2_OK;
}

b2_status b2_factor_destroy(b2_factor* f) {
  if (!f) return B2_OK;
  cudaSetDevice(f->ctx->device);
  if (f->self_set) b2_factor_set_destroy(f->self_set);
  if (f->d_corr) cudaFree(f->d_corr);
  if (f->d_target_records) cudaFree(f->d_target_records);
  if (f->d_lin_pose) cudaFree(f->d_lin_pose);
  delete f;
  return B2_OK;
}

b2_status b2_factor_set_max_correspondence_distance(b2_factor* f, double dist) {
  B2_REQUIRE(f != nullptr, "b2_factor_set_max_correspondence_distance: factor is NULL");
  B2_REQUIRE(dist >= 0.0, "b2_factor_set_max_correspondence_distance: negative distance");
  f->max_corr_sq = dist * dist;  // integrated_gicp_factor.hpp:98-101
  f->params_gen++;               // every factor set holding this factor refreshes its device descriptor before its next launch
  return B2_OK;
}

b2_status b2_factor_set_correspondence_update_tolerance(b2_factor* f, double angle, double trans) {
  B2_REQUIRE(f != nullptr, "b2_factor_set_correspondence_update_tolerance: factor is NULL");
  B2_REQUIRE(f->kind != B2_FACTOR_VGICP, "b2_factor_set_correspondence_update_tolerance: the VGICP factor has no correspondence-update tolerance (integrated_vgicp_factor.hpp)");
  B2_REQUIRE(angle >= 0.0 && trans >= 0.0, "b2_factor_set_correspondence_update_tolerance: negative tolerance");
  f->corr_tol_rot = angle;  // integrated_gicp_factor.hpp:106-109
  f->corr_tol_trans = trans;
  if (f->self_set) f->self_set->any_tolerance = true;
  return B2_OK;
}

size_t b2_factor_num_points(const b2_factor* f) { return f ? f->source->n : 0; }

b2_status b2_factor_correspondences(const b2_factor* f, int64_t* out) {
  B2_REQUIRE(f && out, "b2_factor_correspondences: NULL argument");
  const size_t n = f->source->n;
  if (n == 0) return B2_OK;
  if (!f->linearized) {
    for (size_t i = 0; i < n; i++) out[i] = -1;
    return B2_OK;
  }
  B2_CUDA(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  long long* d_out = nullptr;
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_out), n * sizeof(long long)));
  scatter_corr_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(f->d_corr, f->source->d_perm, f->kind != B2_FACTOR_VGICP ? f->tree->d_leaf_index : nullptr, n, d_out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * sizeof(long long), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(B2_ERR_CUDA, "b2_factor_correspondences: %s", cudaGetErrorString(e));
  return B2_OK;
}

// ---- factor sets ------------------------------------------------------------------------------------------------

b2_status b2_factor_set_create(b2_ctx* ctx, b2_factor* const* factors, size_t F, b2_factor_set** out) {
  B2_REQUIRE(out != nullptr, "b2_factor_set_create: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_factor_set_create: ctx is NULL");
  B2_REQUIRE(F > 0 && factors != nullptr, "b2_factor_set_create: empty factor list");
  for (size_t i = 0; i < F; i++) {
    B2_REQUIRE(factors[i] != nullptr, "b2_factor_set_create: factor %zu is NULL", i);
    B2_REQUIRE(factors[i]->ctx->device == ctx->device, "b2_factor_set_create: factor %zu lives on another device", i);
  }
  B2_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;

  b2_factor_set* s = new b2_factor_set;
  s->ctx = ctx;
  s->factors.assign(factors, factors + F);

  std::map<std::tuple<int, int, int>, size_t> group_of;
  for (size_t i = 0; i < F; i++) {
    const b2_factor* f = factors[i];
    const auto key = std::make_tuple(static_cast<int>(f->kind), f->source->point_bytes, f->kind >= B2_FACTOR_ICP ? 4 : f->source->cov_bytes);
    auto it = group_of.find(key);
    if (it == group_of.end()) {
      it = group_of.emplace(key, s->groups.size()).first;
      Group g;
      g.kind = f->kind;
      g.pb = f->source->point_bytes;
      g.cb = f->kind >= B2_FACTOR_ICP ? 4 : f->source->cov_bytes;
      s->groups.push_back(g);
    }
    s->groups[it->second].members.push_back(i);
  }

  auto fail_cleanup = [&](b2_status stt) {
    b2_factor_set_destroy(s);
    return stt;
  };

  uint32_t slot_cursor = 0;
  for (auto& g : s->groups) {
    const KernelShape shape = kernel_shape(g.kind, g.pb, g.cb);
    std::vector<FactorDesc> descs(g.members.size());
    std::vector<uint32_t> tile_factor;
    uint32_t tile_cursor = 0;
    for (size_t k = 0; k < g.members.size(); k++) {
      const b2_factor* f = factors[g.members[k]];
      FactorDesc& d = descs[k];
      std::memset(&d, 0, sizeof(d));
      d.pts = f->source->d_points;
      d.covs = f->source->d_covs;
      d.n = static_cast<uint32_t>(f->source->n);
      d.n_pad = static_cast<uint32_t>(f->source->n_pad);
      if (f->kind == B2_FACTOR_VGICP) {
        d.buckets = f->voxelmap->d_buckets;
        d.bucket_mask = static_cast<uint32_t>(f->voxelmap->num_buckets / kGroup - 1);  // group mask
        d.inv_leaf = f->voxelmap->inv_resolution;
        d.records = f->voxelmap->d_records;
        d.num_records = static_cast<uint32_t>(f->voxelmap->num_voxels);
      } else {
        d.nodes = f->tree->d_nodes;
        d.leaf_pts = f->tree->d_leaf_points;
        d.leaf_f32 = f->tree->leaf_f32 ? 1u : 0u;
        d.max_sq = f->max_corr_sq;
        d.records = f->d_target_records;
      }
      d.corr = f->d_corr;
      d.lin_pose = f->d_lin_pose;
      d.tile_begin = tile_cursor;
      d.num_tiles = std::max<uint32_t>(1u, (d.n + shape.tile - 1) / shape.tile);
      d.perm_stride = golden_stride(d.num_tiles);
      d.out_index = static_cast<uint32_t>(g.members[k]);
      g.gen.push_back(f->params_gen);
      g.vm_gen.push_back(f->voxelmap ? f->voxelmap->generation : 0);
      tile_cursor += d.num_tiles;
      tile_factor.insert(tile_factor.end(), d.num_tiles, static_cast<uint32_t>(k));
    }
    g.num_tiles = tile_cursor;
    auto keep_host_copy = [&]() { g.h_descs = descs; };
    g.probe = pick_probe(g.kind, g.pb, false);
    g.probe_single = F == 1 ? pick_probe(g.kind, g.pb, true) : nullptr;
    for (int mode = 0; mode < 2; mode++) {
      g.fn[mode] = pick_kernel(g.kind, mode, g.pb, g.cb, false);
      g.fn_single[mode] = F == 1 ? pick_kernel(g.kind, mode, g.pb, g.cb, true) : nullptr;  // by-value pose (VGICP, GICP)
      g.dyn_smem[mode] = kernel_smem(g.kind, mode, g.pb, g.cb);
      int per_sm = 0;
      cudaError_t e = cudaSuccess;
      for (KernelFn fn : {g.fn[mode], g.fn_single[mode]}) {
        if (fn == nullptr || e != cudaSuccess) continue;
        e = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g.dyn_smem[mode]));
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, shape.threads, g.dyn_smem[mode]);
        if (e == cudaSuccess && per_sm < 1) e = cudaErrorLaunchOutOfResources;
      }
      if (e != cudaSuccess) return fail_cleanup(fail(B2_ERR_CUDA, "b2_factor_set_create: the kernel does not fit one CTA per SM (%s)", cudaGetErrorString(e)));
      // persistent: one CTA per SM (the kernel redistributes the SM's whole register file between its warp roles)
      const uint32_t G = std::min<uint32_t>(g.num_tiles, static_cast<uint32_t>(ctx->sm_count));
      g.grid[mode] = G;
      // CTA c owns tiles [c T / G, (c + 1) T / G): a factor's partial-sum slots are the CTAs whose range touches it
      uint32_t c = 0;
      for (auto& d : descs) {
        while (vgicp::cta_tile_begin(c + 1, g.num_tiles, G) <= d.tile_begin) c++;
        uint32_t c_last = c;
        while (vgicp::cta_tile_begin(c_last + 1, g.num_tiles, G) < d.tile_begin + d.num_tiles) c_last++;
        d.cta_first[mode] = c;
        d.num_slots[mode] = c_last - c + 1;
        d.slot_begin[mode] = slot_cursor;
        slot_cursor += d.num_slots[mode];
      }
    }
    keep_host_copy();  // after the per-mode slot bookkeeping above has been written into `descs`
    cudaError_t e;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&g.d_descs), descs.size() * sizeof(FactorDesc))) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void**>(&g.d_tile_factor), tile_factor.size() * sizeof(uint32_t))) != cudaSuccess) {
      return fail_cleanup(fail(B2_ERR_OUT_OF_MEMORY, "b2_factor_set_create: %s", cudaGetErrorString(e)));
    }
    if ((e = cudaMemcpyAsync(g.d_descs, descs.data(), descs.size() * sizeof(FactorDesc), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(g.d_tile_factor, tile_factor.data(), tile_factor.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess) {
      return fail_cleanup(fail(B2_ERR_CUDA, "b2_factor_set_create: %s", cudaGetErrorString(e)));
    }
  }

  cudaError_t e;
  if ((e = cudaMalloc(reinterpret_cast<void**>(&s->d_partials), static_cast<size_t>(slot_cursor) * kAcc * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&s->d_counters), F * sizeof(unsigned int))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&s->d_poses_lin), F * 16 * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&s->d_poses_eval), F * 16 * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&s->d_out), F * B2_LINEARIZED_DOUBLES * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&s->d_err), F * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void**>(&s->d_frozen), F * sizeof(uint32_t))) != cudaSuccess) {
    return fail_cleanup(fail(B2_ERR_OUT_OF_MEMORY, "b2_factor_set_create: %s", cudaGetErrorString(e)));
  }
  if ((e = cudaMemsetAsync(s->d_counters, 0, F * sizeof(unsigned int), st)) != cudaSuccess || (e = cudaStreamSynchronize(st)) != cudaSuccess) {
    return fail_cleanup(fail(B2_ERR_CUDA, "b2_factor_set_create: %s", cudaGetErrorString(e)));
  }
  for (size_t i = 0; i < F; i++) s->any_tolerance |= factors[i]->corr_tol_rot > 0.0 || factors[i]->corr_tol_trans > 0.0;
  b2_status ss = ctx->ensure_stage(F * (B2_LINEARIZED_DOUBLES + 32) * sizeof(double), 0);
  if (ss != B2_OK) return fail_cleanup(ss);
  *out = s;
  return B2_OK;
}

b2_status b2_factor_set_destroy(b2_factor_set* s) {
  if (!s) return B2_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  for (auto& g : s->groups) {
    if (g.d_descs) cudaFree(g.d_descs);
    if (g.d_tile_factor) cudaFree(g.d_tile_factor);
  }
  if (s->d_partials) cudaFree(s->d_partials);
  if (s->d_counters) cudaFree(s->d_counters);
  if (s->d_poses_lin) cudaFree(s->d_poses_lin);
  if (s->d_poses_eval) cudaFree(s->d_poses_eval);
  if (s->d_out) cudaFree(s->d_out);
  if (s->d_err) cudaFree(s->d_err);
  if (s->d_frozen) cudaFree(s->d_frozen);
  delete s;
  return B2_OK;
}

size_t b2_factor_set_size(const b2_factor_set* s) { return s ? s->factors.size() : 0; }
uint64_t b2_factor_set_launch_count(const b2_factor_set* s) { return s ? s->launches : 0; }

b2_status b2_factor_set_linearize_device(b2_factor_set* s, const double* d_deltas, double* d_out) {
  B2_REQUIRE(s && d_deltas && d_out, "b2_factor_set_linearize_device: NULL argument");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  const size_t F = s->factors.size();
  (void)F;
  // one launch per group, nothing else: the kernel's per-factor epilogue also stores the linearization point on the device
  B2_TRY(launch_groups(s, MODE_LINEARIZE, d_deltas, d_deltas, d_out));
  for (auto* f : s->factors) f->linearized = true;
  return B2_OK;
}

b2_status b2_factor_set_error_device(b2_factor_set* s, const double* d_deltas_eval, double* d_out_errors) {
  B2_REQUIRE(s && d_deltas_eval && d_out_errors, "b2_factor_set_error_device: NULL argument");
  for (auto* f : s->factors)
    if (!f->linearized) return fail(B2_ERR_INVALID_STATE, "b2_factor_set_error_device: a factor of the set has not been linearized yet");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  B2_TRY(launch_groups(s, MODE_ERROR, nullptr, d_deltas_eval, d_out_errors));
  return B2_OK;
}

b2_status b2_factor_set_issue_linearize(b2_factor_set* s, const double* deltas, double* d_out) {
  B2_REQUIRE(s && deltas, "b2_factor_set_issue_linearize: NULL argument");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  const size_t F = s->factors.size();
  if (d_out == nullptr) d_out = s->d_out;
  const uint32_t* d_frozen = nullptr;
  B2_TRY(prepare_frozen_flags(s, deltas, &d_frozen));
  if (F == 1 && s->groups[0].fn_single[MODE_LINEARIZE] != nullptr) {
    B2_TRY(launch_groups(s, MODE_LINEARIZE, nullptr, nullptr, d_out, DoneSignal{}, deltas));  // pose by value: nothing but the launch
  } else {
    // pageable source: the runtime stages it before returning, so the caller's buffer is free immediately
    B2_CUDA(cudaMemcpyAsync(s->d_poses_lin, deltas, F * 16 * sizeof(double), cudaMemcpyHostToDevice, s->ctx->stream));
    B2_TRY(launch_groups(s, MODE_LINEARIZE, s->d_poses_lin, s->d_poses_lin, d_out, DoneSignal{}, nullptr, d_frozen));
  }
  for (size_t i = 0; i < F; i++) {
    s->factors[i]->linearized = true;
    std::memcpy(s->factors[i]->lin_delta, deltas + i * 16, 16 * sizeof(double));
  }
  return B2_OK;
}

b2_status b2_factor_set_issue_error(b2_factor_set* s, const double* deltas_eval, double* d_out_errors) {
  B2_REQUIRE(s && deltas_eval, "b2_factor_set_issue_error: NULL argument");
  for (auto* f : s->factors)
    if (!f->linearized) return fail(B2_ERR_INVALID_STATE, "b2_factor_set_issue_error: a factor of the set has not been linearized yet");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  const size_t F = s->factors.size();
  if (d_out_errors == nullptr) d_out_errors = s->d_err;
  if (F == 1 && s->groups[0].fn_single[MODE_ERROR] != nullptr) {
    B2_TRY(launch_groups(s, MODE_ERROR, nullptr, nullptr, d_out_errors, DoneSignal{}, deltas_eval));
  } else {
    B2_CUDA(cudaMemcpyAsync(s->d_poses_eval, deltas_eval, F * 16 * sizeof(double), cudaMemcpyHostToDevice, s->ctx->stream));
    B2_TRY(launch_groups(s, MODE_ERROR, nullptr, s->d_poses_eval, d_out_errors));
  }
  return B2_OK;
}

b2_status b2_factor_set_sync(b2_factor_set* s) {
  B2_REQUIRE(s != nullptr, "b2_factor_set_sync: set is NULL");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  B2_CUDA(cudaStreamSynchronize(s->ctx->stream));
  return B2_OK;
}

b2_status b2_factor_set_store_linearized(b2_factor_set* s, b2_linearized* out) {
  B2_REQUIRE(s && out, "b2_factor_set_store_linearized: NULL argument");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  B2_CUDA(cudaMemcpyAsync(out, s->d_out, s->factors.size() * sizeof(b2_linearized), cudaMemcpyDeviceToHost, s->ctx->stream));
  B2_CUDA(cudaStreamSynchronize(s->ctx->stream));
  return B2_OK;
}

b2_status b2_factor_set_store_errors(b2_factor_set* s, double* out_errors) {
  B2_REQUIRE(s && out_errors, "b2_factor_set_store_errors: NULL argument");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  B2_CUDA(cudaMemcpyAsync(out_errors, s->d_err, s->factors.size() * sizeof(double), cudaMemcpyDeviceToHost, s->ctx->stream));
  B2_CUDA(cudaStreamSynchronize(s->ctx->stream));
  return B2_OK;
}

b2_status b2_factor_set_linearize_exchange(b2_factor_set* s, const double* d_deltas, double* d_out, double* const* peer_out, unsigned int* const* peer_flags,
                                           int n_peers, int my_rank, unsigned int seq) {
  B2_REQUIRE(s && d_deltas && d_out && peer_out && peer_flags, "b2_factor_set_linearize_exchange: NULL argument");
  B2_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && my_rank >= 0 && my_rank < n_peers, "b2_factor_set_linearize_exchange: bad peer count / rank");
  B2_REQUIRE(seq != 0u, "b2_factor_set_linearize_exchange: seq must be non-zero");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  DoneSignal sig{};
  sig.counter = s->ctx->d_done_counter;
  sig.flag = peer_flags[my_rank];
  sig.total = static_cast<unsigned int>(s->factors.size());
  sig.seq = seq;
  sig.n_peers = n_peers;
  sig.my_rank = my_rank;
  for (int p = 0; p < n_peers; p++) {
    B2_REQUIRE(peer_flags[p] != nullptr && (p == my_rank || peer_out[p] != nullptr), "b2_factor_set_linearize_exchange: NULL peer pointer");
    sig.peer_out[p] = peer_out[p];
    sig.peer_flag[p] = peer_flags[p];
  }
  B2_TRY(launch_groups(s, MODE_LINEARIZE, d_deltas, d_deltas, d_out, sig));
  for (auto* f : s->factors) f->linearized = true;
  return B2_OK;
}

namespace {
// Completes the exchange: returns (on the stream) once flags[r] == seq for every rank r, i.e. every GPU's records have
// landed in this GPU's buffer.  One warp, lane r polls rank r's word; sleeps between polls, traps instead of hanging.
__global__ void wait_flags_kernel(const volatile unsigned int* __restrict__ flags, int n, unsigned int seq) {
  const int r = threadIdx.x;
  if (r < n) {
    unsigned polls = 0;
    while (flags[r] != seq) {
      __nanosleep(100);
      if (++polls > (1u << 28)) __trap();  // ~30 s: another rank may still be loading its modules at the first step
    }
  }
  __syncwarp();
  __threadfence_system();
}
}  // namespace

namespace {
// A rank that owns no factor still has to take part in the exchange: raise its flag on every GPU.
__global__ void raise_flags_kernel(DoneSignal sig) {
  if (threadIdx.x < sig.n_peers) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int*>(sig.peer_flag[threadIdx.x] + sig.my_rank) = sig.seq;
  }
}
}  // namespace

b2_status b2_exchange_signal(b2_ctx* ctx, unsigned int* const* peer_flags, int n_peers, int my_rank, unsigned int seq) {
  B2_REQUIRE(ctx && peer_flags, "b2_exchange_signal: NULL argument");
  B2_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && my_rank >= 0 && my_rank < n_peers && seq != 0u, "b2_exchange_signal: bad arguments");
  B2_CUDA(cudaSetDevice(ctx->device));
  DoneSignal sig{};
  sig.n_peers = n_peers;
  sig.my_rank = my_rank;
  sig.seq = seq;
  for (int p = 0; p < n_peers; p++) {
    B2_REQUIRE(peer_flags[p] != nullptr, "b2_exchange_signal: NULL peer pointer");
    sig.peer_flag[p] = peer_flags[p];
  }
  raise_flags_kernel<<<1, 32, 0, ctx->stream>>>(sig);
  B2_CUDA(cudaGetLastError());
  return B2_OK;
}

b2_status b2_exchange_wait(b2_ctx* ctx, const unsigned int* d_flags, int n_peers, unsigned int seq) {
  B2_REQUIRE(ctx && d_flags, "b2_exchange_wait: NULL argument");
  B2_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers, "b2_exchange_wait: bad peer count");
  B2_CUDA(cudaSetDevice(ctx->device));
  wait_flags_kernel<<<1, 32, 0, ctx->stream>>>(d_flags, n_peers, seq);
  B2_CUDA(cudaGetLastError());
  return B2_OK;
}

#ifdef B2_WS_TIMING
// development aid (scripts/cta_times.py; never in the production build): per-CTA {start, probe warps done, accumulate warps done,
// flush done} timestamps (ns) of the VGICP kernel's launches since the previous call; resets them
__attribute__((visibility("default"))) void b2_debug_cta_times(unsigned long long* out, int n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, b2::ws::g_cta_times, static_cast<size_t>(n) * 4 * sizeof(unsigned long long));
  static unsigned long long zeros[1024 * 4];
  cudaMemcpyToSymbol(b2::ws::g_cta_times, zeros, sizeof(zeros));
}
#endif

// ---- exchange objects: peer-mapped result blocks through CUDA IPC / peer access, no framework above the ABI ------------------
}  // extern "C"

struct b2_exchange {
  b2_ctx* ctx = nullptr;
  int n = 1, rank = 0;
  size_t num_records = 0;
  double* d_block = nullptr;          // [2][num_records][128] doubles, then 2 x 8 flag words + 8 rendezvous words (padded to 256 bytes)
  unsigned int barrier_seq = 0;       // rendezvous counter (monotonic: the words are never reset)
  double* peer_block[kMaxPeers] = {nullptr};
  bool ipc_opened[kMaxPeers] = {false};
  size_t rec_doubles() const { return num_records * B2_LINEARIZED_DOUBLES; }
  size_t bytes() const { return 2 * rec_doubles() * sizeof(double) + 256; }
};

extern "C" {

b2_status b2_exchange_create(b2_ctx* ctx, int n_ranks, int my_rank, size_t num_records, b2_exchange** out) {
  B2_REQUIRE(out != nullptr, "b2_exchange_create: out is NULL");
  *out = nullptr;
  B2_REQUIRE(ctx != nullptr, "b2_exchange_create: ctx is NULL");
  B2_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxPeers && my_rank >= 0 && my_rank < n_ranks, "b2_exchange_create: bad rank %d of %d (at most %d GPUs of one node)", my_rank, n_ranks, kMaxPeers);
  B2_REQUIRE(num_records > 0, "b2_exchange_create: num_records must be positive");
  B2_CUDA(cudaSetDevice(ctx->device));
  b2_exchange* ex = new b2_exchange;
  ex->ctx = ctx, ex->n = n_ranks, ex->rank = my_rank, ex->num_records = num_records;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->d_block), ex->bytes());
  if (e == cudaSuccess) e = cudaMemset(ex->d_block, 0, ex->bytes());
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (ex->d_block) cudaFree(ex->d_block);
    delete ex;
    return fail(B2_ERR_OUT_OF_MEMORY, "b2_exchange_create: %s", cudaGetErrorString(e));
  }
  ex->peer_block[my_rank] = ex->d_block;
  *out = ex;
  return B2_OK;
}

b2_status b2_exchange_destroy(b2_exchange* ex) {
  if (!ex) return B2_OK;
  cudaSetDevice(ex->ctx->device);
  cudaStreamSynchronize(ex->ctx->stream);
  for (int p = 0; p < ex->n; p++)
    if (ex->ipc_opened[p]) cudaIpcCloseMemHandle(ex->peer_block[p]);
  if (ex->d_block) cudaFree(ex->d_block);
  delete ex;
  return B2_OK;
}

b2_status b2_exchange_export(const b2_exchange* ex, unsigned char handle[B2_IPC_HANDLE_BYTES]) {
  B2_REQUIRE(ex && handle, "b2_exchange_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == B2_IPC_HANDLE_BYTES, "IPC handle size");
  B2_CUDA(cudaSetDevice(ex->ctx->device));
  cudaIpcMemHandle_t h;
  B2_CUDA(cudaIpcGetMemHandle(&h, ex->d_block));
  std::memcpy(handle, &h, sizeof(h));
  return B2_OK;
}

b2_status b2_exchange_import(b2_exchange* ex, int peer_rank, const unsigned char handle[B2_IPC_HANDLE_BYTES]) {
  B2_REQUIRE(ex && handle, "b2_exchange_import: NULL argument");
  B2_REQUIRE(peer_rank >= 0 && peer_rank < ex->n && peer_rank != ex->rank, "b2_exchange_import: bad peer rank %d", peer_rank);
  B2_CUDA(cudaSetDevice(ex->ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  B2_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ex->peer_block[peer_rank] = static_cast<double*>(p);
  ex->ipc_opened[peer_rank] = true;
  return B2_OK;
}

b2_status b2_exchange_enable_peer(b2_exchange* ex, int peer_rank, const b2_exchange* peer) {
  B2_REQUIRE(ex && peer, "b2_exchange_enable_peer: NULL argument");
  B2_REQUIRE(peer_rank >= 0 && peer_rank < ex->n && peer_rank != ex->rank && peer->rank == peer_rank && peer->num_records == ex->num_records, "b2_exchange_enable_peer: mismatched exchange objects");
  B2_CUDA(cudaSetDevice(ex->ctx->device));
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer->ctx->device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(B2_ERR_CUDA, "b2_exchange_enable_peer: %s", cudaGetErrorString(e));
  (void)cudaGetLastError();
  ex->peer_block[peer_rank] = peer->d_block;
  return B2_OK;
}

namespace {
// Stream-ordered rendezvous of the GPUs of an exchange: raise word [my_rank] = seq in every GPU's rendezvous array, then wait
// until every rank's word in THIS GPU's array has reached seq (monotonic counters, so a fast peer's next round cannot be missed).
__global__ void rendezvous_kernel(DoneSignal sig) {
  const int r = threadIdx.x;
  if (r < sig.n_peers) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int*>(sig.peer_flag[r] + sig.my_rank) = sig.seq;
    const volatile unsigned int* mine = sig.peer_flag[sig.my_rank];
    unsigned polls = 0;
    while (static_cast<int>(mine[r] - sig.seq) < 0) {
      __nanosleep(64);
      if (++polls > (1u << 28)) __trap();
    }
  }
  __syncwarp();
  __threadfence_system();
}
}  // namespace

b2_status b2_exchange_barrier(b2_exchange* ex) {
  B2_REQUIRE(ex != nullptr, "b2_exchange_barrier: exchange is NULL");
  for (int p = 0; p < ex->n; p++) B2_REQUIRE(ex->peer_block[p] != nullptr, "b2_exchange_barrier: rank %d has not been imported", p);
  B2_CUDA(cudaSetDevice(ex->ctx->device));
  DoneSignal sig{};
  sig.n_peers = ex->n;
  sig.my_rank = ex->rank;
  sig.seq = ++ex->barrier_seq;
  for (int p = 0; p < ex->n; p++) sig.peer_flag[p] = reinterpret_cast<unsigned int*>(ex->peer_block[p] + 2 * ex->rec_doubles()) + 16;
  rendezvous_kernel<<<1, 32, 0, ex->ctx->stream>>>(sig);
  B2_CUDA(cudaGetLastError());
  return B2_OK;
}

const double* b2_exchange_records(const b2_exchange* ex, unsigned int step) { return ex ? ex->d_block + (step & 1u) * ex->rec_doubles() : nullptr; }

namespace {
__global__ void mirror_kernel(DoneSignal sig) {
  mirror_to_host(sig, threadIdx.x, blockDim.x);
  __syncthreads();
  if (threadIdx.x == 0) *sig.mirror_flag = sig.mirror_seq;
}
b2_status exchange_linearize_impl(b2_exchange* ex, b2_factor_set* s, const double* deltas, size_t first_slot, unsigned int step, double* out_host);
}  // namespace

b2_status b2_exchange_linearize(b2_exchange* ex, b2_factor_set* s, const double* deltas, size_t first_slot, unsigned int step) {
  return exchange_linearize_impl(ex, s, deltas, first_slot, step, nullptr);
}

b2_status b2_exchange_linearize_host(b2_exchange* ex, b2_factor_set* s, const double* deltas, size_t first_slot, unsigned int step, double* out_records) {
  B2_REQUIRE(out_records != nullptr, "b2_exchange_linearize_host: out_records is NULL");
  return exchange_linearize_impl(ex, s, deltas, first_slot, step, out_records);
}

namespace {
b2_status exchange_linearize_impl(b2_exchange* ex, b2_factor_set* s, const double* deltas, size_t first_slot, unsigned int step, double* out_host) {
  B2_REQUIRE(ex != nullptr, "b2_exchange_linearize: exchange is NULL");
  B2_REQUIRE(step != 0u, "b2_exchange_linearize: step starts at 1");
  for (int p = 0; p < ex->n; p++) B2_REQUIRE(ex->peer_block[p] != nullptr, "b2_exchange_linearize: rank %d has not been imported", p);
  const size_t F = s ? s->factors.size() : 0;
  B2_REQUIRE(F == 0 || deltas != nullptr, "b2_exchange_linearize: deltas is NULL");
  B2_REQUIRE(first_slot + F <= ex->num_records, "b2_exchange_linearize: slots [%zu, %zu) exceed the %zu records of the exchange", first_slot, first_slot + F, ex->num_records);
  B2_CUDA(cudaSetDevice(ex->ctx->device));
  const unsigned par = step & 1u;
  DoneSignal sig{};
  sig.n_peers = ex->n;
  sig.my_rank = ex->rank;
  sig.seq = step;
  sig.wait_in_kernel = 1;
  for (int p = 0; p < ex->n; p++) {
    sig.peer_out[p] = ex->peer_block[p] + par * ex->rec_doubles() + first_slot * B2_LINEARIZED_DOUBLES;
    sig.peer_flag[p] = reinterpret_cast<unsigned int*>(ex->peer_block[p] + 2 * ex->rec_doubles()) + par * 8;
  }
  unsigned int host_seq = 0;
  const size_t mirror_bytes = ex->rec_doubles() * sizeof(double);
  if (out_host != nullptr) {
    // every rank's records of this step are delivered to pinned mapped host memory by the kernel that waited for them
    B2_TRY(ex->ctx->ensure_stage(mirror_bytes, 0));
    double* d_mirror = nullptr;
    B2_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_mirror), ex->ctx->h_stage, 0));
    sig.mirror_src = ex->d_block + par * ex->rec_doubles();
    sig.mirror = d_mirror;
    sig.mirror_doubles = static_cast<unsigned int>(ex->rec_doubles());
    sig.mirror_flag = ex->ctx->d_done_flag;
    host_seq = ++ex->ctx->done_seq;
    sig.mirror_seq = host_seq;
  }
  auto deliver = [&]() -> b2_status {
    if (out_host == nullptr) return B2_OK;
    B2_TRY(wait_done(ex->ctx, host_seq));
    std::memcpy(out_host, ex->ctx->h_stage, mirror_bytes);
    return B2_OK;
  };
  if (F == 0) {  // nothing to linearize on this rank: still take part (raise, then wait)
    raise_flags_kernel<<<1, 32, 0, ex->ctx->stream>>>(sig);
    wait_flags_kernel<<<1, 32, 0, ex->ctx->stream>>>(sig.peer_flag[ex->rank], ex->n, step);
    if (out_host != nullptr) mirror_kernel<<<1, 256, 0, ex->ctx->stream>>>(sig);
    B2_CUDA(cudaGetLastError());
    return deliver();
  }
  B2_REQUIRE(s->ctx == ex->ctx, "b2_exchange_linearize: set and exchange live on different contexts");
  sig.counter = s->ctx->d_done_counter;
  sig.flag = sig.peer_flag[ex->rank];
  sig.total = static_cast<unsigned int>(F);
  const uint32_t* d_frozen = nullptr;
  B2_TRY(prepare_frozen_flags(s, deltas, &d_frozen));
  double* d_out = sig.peer_out[ex->rank];
  if (F == 1 && s->groups[0].fn_single[MODE_LINEARIZE] != nullptr) {
    B2_TRY(launch_groups(s, MODE_LINEARIZE, nullptr, nullptr, d_out, sig, deltas));
  } else {
    B2_CUDA(cudaMemcpyAsync(s->d_poses_lin, deltas, F * 16 * sizeof(double), cudaMemcpyHostToDevice, s->ctx->stream));
    B2_TRY(launch_groups(s, MODE_LINEARIZE, s->d_poses_lin, s->d_poses_lin, d_out, sig, nullptr, d_frozen));
  }
  for (size_t i = 0; i < F; i++) {
    s->factors[i]->linearized = true;
    std::memcpy(s->factors[i]->lin_delta, deltas + i * 16, 16 * sizeof(double));
  }
  return deliver();
}
}  // namespace

b2_status b2_factor_set_linearize(b2_factor_set* s, const double* deltas, b2_linearized* out) {
  B2_REQUIRE(s && deltas && out, "b2_factor_set_linearize: NULL argument");
  static const bool trace = std::getenv("B2_TRACE") != nullptr;  // development aid: host-side time split of this call on stderr
  const auto t0 = std::chrono::steady_clock::now();
  B2_CUDA(cudaSetDevice(s->ctx->device));
  cudaStream_t st = s->ctx->stream;
  const size_t F = s->factors.size();
  const size_t in_bytes = F * 16 * sizeof(double), out_bytes = F * sizeof(b2_linearized);
  B2_TRY(s->ctx->ensure_stage(in_bytes + out_bytes, 0));
  char* h = static_cast<char*>(s->ctx->h_stage);
  std::memcpy(h, deltas, in_bytes);
  auto t1 = t0, t2 = t0;
  const uint32_t* d_frozen = nullptr;
  B2_TRY(prepare_frozen_flags(s, deltas, &d_frozen));
  if (F <= kZeroCopyMaxFactors) {
    // Small sets: ONE launch and one stream sync.  The staging buffer is pinned, mapped host memory: the kernel reads the
    // poses straight from it (128 B per factor over PCIe) and its per-factor epilogue writes the 1 KiB result record
    // straight back, so there is no separate H2D / D2H copy operation on the stream (NonlinearFactorSetGPU's protocol
    // needs 2 copies + 2 syncs, nonlinear_factor_set_gpu.cpp:91-124).  The linearization point is kept on the device by
    // the same epilogue (lin_store).
    double* d_in = nullptr;
    B2_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_in), h, 0));
    double* d_res = reinterpret_cast<double*>(reinterpret_cast<char*>(d_in) + in_bytes);
    t1 = std::chrono::steady_clock::now();
    DoneSignal sig{};
    sig.counter = s->ctx->d_done_counter, sig.flag = s->ctx->d_done_flag, sig.total = static_cast<unsigned int>(F), sig.seq = ++s->ctx->done_seq;
    B2_TRY(launch_groups(s, MODE_LINEARIZE, d_in, d_in, d_res, sig, F == 1 ? deltas : nullptr, d_frozen));
    t2 = std::chrono::steady_clock::now();
    B2_TRY(wait_done(s->ctx, sig.seq));
  } else {
    B2_CUDA(cudaMemcpyAsync(s->d_poses_lin, h, in_bytes, cudaMemcpyHostToDevice, st));
    t1 = std::chrono::steady_clock::now();
    B2_TRY(launch_groups(s, MODE_LINEARIZE, s->d_poses_lin, s->d_poses_lin, s->d_out, DoneSignal{}, nullptr, d_frozen));
    B2_CUDA(cudaMemcpyAsync(h + in_bytes, s->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    t2 = std::chrono::steady_clock::now();
    B2_CUDA(cudaStreamSynchronize(st));
  }
  const auto t3 = std::chrono::steady_clock::now();
  std::memcpy(out, h + in_bytes, out_bytes);
  for (size_t i = 0; i < F; i++) {
    s->factors[i]->linearized = true;
    std::memcpy(s->factors[i]->lin_delta, deltas + i * 16, 16 * sizeof(double));
  }
  if (trace) {
    const auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    std::fprintf(stderr, "[b2 trace] linearize F=%zu: prepare %.1f us, launch %.1f us, wait %.1f us, finish %.1f us\n", F, us(t0, t1), us(t1, t2), us(t2, t3),
                 us(t3, std::chrono::steady_clock::now()));
  }
  return B2_OK;
}

b2_status b2_factor_set_error(b2_factor_set* s, const double* deltas_eval, double* out_errors) {
  B2_REQUIRE(s && deltas_eval && out_errors, "b2_factor_set_error: NULL argument");
  B2_CUDA(cudaSetDevice(s->ctx->device));
  cudaStream_t st = s->ctx->stream;
  const size_t F = s->factors.size();
  // A factor evaluated before its first linearization establishes its correspondences at the evaluation point
  // (integrated_vgicp_factor_impl.hpp:183-185): linearize exactly those factors, individually, at delta_eval first.
  for (size_t i = 0; i < F; i++) {
    if (!s->factors[i]->linearized) {
      b2_linearized tmp;
      B2_TRY(b2_factor_linearize(s->factors[i], deltas_eval + i * 16, &tmp));
    }
  }
  const size_t in_bytes = F * 16 * sizeof(double), out_bytes = F * sizeof(double);
  B2_TRY(s->ctx->ensure_stage(in_bytes + out_bytes, 0));
  char* h = static_cast<char*>(s->ctx->h_stage);
  std::memcpy(h, deltas_eval, in_bytes);
  if (F <= kZeroCopyMaxFactors) {
    double* d_in = nullptr;
    B2_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_in), h, 0));
    double* d_res = reinterpret_cast<double*>(reinterpret_cast<char*>(d_in) + in_bytes);
    DoneSignal sig{};
    sig.counter = s->ctx->d_done_counter, sig.flag = s->ctx->d_done_flag, sig.total = static_cast<unsigned int>(F), sig.seq = ++s->ctx->done_seq;
    B2_TRY(launch_groups(s, MODE_ERROR, nullptr, d_in, d_res, sig, F == 1 ? deltas_eval : nullptr));
    B2_TRY(wait_done(s->ctx, sig.seq));
  } else {
    B2_CUDA(cudaMemcpyAsync(s->d_poses_eval, h, in_bytes, cudaMemcpyHostToDevice, st));
    B2_TRY(launch_groups(s, MODE_ERROR, nullptr, s->d_poses_eval, s->d_err));
    B2_CUDA(cudaMemcpyAsync(h + in_bytes, s->d_err, out_bytes, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
  }
  std::memcpy(out_errors, h + in_bytes, out_bytes);
  return B2_OK;
}

#ifdef B2_V2_TIMING
// development aid (not part of the ABI): per-warp cycle totals of the last VGICP launch, [block][32 warps][8 slots]
__attribute__((visibility("default"))) int b2_debug_warp_cycles(unsigned long long* out, int blocks) {
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(out, v2::g_warp_cycles, sizeof(unsigned long long) * blocks * 32 * 8));
}
#endif

// ---- single-factor conveniences ---------------------------------------------------------------------------------

static b2_status ensure_self_set(b2_factor* f) {
  if (f->self_set) return B2_OK;
  b2_factor* one[1] = {f};
  return b2_factor_set_create(f->ctx, one, 1, &f->self_set);
}

b2_status b2_factor_linearize(b2_factor* f, const double* delta, b2_linearized* out) {
  B2_REQUIRE(f && delta && out, "b2_factor_linearize: NULL argument");
  B2_TRY(ensure_self_set(f));
  return b2_factor_set_linearize(f->self_set, delta, out);
}

b2_status b2_factor_issue_linearize(b2_factor* f, const double* delta, double* d_out) {
  B2_REQUIRE(f && delta && d_out, "b2_factor_issue_linearize: NULL argument");
  B2_TRY(ensure_self_set(f));
  return b2_factor_set_issue_linearize(f->self_set, delta, d_out);
}

b2_status b2_factor_issue_error(b2_factor* f, const double* delta_eval, double* d_out_error) {
  B2_REQUIRE(f && delta_eval && d_out_error, "b2_factor_issue_error: NULL argument");
  B2_TRY(ensure_self_set(f));
  if (!f->linearized) {  // integrated_vgicp_factor_impl.hpp:183-185: the first evaluation establishes the correspondences
    b2_linearized tmp;
    B2_TRY(b2_factor_set_linearize(f->self_set, delta_eval, &tmp));
  }
  return b2_factor_set_issue_error(f->self_set, delta_eval, d_out_error);
}

b2_status b2_factor_sync(b2_factor* f) {
  B2_REQUIRE(f != nullptr, "b2_factor_sync: factor is NULL");
  B2_CUDA(cudaSetDevice(f->ctx->device));
  B2_CUDA(cudaStreamSynchronize(f->ctx->stream));
  return B2_OK;
}

b2_status b2_factor_error(b2_factor* f, const double* delta_eval, double* out_error) {
  B2_REQUIRE(f && delta_eval && out_error, "b2_factor_error: NULL argument");
  B2_TRY(ensure_self_set(f));
  return b2_factor_set_error(f->self_set, delta_eval, out_error);
}

}  // extern "C"
