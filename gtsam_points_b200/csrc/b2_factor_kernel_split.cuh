// b2_factor_kernel_split.cuh -- the factor path as TWO kernels with the occupancy each half wants.
// Included by b2_factors.cu after the shared pieces (FactorDesc, accumulate_point_f, warp_reduce32, epilogue).
//
// Why.  Every single-launch form (b2_factor_kernel_ws.cuh, b2_factor_kernel_v2.cuh) is one persistent CTA per SM whose
// 16-20 warps are split between a correspondence-search role and an accumulate role: the 29 float64 accumulators cap the
// warp count, each role is a latency-bound dependent chain, and the hand-offs between the roles (rings, stages, drains)
// couple those chains (profiles/r02_experiments.md: nothing saturates, issue slots 35 % busy).  Here
//   * `probe_kernel` (correspondence search) is an ordinary streaming kernel: one thread per source point, ~40 registers,
//     full occupancy (60 warps / SM) -- coordinates -> rotate -> floor -> hash -> bucket group -> corr[];
//   * `factor_kernel` (residual + Jacobian + reduction) is persistent, one CTA of kWarps accumulate warps per SM and NOTHING
//     else: every warp streams its own batches of 32 points through a private kStages-deep cp.async pipeline
//     (coordinates, covariance planes, the voxel record named by corr[] -- lane-private shared-memory slots, no barriers, no
//     rings), so all warps of the SM do float64 work and the operand latency hides behind kStages - 1 batches of arithmetic.
// The correspondences travel through corr[] (4 bytes per point, L2-resident between the two launches: the reference's
// own update_correspondences -> linearize split, integrated_vgicp_factor_gpu.cu), the coordinates are re-read from L2.
// error() launches only the second kernel (frozen correspondences), exactly like the reference.
//
// Determinism: batch -> warp assignment is static, lanes accumulate their own points in batch order, cross-warp / cross-CTA
// sums run in slot order => bit-reproducible.  The correspondence indices are computed with the same individually rounded
// float64 operations as before => bit-identical to the CPU oracle.

namespace b2 {
namespace sp {

#ifndef B2_SP_WARPS
#define B2_SP_WARPS 12
#endif
#ifndef B2_SP_STAGES
#define B2_SP_STAGES 3
#endif
constexpr int kWarps = B2_SP_WARPS;
constexpr int kThreads = kWarps * 32;
constexpr int kTile = kThreads;  // source points per tile: one batch of 32 per accumulate warp, one point per probe thread
constexpr int kStages = B2_SP_STAGES;
static_assert(kStages >= 2 && kStages <= 6, "pipeline depth");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
template <int BYTES>
__device__ __forceinline__ void cp_async_small(uint32_t dst, const void* src) {  // 4 or 8 bytes
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One pipeline stage of one warp: lane-private slots, plane-major so that the read-back is conflict-free.
//   rec  5 planes x 32 lanes x 16 B   (mean | cov upper | count of the target record)
//   cov  6 planes x 32 lanes x sizeof(CT)
//   xyz  3 planes x 32 lanes x sizeof(PT)
//   id   32 x 4 B
template <typename PT, typename CT>
struct StageLayout {
  static constexpr uint32_t kRec = 0;
  static constexpr uint32_t kCov = 5 * 32 * 16;
  static constexpr uint32_t kXyz = kCov + 6 * 32 * sizeof(CT);
  static constexpr uint32_t kId = kXyz + 3 * 32 * sizeof(PT);
  static constexpr uint32_t kBytes = (kId + 32 * 4 + 15u) & ~15u;
  static constexpr size_t kTotal = static_cast<size_t>(kWarps) * kStages * kBytes;
};

struct Shared {
  FactorDesc desc;
  double red[kWarps][kAcc];
  double tot[kAcc];
  double A[36], X[36], D[36];
  double R[9], t[3];  // pose the residuals of the current factor run are evaluated at
  double RL[9];       // rotation of its linearization point (== R when linearizing)
  int flag;
};

__host__ __device__ __forceinline__ uint32_t cta_tile_begin(uint32_t c, uint32_t T, uint32_t G) {
  return static_cast<uint32_t>(static_cast<unsigned long long>(c) * T / G);
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 1: correspondence search, VGICP (voxel hash probe).  One CTA per tile of the set's tile list, one thread per point.
// Reference: integrated_vgicp_factor_gpu.cu / vgicp_derivatives (lookup_voxels kernel): corr[i] = voxel of (T p_i) or -1.
// ---------------------------------------------------------------------------------------------------------------
template <typename PT, bool SINGLE>
__global__ void __launch_bounds__(kTile) probe_kernel(const FactorDesc* __restrict__ descs, const uint32_t* __restrict__ tile_factor, const double* __restrict__ poses_lin,
                                                      const __grid_constant__ PoseArg pose, const uint32_t* __restrict__ frozen_flags) {
  const uint32_t tile = blockIdx.x;
  const FactorDesc* __restrict__ dg = descs + (SINGLE ? 0u : __ldg(tile_factor + tile));
  const uint32_t out_index = dg->out_index;
  // correspondence-update tolerance (integrated_gicp_factor_impl.hpp:135-147): this factor keeps its stored correspondences
  if (frozen_flags != nullptr && __ldg(frozen_flags + out_index) != 0u) return;
  __shared__ double P[12];
  if (!SINGLE) {
    if (threadIdx.x < 12) {
      const double* pe = poses_lin + static_cast<size_t>(out_index) * 16;
      P[threadIdx.x] = __ldg(pe + (threadIdx.x < 9 ? (threadIdx.x / 3) * 4 + threadIdx.x % 3 : (threadIdx.x - 9) * 4 + 3));
    }
    __syncthreads();
  }
  auto Rm = [&](int i) -> double { return SINGLE ? pose.m[(i / 3) * 4 + (i % 3)] : P[i]; };
  auto tvec = [&](int i) -> double { return SINGLE ? pose.m[i * 4 + 3] : P[9 + i]; };
  const uint32_t i = (tile - dg->tile_begin) * kTile + threadIdx.x;
  if (i >= dg->n) return;
  const PT* __restrict__ px = static_cast<const PT*>(dg->pts);
  const size_t n_pad = dg->n_pad;
  const double x = static_cast<double>(__ldg(px + i)), y = static_cast<double>(__ldg(px + n_pad + i)), z = static_cast<double>(__ldg(px + 2 * n_pad + i));
  const VoxelBucket* __restrict__ buckets = dg->buckets;
  const uint32_t bucket_mask = dg->bucket_mask;
  const double inv_leaf = dg->inv_leaf;
  // q = R p + t : coefficient sums in index order, each operation individually rounded (bit-parity with the CPU float64 path)
  const double q0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Rm(0), x), __dmul_rn(Rm(1), y)), __dmul_rn(Rm(2), z)), tvec(0));
  const double q1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Rm(3), x), __dmul_rn(Rm(4), y)), __dmul_rn(Rm(5), z)), tvec(1));
  const double q2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Rm(6), x), __dmul_rn(Rm(7), y)), __dmul_rn(Rm(8), z)), tvec(2));
  const int cx = voxel_coord1(q0, inv_leaf), cy = voxel_coord1(q1, inv_leaf), cz = voxel_coord1(q2, inv_leaf);
  uint32_t g = voxel_hash(cx, cy, cz) & bucket_mask;
  int id = match_group(load_group(buckets, g), cx, cy, cz);
  while (id == -2) {  // rare: the home group is full, walk on
    g = (g + 1) & bucket_mask;
    id = match_group(load_group(buckets, g), cx, cy, cz);
  }
  dg->corr[i] = id;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-factor flush: warp butterfly -> cross-warp sum -> fixed slot; the last CTA of the factor sums the slots in slot
// order and runs the epilogue (H_t = X^T A' X, ...).
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void flush_factor(Shared& sh, double (&v)[kAcc], int tid, double* __restrict__ partials, unsigned int* __restrict__ counters, double* __restrict__ out,
                                             const double* __restrict__ pose_lin /* this factor's linearization pose (16 doubles) */, const DoneSignal& sig) {
  const int lane = tid & 31, warp = tid >> 5;
  const double w = warp_reduce32(v, lane);
  sh.red[warp][lane] = w;
  __syncthreads();
  const FactorDesc& d = sh.desc;
  const uint32_t slot = blockIdx.x - d.cta_first[MODE];
  if (warp == 0) {
    double s = sh.red[0][lane];
#pragma unroll
    for (int k = 1; k < kWarps; k++) s += sh.red[k][lane];
    partials[(static_cast<size_t>(d.slot_begin[MODE]) + slot) * kAcc + lane] = s;
    __threadfence();  // only the writing warp pays for the fence
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(&counters[d.out_index], 1u);
    sh.flag = (prev == d.num_slots[MODE] - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!sh.flag) return;

  // ---- last CTA of this factor ----
  __threadfence();
  {
    double s = 0.0;
    for (uint32_t sl = warp; sl < d.num_slots[MODE]; sl += kWarps) s += __ldcg(&partials[(static_cast<size_t>(d.slot_begin[MODE]) + sl) * kAcc + lane]);
    sh.red[warp][lane] = s;
  }
  __syncthreads();
  if (tid < kAcc) {
    double s = sh.red[0][tid];
#pragma unroll
    for (int k = 1; k < kWarps; k++) s += sh.red[k][tid];
    sh.tot[tid] = s;
  }
  if (tid == 0) counters[d.out_index] = 0u;  // re-arm for the next launch
  __syncthreads();

  if (MODE == MODE_ERROR) {
    if (tid == 0) {
      out[d.out_index] = sh.tot[27];
      __threadfence_system();  // `out` may be mapped host memory
      signal_done(sig);
    }
    __syncthreads();
    return;
  }
  epilogue_build(sh.A, sh.X, sh.D, sh.tot, sh.R, sh.t, tid);
  __syncthreads();
  double* rec = out + static_cast<size_t>(d.out_index) * B2_LINEARIZED_DOUBLES;
  epilogue_store(rec, sh.A, sh.X, sh.D, sh.tot, tid);
  if (tid >= 100 && tid < 116) {
    // remember the linearization point with the factor (error-only launches of ANY set read it back)
    d.lin_pose[tid - 100] = pose_lin[tid - 100];
  }
  __threadfence_system();  // `out` may be mapped host memory (zero-copy host API)
  __syncthreads();
  if (sig.n_peers > 1) {
    // multi-GPU exchange fused into the epilogue: copy the finished record into the same slot of every peer's buffer
    // (plain stores to peer memory over NVLink), fence at system scope, then signal
    for (int p = 0; p < sig.n_peers; p++) {
      if (p == sig.my_rank) continue;
      double* dst = sig.peer_out[p] + static_cast<size_t>(d.out_index) * B2_LINEARIZED_DOUBLES;
      for (int i = tid; i < B2_LINEARIZED_DOUBLES; i += kThreads) dst[i] = __ldcg(rec + i);
    }
    __threadfence_system();
    __syncthreads();
  }
  if (tid == 0) signal_done(sig);  // every writer of this record fenced before the barrier
}

template <int KIND>
struct MetricOf {
  static constexpr int value = KIND <= 1 ? 0 : KIND - 1;
};

// ---------------------------------------------------------------------------------------------------------------
// Kernel 2: residuals, Jacobians and the reduction over the correspondences stored in corr[].
// KIND: 0 VGICP, 1 GICP, 2 point-to-point ICP, 3 point-to-plane ICP (only the metric and the covariance stream differ).
// SINGLE: the launch covers exactly one factor and its pose is the by-value parameter `pose`.
// ---------------------------------------------------------------------------------------------------------------
template <typename PT, typename CT, int KIND, int MODE, bool SINGLE = false>
__global__ void __launch_bounds__(kThreads, 1)
factor_kernel(const FactorDesc* __restrict__ descs, const uint32_t* __restrict__ tile_factor, uint32_t num_tiles, const double* __restrict__ poses_lin,
              const double* __restrict__ poses_eval, double* __restrict__ partials, unsigned int* __restrict__ counters, double* __restrict__ out,
              const __grid_constant__ DoneSignal sig, const __grid_constant__ PoseArg pose, const uint32_t* __restrict__ frozen_flags) {
  using L = StageLayout<PT, CT>;
  __shared__ Shared sh;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  (void)frozen_flags;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t stage0 = smem_u32(dyn_smem) + static_cast<uint32_t>(warp) * kStages * L::kBytes;  // this warp's stages
  const unsigned char* const stage0_g = dyn_smem + static_cast<size_t>(warp) * kStages * L::kBytes;
  const uint32_t G = gridDim.x;
  const uint32_t tile_lo = cta_tile_begin(blockIdx.x, num_tiles, G);
  const uint32_t tile_hi = cta_tile_begin(blockIdx.x + 1, num_tiles, G);

  uint32_t tile = tile_lo;
  double acc[kAcc];
  while (tile < tile_hi) {
    const uint32_t f = SINGLE ? 0u : __ldg(tile_factor + tile);
    __syncthreads();  // previous flush is done with sh.desc
    if (tid < static_cast<int>(sizeof(FactorDesc) / 4)) reinterpret_cast<uint32_t*>(&sh.desc)[tid] = __ldg(reinterpret_cast<const uint32_t*>(descs + f) + tid);
    __syncthreads();
    const FactorDesc& d = sh.desc;
    if (tid < 21) {
      const double* pe = SINGLE ? pose.m : ((MODE == MODE_ERROR ? poses_eval : poses_lin) + static_cast<size_t>(d.out_index) * 16);
      const double* pl = (MODE == MODE_ERROR) ? d.lin_pose : pe;
      if (tid < 9)
        sh.R[tid] = pe[(tid / 3) * 4 + tid % 3];
      else if (tid < 12)
        sh.t[tid - 9] = pe[(tid - 9) * 4 + 3];
      else
        sh.RL[tid - 12] = pl[((tid - 12) / 3) * 4 + (tid - 12) % 3];
    }
    __syncthreads();
    // Pose operands.  SINGLE: uniform-register / constant-bank operands straight from the parameter.  Otherwise registers;
    // when linearizing, the evaluation rotation IS the linearization rotation (one copy).
    constexpr bool kConstR = SINGLE;                              // evaluation rotation + translation from `pose`
    constexpr bool kConstRL = SINGLE && MODE == MODE_LINEARIZE;   // linearization rotation from `pose`
    constexpr bool kSharedRL = !SINGLE && MODE == MODE_LINEARIZE; // RL aliases the R registers
    double Rr[kConstR ? 1 : 9], tr[kConstR ? 1 : 3], RLr[(kConstRL || kSharedRL) ? 1 : 9];
    if (!kConstR) {
#pragma unroll
      for (int k = 0; k < 9; k++) Rr[k] = sh.R[k];
#pragma unroll
      for (int k = 0; k < 3; k++) tr[k] = sh.t[k];
    }
    if (!kConstRL && !kSharedRL) {
#pragma unroll
      for (int k = 0; k < 9; k++) RLr[k] = sh.RL[k];
    }
    auto rm = [&](int i) -> double { return kConstR ? pose.m[(i / 3) * 4 + (i % 3)] : Rr[kConstR ? 0 : i]; };
    auto tt = [&](int i) -> double { return kConstR ? pose.m[i * 4 + 3] : tr[kConstR ? 0 : i]; };
    auto rl = [&](int i) -> double { return kConstRL ? pose.m[(i / 3) * 4 + (i % 3)] : (kSharedRL ? Rr[kConstR ? 0 : i] : RLr[(kConstRL || kSharedRL) ? 0 : i]); };
#pragma unroll
    for (int k = 0; k < kAcc; k++) acc[k] = 0.0;

    const double* __restrict__ records = d.records;
    const PT* __restrict__ px = static_cast<const PT*>(d.pts);
    const CT* __restrict__ cv = static_cast<const CT*>(d.covs);
    const int32_t* __restrict__ corr = d.corr;
    const size_t n_pad = d.n_pad;
    const uint32_t n = d.n;
    const uint32_t f_tile_begin = d.tile_begin, f_num_tiles = d.num_tiles, perm_stride = d.perm_stride;
    const uint32_t run_end = min(tile_hi, f_tile_begin + f_num_tiles);
    const uint32_t nb = run_end - tile;  // batches of this warp in this run (one per tile)

    // virtual tile v of the factor is physical tile (v * S) mod n_tiles, S ~ 0.618 n_tiles coprime to n_tiles: every CTA samples
    // the (Morton-ordered) cloud quasi-uniformly, dense and empty regions spread evenly over the SMs
    auto next_tile = [&](uint32_t pt) {
      pt += perm_stride;
      return pt >= f_num_tiles ? pt - f_num_tiles : pt;
    };
    const uint32_t pt_first = static_cast<uint32_t>(static_cast<unsigned long long>(tile - f_tile_begin) * perm_stride % f_num_tiles);
    const uint32_t lane_off = static_cast<uint32_t>(warp) * 32u + static_cast<uint32_t>(lane);
    auto load_id = [&](uint32_t pt) -> int {
      const uint32_t i = pt * kTile + lane_off;
      return i < n ? __ldg(corr + i) : -1;
    };
    // request the operands of one batch into stage `s` (lane-private slots; misses request nothing)
    auto issue = [&](uint32_t s, uint32_t pt, int id) {
      const uint32_t base = stage0 + s * L::kBytes;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + L::kId + lane * 4u), "r"(id) : "memory");
      if (id >= 0) {
        const uint32_t i = pt * kTile + lane_off;
        const double* rec = records + static_cast<size_t>(id) * kRecordDoubles;
#pragma unroll
        for (int k = 0; k < 5; k++) cp_async16(base + L::kRec + (k * 32u + lane) * 16u, rec + 2 * k);
        if (KIND <= 1) {
#pragma unroll
          for (int k = 0; k < 6; k++) cp_async_small<static_cast<int>(sizeof(CT))>(base + L::kCov + (k * 32u + lane) * static_cast<uint32_t>(sizeof(CT)), cv + static_cast<size_t>(k) * n_pad + i);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) cp_async_small<static_cast<int>(sizeof(PT))>(base + L::kXyz + (k * 32u + lane) * static_cast<uint32_t>(sizeof(PT)), px + static_cast<size_t>(k) * n_pad + i);
      }
    };

    // ---- prologue: ids of the first kStages batches, requests of the first kStages - 1 ----
    uint32_t pt_issue = pt_first;  // physical tile of the next batch to request
    {
      int idq[kStages - 1];
      uint32_t pt = pt_first;
#pragma unroll
      for (int j = 0; j < kStages - 1; j++) {
        idq[j] = static_cast<uint32_t>(j) < nb ? load_id(pt) : -1;
        pt = next_tile(pt);
      }
#pragma unroll
      for (int j = 0; j < kStages - 1; j++) {
        if (static_cast<uint32_t>(j) < nb) issue(static_cast<uint32_t>(j), pt_issue, idq[j]);
        cp_async_commit();
        pt_issue = next_tile(pt_issue);
      }
    }
    uint32_t pt_id = pt_issue;  // physical tile of the next batch whose id is to be loaded
    int id_ahead = (kStages - 1 < nb) ? load_id(pt_id) : -1;
    pt_id = next_tile(pt_id);
    uint32_t s_cur = 0u, s_issue = kStages - 1;

#pragma unroll 1
    for (uint32_t k = 0; k < nb; k++) {
      if (k + (kStages - 1) < nb) issue(s_issue, pt_issue, id_ahead);
      cp_async_commit();
      pt_issue = next_tile(pt_issue);
      if (k + kStages < nb) id_ahead = load_id(pt_id);
      pt_id = next_tile(pt_id);
      cp_async_wait<kStages - 1>();  // this lane's requests of batch k have landed (lane-private slots: no barrier needed)

      const unsigned char* sb = stage0_g + s_cur * L::kBytes;
      const int id = *reinterpret_cast<const volatile int*>(sb + L::kId + lane * 4);
      if (id >= 0) {
        TargetRec T;
        const double2* rp = reinterpret_cast<const double2*>(sb + L::kRec) + lane;
        T.r01 = rp[0];
        T.r23 = rp[32];
        T.r45 = rp[64];
        T.r67 = rp[96];
        T.r89 = rp[128];
        SourceCov A;
        if (KIND <= 1) {
          const CT* cp = reinterpret_cast<const CT*>(sb + L::kCov) + lane;
          A.a00 = static_cast<double>(cp[0]);
          A.a01 = static_cast<double>(cp[32]);
          A.a02 = static_cast<double>(cp[64]);
          A.a11 = static_cast<double>(cp[96]);
          A.a12 = static_cast<double>(cp[128]);
          A.a22 = static_cast<double>(cp[160]);
        } else {
          A = SourceCov{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        }
        const PT* xp = reinterpret_cast<const PT*>(sb + L::kXyz) + lane;
        const double x = static_cast<double>(xp[0]), y = static_cast<double>(xp[32]), z = static_cast<double>(xp[64]);
        // u = R p, the same individually rounded operations as the correspondence search
        const double u0 = __dadd_rn(__dadd_rn(__dmul_rn(rm(0), x), __dmul_rn(rm(1), y)), __dmul_rn(rm(2), z));
        const double u1 = __dadd_rn(__dadd_rn(__dmul_rn(rm(3), x), __dmul_rn(rm(4), y)), __dmul_rn(rm(5), z));
        const double u2 = __dadd_rn(__dadd_rn(__dmul_rn(rm(6), x), __dmul_rn(rm(7), y)), __dmul_rn(rm(8), z));
        accumulate_point_f<MODE, MetricOf<KIND>::value>(acc, rl, tt, u0, u1, u2, T, A);
      }
      s_cur = (s_cur + 1 == kStages) ? 0u : s_cur + 1;
      s_issue = (s_issue + 1 == kStages) ? 0u : s_issue + 1;
    }
    cp_async_wait<0>();
    flush_factor<MODE>(sh, acc, tid, partials, counters, out, SINGLE ? pose.m : (poses_lin + static_cast<size_t>(sh.desc.out_index) * 16), sig);
    tile = run_end;
  }
}

}  // namespace sp
}  // namespace b2
