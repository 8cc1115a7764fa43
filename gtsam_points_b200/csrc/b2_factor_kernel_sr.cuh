// b2_factor_kernel_sr.cuh -- the VGICP factor kernel, single-role form: every warp runs the WHOLE per-point chain as a software
// pipeline through lane-private shared-memory slots.  Included by b2_factors.cu after the shared pieces (FactorDesc,
// accumulate_point_f, warp_reduce32, epilogue).
//
// Why (profiles/r02_experiments.md).  The warp-specialised forms hand every hit from a search warp to an accumulate warp
// (rings / mbarrier stages): with 16-20 latency-bound warps per SM the hand-offs couple the chains and the SM idles (issue
// slots 35 % busy, no unit saturated).  Measured on the two-launch form (b2_factor_kernel_split.cuh): once the operands arrive
// asynchronously a warp issues one instruction per ~6.6 cycles -- the dependent-issue latency of its float64 chain -- so the
// time is (instructions per point) x 6.6 cycles / (independent chains per scheduler).  The 29 accumulators (58 registers) cap
// the WARPS per SM, but not the chains: a lane that carries kPPL points through ONE basic block shares its accumulators
// between kPPL independent chains.  Here each of the kWarps warps of the persistent CTA owns a static list of batches of
// 32 x kPPL points and, in iteration `it`, does for every one of its kPPL points per lane
//     T1  batch it+2 : coordinates (landed) -> R p + t -> floor -> hash -> REQUEST the home bucket group      (cp.async)
//     T2  batch it+1 : bucket group (landed) -> match -> corr[] -> REQUEST voxel record + source covariance   (cp.async)
//     T0  batch it+3 : REQUEST the coordinates                                                                (cp.async)
//     T3  batch it   : record, covariance, coordinates (landed) -> u = R p -> 154 float64 operations -> 29 accumulators
// Every request is a lane-private cp.async: a lane only ever reads back what it requested itself, so there are no barriers,
// no rings, no mbarriers and no warp ever waits for another warp; ONE `cp.async.wait_group 0` at the top of an iteration is
// the only synchronisation (everything requested during the previous iteration -- whose T3 gave it kPPL x ~1000 cycles -- has
// landed).  T3 is branch-free: a lane without a correspondence computes on stand-in operands (record 0, its own covariance)
// with its weight matrix scaled by an exact 0.0, so the kPPL chains interleave instead of diverging.
// error() (frozen correspondences): T1 disappears, T0 also requests corr[] and T2 only forwards it.
//
// Determinism: batch -> warp assignment is static, lanes accumulate their own points in batch order, cross-warp / cross-CTA
// sums run in slot order => bit-reproducible.  The correspondence search uses the same individually rounded float64
// operations as every other form => indices bit-identical to the CPU oracle.

namespace b2 {
namespace sr {

#ifndef B2_SR_WARPS
#define B2_SR_WARPS 8
#endif
#ifndef B2_SR_PPL
#define B2_SR_PPL 2
#endif
constexpr int kWarps = B2_SR_WARPS;
constexpr int kPPL = B2_SR_PPL;                // points per lane and batch (independent chains sharing one accumulator set)
constexpr int kThreads = kWarps * 32;
constexpr int kWarpPoints = 32 * kPPL;
constexpr int kTile = kWarps * kWarpPoints;    // source points per tile: one batch per warp
constexpr int kXS = 4;                         // coordinate slots: requested at it - 3, hashed at it - 2, used again by T3 at it
constexpr int kGS = 2;                         // bucket-group slots: requested at it - 2, matched at it - 1
constexpr int kRS = 2;                         // operand slots: requested at it - 1, consumed at it

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
__device__ __forceinline__ void st_shared_b32(uint32_t addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Per-warp shared memory (all slots lane-private, plane-major so that every access is conflict-free); every plane holds
// kPPL x 32 entries (point p of lane l at entry p * 32 + l):
//   X  kXS slots: 3 coordinate planes x sizeof(PT)
//   G  linearize: kGS slots of [kGroup bucket planes x 16 B | 3 voxel-coordinate planes x 4 B]
//      error:     3 slots of one 4 B plane (the frozen correspondences on their way from corr[] to T2)
//   R  kRS slots: 5 record planes x 16 B | 6 covariance planes x sizeof(CT) | one id plane x 4 B
template <typename PT, typename CT, int MODE>
struct Layout {
  static constexpr uint32_t kN = kWarpPoints;  // entries per plane
  static constexpr uint32_t kXSlot = 3 * kN * sizeof(PT);
  static constexpr uint32_t kGCoord = kGroup * kN * 16;
  static constexpr uint32_t kGSlot = MODE == MODE_LINEARIZE ? kGCoord + 3 * kN * 4 : kN * 4;
  static constexpr uint32_t kGSlots = MODE == MODE_LINEARIZE ? kGS : 3;  // error: requested at it - 3, forwarded at it - 1
  static constexpr uint32_t kRCov = 5 * kN * 16;
  static constexpr uint32_t kRId = kRCov + 6 * kN * sizeof(CT);
  static constexpr uint32_t kRSlot = kRId + kN * 4;
  static constexpr uint32_t kX = 0;
  static constexpr uint32_t kG = kX + kXS * kXSlot;
  static constexpr uint32_t kR = kG + kGSlots * kGSlot;
  static constexpr uint32_t kWarpBytes = kR + kRS * kRSlot;
  static_assert(kXSlot % 16 == 0 && kGSlot % 16 == 0 && kRSlot % 16 == 0 && kG % 16 == 0 && kR % 16 == 0, "16-byte aligned slots");
  static constexpr size_t kTotal = static_cast<size_t>(kWarps) * kWarpBytes;
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

// ---------------------------------------------------------------------------------------------------------------
// The kernel (KIND 0 = VGICP only: voxel hash probe).  MODE: linearize (search + accumulate) / error-only (frozen corr[]).
// SINGLE: the launch covers exactly one factor and its pose is the by-value parameter `pose` (uniform-register operands).
// ---------------------------------------------------------------------------------------------------------------
template <typename PT, typename CT, int KIND, int MODE, bool SINGLE = false>
__global__ void __launch_bounds__(kThreads, 1)
factor_kernel(const FactorDesc* __restrict__ descs, const uint32_t* __restrict__ tile_factor, uint32_t num_tiles, const double* __restrict__ poses_lin,
              const double* __restrict__ poses_eval, double* __restrict__ partials, unsigned int* __restrict__ counters, double* __restrict__ out,
              const __grid_constant__ DoneSignal sig, const __grid_constant__ PoseArg pose, const uint32_t* __restrict__ frozen_flags) {
  static_assert(KIND == 0, "the single-role kernel serves the voxel-map path");
  using L = Layout<PT, CT, MODE>;
  __shared__ Shared sh;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  (void)frozen_flags;  // VGICP has no correspondence-update tolerance (integrated_vgicp_factor_impl.hpp: always re-associates)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned char* const wb_g = dyn_smem + static_cast<size_t>(warp) * L::kWarpBytes;  // this warp's slots (generic address)
  const uint32_t wb = smem_u32(wb_g);                                                      // ... shared-window address
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
    constexpr bool kConstR = SINGLE;                               // evaluation rotation + translation from `pose`
    constexpr bool kConstRL = SINGLE && MODE == MODE_LINEARIZE;    // linearization rotation from `pose`
    constexpr bool kSharedRL = !SINGLE && MODE == MODE_LINEARIZE;  // RL aliases the R registers
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
    int32_t* __restrict__ corr = d.corr;
    const VoxelBucket* __restrict__ buckets = d.buckets;
    const uint32_t bucket_mask = d.bucket_mask;
    const double inv_leaf = d.inv_leaf;
    const size_t n_pad = d.n_pad;
    const uint32_t n = d.n;
    const uint32_t f_tile_begin = d.tile_begin, f_num_tiles = d.num_tiles, perm_stride = d.perm_stride;
    const uint32_t run_end = min(tile_hi, f_tile_begin + f_num_tiles);
    const int nb = static_cast<int>(run_end - tile);  // batches of this warp in this run (one per tile)

    // virtual tile v of the factor is physical tile (v * S) mod n_tiles, S ~ 0.618 n_tiles coprime to n_tiles: every CTA samples
    // the (Morton-ordered) cloud quasi-uniformly, dense and empty regions spread evenly over the SMs
    auto next_tile = [&](uint32_t pt) {
      pt += perm_stride;
      return pt >= f_num_tiles ? pt - f_num_tiles : pt;
    };
    const uint32_t pt_first = static_cast<uint32_t>(static_cast<unsigned long long>(tile - f_tile_begin) * perm_stride % f_num_tiles);
    const uint32_t lane_off = static_cast<uint32_t>(warp) * kWarpPoints + static_cast<uint32_t>(lane);  // + p * 32: this lane's p-th point of a tile
    uint32_t pt_t0 = pt_first;  // physical tile of the batch T0 handles next
    uint32_t pt_t2 = pt_first;  // ... T2
    const bool has_records = d.num_records != 0u;  // an empty map has no stand-in record for the branch-free T3: skip it altogether
    // stand-in record of a lane without a correspondence: record (i mod 2^k), 2^k <= num_records -- spread over the records like
    // the points themselves (ONE common stand-in would be a single hot L2 line hammered by every SM: measured 8x slower)
    const uint32_t standin_mask = has_records ? (0x80000000u >> __clz(d.num_records)) - 1u : 0u;
    int hits = 0;                                  // this lane's inlier count (kept as an integer; acc[28] at the flush)

    // iteration `it` runs T1(it + 2), T2(it + 1), T0(it + 3), T3(it); batches outside [0, nb) are skipped, so the first three
    // iterations fill the pipeline and the last ones drain it
#pragma unroll 1
    for (int it = -3; it < nb; it++) {
      const uint32_t ub = static_cast<uint32_t>(it + 4);  // non-negative, same residues mod 4 / 2 as `it`
      cp_async_wait<0>();  // everything requested in the previous iteration has landed

      // ---- T1: hash batch it + 2, request its home bucket groups ----
      if (MODE == MODE_LINEARIZE) {
        const int j = it + 2;
        if (j >= 0 && j < nb) {
          const PT* xp = reinterpret_cast<const PT*>(wb_g + L::kX + ((ub + 2) % kXS) * L::kXSlot) + lane;
          const uint32_t gs = wb + L::kG + (ub % kGS) * L::kGSlot;
#pragma unroll
          for (int p = 0; p < kPPL; p++) {
            const double x = static_cast<double>(xp[p * 32]), y = static_cast<double>(xp[L::kN + p * 32]), z = static_cast<double>(xp[2 * L::kN + p * 32]);
            // q = R p + t : coefficient sums in index order, each operation individually rounded (bit-parity with the CPU float64 path)
            const double q0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(rm(0), x), __dmul_rn(rm(1), y)), __dmul_rn(rm(2), z)), tt(0));
            const double q1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(rm(3), x), __dmul_rn(rm(4), y)), __dmul_rn(rm(5), z)), tt(1));
            const double q2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(rm(6), x), __dmul_rn(rm(7), y)), __dmul_rn(rm(8), z)), tt(2));
            const int cx = voxel_coord1(q0, inv_leaf), cy = voxel_coord1(q1, inv_leaf), cz = voxel_coord1(q2, inv_leaf);
            const uint32_t g = voxel_hash(cx, cy, cz) & bucket_mask;
            const int4* gp = reinterpret_cast<const int4*>(buckets) + static_cast<size_t>(g) * kGroup;
            const uint32_t e = p * 32u + lane;
#pragma unroll
            for (int k = 0; k < kGroup; k++) cp_async16(gs + (k * L::kN + e) * 16u, gp + k);
            st_shared_b32(gs + L::kGCoord + e * 4u, cx);
            st_shared_b32(gs + L::kGCoord + (L::kN + e) * 4u, cy);
            st_shared_b32(gs + L::kGCoord + (2u * L::kN + e) * 4u, cz);
          }
        }
      }

      // ---- T2: match batch it + 1 (or forward its frozen correspondence), request record + covariance ----
      {
        const int j = it + 1;
        if (j >= 0 && j < nb) {
          const uint32_t rs = wb + L::kR + ((ub + 1) % kRS) * L::kRSlot;
          const unsigned char* gsl = wb_g + L::kG + ((ub + 1) % L::kGSlots) * L::kGSlot;
#pragma unroll
          for (int p = 0; p < kPPL; p++) {
            const uint32_t e = p * 32u + lane;
            const uint32_t i = pt_t2 * kTile + lane_off + p * 32u;
            int id;
            if (MODE == MODE_LINEARIZE) {
              BucketGroup grp;
#pragma unroll
              for (int k = 0; k < kGroup; k++) grp.b[k] = reinterpret_cast<const int4*>(gsl)[k * L::kN + e];
              const int* cp = reinterpret_cast<const int*>(gsl + L::kGCoord) + e;
              const int cx = cp[0], cy = cp[L::kN], cz = cp[2 * L::kN];
              id = match_group(grp, cx, cy, cz);
              if (id == -2) {  // the home group is full and does not hold the key (~5 % of the points): walk on with ordinary loads
                uint32_t g = voxel_hash(cx, cy, cz) & bucket_mask;
                do {
                  g = (g + 1) & bucket_mask;
                  id = match_group(load_group(buckets, g), cx, cy, cz);
                } while (id == -2);
              }
              if (i >= n) id = -1;
              if (i < n) corr[i] = id;
            } else {
              id = *reinterpret_cast<const volatile int*>(gsl + e * 4);
              if (i >= n) id = -1;
            }
            st_shared_b32(rs + L::kRId + e * 4u, id);
            if (has_records) {
              // lanes without a correspondence fetch stand-ins (some record, the covariance of point i or 0): T3 is branch-free
              const double* rec = records + static_cast<size_t>(id < 0 ? (i & standin_mask) : static_cast<uint32_t>(id)) * kRecordDoubles;
              const uint32_t ic = i < n ? i : 0u;
#pragma unroll
              for (int k = 0; k < 5; k++) cp_async16(rs + (k * L::kN + e) * 16u, rec + 2 * k);
#pragma unroll
              for (int k = 0; k < 6; k++) cp_async_small<static_cast<int>(sizeof(CT))>(rs + L::kRCov + (k * L::kN + e) * static_cast<uint32_t>(sizeof(CT)), cv + static_cast<size_t>(k) * n_pad + ic);
            }
          }
        }
        if (j >= 0) pt_t2 = next_tile(pt_t2);
      }

      // ---- T0: request the coordinates (error mode: and the frozen correspondence) of batch it + 3 ----
      {
        const int j = it + 3;
        if (j < nb) {
          const uint32_t xs = wb + L::kX + ((ub + 3) % kXS) * L::kXSlot;
#pragma unroll
          for (int p = 0; p < kPPL; p++) {
            const uint32_t e = p * 32u + lane;
            const uint32_t i = pt_t0 * kTile + lane_off + p * 32u;
            const uint32_t ic = i < n ? i : 0u;  // out-of-range lanes read element 0 (always allocated) and are masked through id = -1
#pragma unroll
            for (int k = 0; k < 3; k++) cp_async_small<static_cast<int>(sizeof(PT))>(xs + (k * L::kN + e) * static_cast<uint32_t>(sizeof(PT)), px + static_cast<size_t>(k) * n_pad + ic);
            if (MODE == MODE_ERROR) cp_async_small<4>(wb + L::kG + ((ub + 3) % L::kGSlots) * L::kGSlot + e * 4u, corr + ic);
          }
        }
        pt_t0 = next_tile(pt_t0);
      }
      cp_async_commit();

      // ---- T3: residual, Jacobian and accumulation of batch it: kPPL independent chains per lane in one basic block ----
      if (it >= 0 && has_records) {
        const unsigned char* rsl = wb_g + L::kR + (ub % kRS) * L::kRSlot;
        const PT* xp = reinterpret_cast<const PT*>(wb_g + L::kX + (ub % kXS) * L::kXSlot) + lane;
        TargetRec T[kPPL];
        SourceCov A[kPPL];
        double u0[kPPL], u1[kPPL], u2[kPPL], vf[kPPL];
#pragma unroll
        for (int p = 0; p < kPPL; p++) {
          const uint32_t e = p * 32u + lane;
          const int id = *reinterpret_cast<const volatile int*>(rsl + L::kRId + e * 4);
          vf[p] = id >= 0 ? 1.0 : 0.0;
          hits += id >= 0 ? 1 : 0;
          const double2* rp = reinterpret_cast<const double2*>(rsl) + e;
          T[p].r01 = rp[0];
          T[p].r23 = rp[L::kN];
          T[p].r45 = rp[2 * L::kN];
          T[p].r67 = rp[3 * L::kN];
          T[p].r89 = rp[4 * L::kN];
          const CT* cp = reinterpret_cast<const CT*>(rsl + L::kRCov) + e;
          A[p].a00 = static_cast<double>(cp[0]);
          A[p].a01 = static_cast<double>(cp[L::kN]);
          A[p].a02 = static_cast<double>(cp[2 * L::kN]);
          A[p].a11 = static_cast<double>(cp[3 * L::kN]);
          A[p].a12 = static_cast<double>(cp[4 * L::kN]);
          A[p].a22 = static_cast<double>(cp[5 * L::kN]);
          const double x = static_cast<double>(xp[p * 32]), y = static_cast<double>(xp[L::kN + p * 32]), z = static_cast<double>(xp[2 * L::kN + p * 32]);
          // u = R p, the same individually rounded operations as the correspondence search
          u0[p] = __dadd_rn(__dadd_rn(__dmul_rn(rm(0), x), __dmul_rn(rm(1), y)), __dmul_rn(rm(2), z));
          u1[p] = __dadd_rn(__dadd_rn(__dmul_rn(rm(3), x), __dmul_rn(rm(4), y)), __dmul_rn(rm(5), z));
          u2[p] = __dadd_rn(__dadd_rn(__dmul_rn(rm(6), x), __dmul_rn(rm(7), y)), __dmul_rn(rm(8), z));
        }
#pragma unroll
        for (int p = 0; p < kPPL; p++) accumulate_point_f<MODE, 0, true>(acc, rl, tt, u0[p], u1[p], u2[p], T[p], A[p], vf[p]);
      }
    }
    acc[28] = static_cast<double>(hits);
    cp_async_wait<0>();
    flush_factor<MODE>(sh, acc, tid, partials, counters, out, SINGLE ? pose.m : (poses_lin + static_cast<size_t>(sh.desc.out_index) * 16), sig);
    tile = run_end;
  }
}

}  // namespace sr
}  // namespace b2
