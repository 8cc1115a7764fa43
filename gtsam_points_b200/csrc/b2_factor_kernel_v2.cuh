// b2_factor_kernel_v2.cuh -- the hot kernel, second form: warp-specialised, every operand of the float64 arithmetic staged
// asynchronously into shared memory.  Included by b2_factors.cu after the shared pieces (FactorDesc, accumulate arithmetic,
// warp_reduce32, epilogue).
//
// Round-1's kernel (b2_factor_kernel_ws.cuh, still used by the kd-tree path) was latency bound: 16 warps per SM, each a
// long dependent chain with its global loads on the scoreboard (probe warps: 48 % long-scoreboard stalls on the streamed
// coordinates and the bucket group; accumulate warps: 11 gathers in front of every batch).  Here
//   * the coordinate STREAM (3 planes, + the frozen correspondences in error mode) arrives through the TMA unit: a dedicated
//     PRODUCER warp (one lane) issues 1-D bulk copies (cp.async.bulk, one per plane and CTA tile of kTile = 512 points,
//     2-4 KB each) into a multi-stage shared-memory ring, completion on `full` mbarriers (complete_tx::bytes); stages come
//     back through `empty` mbarriers on which every probe warp arrives once.  The probe warps compute no stream address.
//   * the BUCKET GROUP of the hash probe (the data-dependent access that bounds the probe side) is requested by the probe
//     warp kSB - 1 tiles ahead with cp.async (LDGSTS: no register, no scoreboard) into a small per-warp staging ring, next
//     to the rotated point R p it belongs to; when the probe warp later matches the group it reads shared memory only.
//     In-flight probes per SM = kP x kWarpPoints x (kSB - 1) without holding a single register for them.
//   * what the arithmetic GATHERS per hit -- the target record (80 B) and the source covariance (6 values) -- is requested
//     by the ACCUMULATE warp for batch k + 1 (cp.async into a private double buffer) while it computes batch k: the float64
//     arithmetic reads nothing but shared memory.  Misses are never fetched.
//   * the pose of single-factor launches is a by-value kernel parameter: DMUL / DFMA take it from uniform registers, and the
//     kernel never touches host memory on its way in.
// What stays on a scoreboard is the bucket group of the hash probe (data-dependent address), kPPL points per lane in flight.
//
// Ring protocol (one ring per probe warp, single producer / single consumer, monotonic 32-bit counters in shared memory,
// 32-byte items (R p, stored position | target id)): `tail` (published with st.release), `head` (released by the accumulate warp
// after it has read a batch), `done` / `ack` (factor-run hand-shake: batches never straddle factors).  The accumulate warp
// takes batches of exactly 32 consecutive hits (the last batch of a run may be shorter), rings are drained in strict
// rotation, cross-warp / cross-CTA sums run in slot order: results are bit-reproducible run to run.
// Liveness: ring capacity >= one batch + one warp tile guarantees that either side can always move.  All waits are bounded (trap, never hang).
//
// This file is included once per kernel configuration (no include guard); the includer defines B2_V2_NAMESPACE and the
// B2_V2_* parameters (see b2_factors.cu).

namespace b2 {
namespace B2_V2_NAMESPACE {

constexpr int kP = B2_V2_PRODUCERS;
constexpr int kC = B2_V2_CONSUMERS;
constexpr int kAux = 4;  // one more warpgroup: warp 0 = coordinate-stream producer, warp 1 = covariance-stream producer, 2-3 exit
constexpr int kThreads = (kP + kC + kAux) * 32;
constexpr int kPPL = B2_V2_PPL;          // points per probe lane and warp tile (independent chains interleaved for ILP)
constexpr int kWarpPoints = 32 * kPPL;   // contiguous source points per probe warp and tile
constexpr int kTile = kP * kWarpPoints;  // source points per CTA tile
constexpr int kRing = B2_V2_RING;        // items per ring (power of two)
constexpr int kSX = B2_V2_XYZ_STAGES;    // coordinate stages (CTA tiles in flight ahead of the probe warps)
constexpr int kSB = B2_V2_BUCKET_STAGES;  // bucket-group stages per probe warp (groups requested kSB - 1 tiles ahead)
constexpr int kOperandFields = 8;        // 16-byte fields per gathered hit: record 5 x | covariance 3 x
constexpr int kRingsPerConsumer = kP / kC;
constexpr uint32_t kBatch = 32u;
static_assert(kP % 4 == 0 && kC % 4 == 0, "setmaxnreg works on warpgroups of 4 warps");
static_assert(kP % kC == 0, "every accumulate warp drains the same number of rings");
static_assert((kRing & (kRing - 1)) == 0 && kRing >= 32 + kWarpPoints, "ring capacity: a batch the accumulate warp can take + a tile's hits");
static_assert(kSX >= 3 && kSB >= 2, "stage counts");
// setmaxnreg moves registers inside the CTA's OWN pool (what the launch allocated: threads x the per-thread count the
// launch bounds give); the SM's unallocated remainder is not available.  A split that asks for more deadlocks the
// accumulate warps in USETMAXREG.TRY_ALLOC.
constexpr int kLaunchRegs = ((65536 / kThreads) / 8) * 8 > 255 ? 248 : ((65536 / kThreads) / 8) * 8;
static_assert(kP * B2_V2_REGS_PRODUCER + kC * B2_V2_REGS_CONSUMER + kAux * B2_V2_REGS_AUX <= (kP + kC + kAux) * kLaunchRegs, "register split exceeds the CTA's pool");
// per-role register count relative to what the launch handed every thread: release, keep, or (blocking until released) acquire
template <int REGS>
__device__ __forceinline__ void set_role_registers() {
  if constexpr (REGS < kLaunchRegs) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
  if constexpr (REGS > kLaunchRegs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}

constexpr unsigned kSpinLimit = 1u << 25;  // polls (with sleeps: >= 2 s) before a wait traps: a protocol bug must not hang the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
// Publication of a ring counter WITHOUT a fence: `st.release` compiles to MEMBAR.ALL.CTA + ST, and that barrier also waits for
// every cp.async (LDGSTS) this thread still has in flight -- the bucket groups / operands requested tiles ahead -- which puts
// the full memory latency back on the critical path of every tile.  What has to be ordered here is shared-memory traffic of
// ONE warp (the item stores / item loads of its lanes, then the counter store of lane 0 after __syncwarp()): the shared-memory
// pipeline executes a warp's accesses in program order, and __syncwarp() orders the lanes, so a volatile store is enough.
// (Round 1 measured the same for `head`; `done` / `ack`, once per factor run, keep release semantics.)
__device__ __forceinline__ void st_publish(uint32_t* p, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
struct Backoff {
  unsigned ns, max_ns, polls;
  __device__ __forceinline__ explicit Backoff(unsigned first_ns, unsigned cap_ns) : ns(first_ns), max_ns(cap_ns), polls(0u) {}
  __device__ __forceinline__ void wait() {
    __nanosleep(ns);
    if (ns < max_ns) ns *= 2u;
    if (++polls > kSpinLimit) __trap();
  }
};
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kC * 32) : "memory"); }

#ifdef B2_V2_TIMING
// development aid: per-warp cycle totals of the last launch: [block][warp][slot]
__device__ unsigned long long g_warp_cycles[160 * 32 * 8];
#define B2_T0(var) const long long var = clock64()
#define B2_TACC(slot, var) tacc[slot] += clock64() - var
#else
#define B2_T0(var)
#define B2_TACC(slot, var)
#endif

// ---- mbarrier / bulk-copy (TMA) / cp.async primitives ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned polls = 0;
  while (!mbar_try_wait(bar, parity)) {  // try_wait suspends in hardware up to a time limit: no explicit sleep needed
    if (++polls > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// global -> shared bulk copy (SASS: UBLKCP), `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

// ---- shared-memory layout ---------------------------------------------------------------------------------------------
// dynamic part: [kSX coordinate stages] (CTA tiles), then one ring per probe warp.  Everything is a multiple of 128 bytes.
template <typename PT, int MODE>
struct XyzStage {
  static constexpr uint32_t kPlane = kTile * sizeof(PT);
  static constexpr uint32_t kCorrOff = 3u * kPlane;
  static constexpr uint32_t kBytes = 3u * kPlane + (MODE == MODE_ERROR ? kTile * 4u : 0u);
};
constexpr uint32_t kRingBytes1 = 2u * kRing * 16u;                                      // ring: (u0,u1) | (u2, position|id), field-major
constexpr uint32_t kStageFields = 2u + kGroup;                                          // bucket stage per point: (u0,u1) (u2,aux) | kGroup buckets
constexpr uint32_t kBucketStageBytes = kStageFields * kWarpPoints * 16u;                // one tile of one probe warp, field-major
constexpr uint32_t kOperandBufBytes = kOperandFields * 32u * 16u;                       // one batch of gathered operands, field-major
// 8-byte coordinates double the stage: one stage less keeps the CTA inside the 227 KB of shared memory
template <typename PT>
struct XStages {
  static constexpr uint32_t value = sizeof(PT) == 8 ? static_cast<uint32_t>(kSX - 1) : static_cast<uint32_t>(kSX);
};
template <typename PT, typename CT, int MODE>
struct Layout {
  static constexpr uint32_t kXyzOff = 0u;
  static constexpr uint32_t kRingOff = XStages<PT>::value * XyzStage<PT, MODE>::kBytes;
  static constexpr uint32_t kBktOff = kRingOff + kP * kRingBytes1;
  static constexpr uint32_t kOpOff = kBktOff + kP * kSB * kBucketStageBytes;
  static constexpr uint32_t kTotal = kOpOff + kC * 2u * kOperandBufBytes;
  static_assert(kRingOff % 128u == 0u, "stage alignment");
};

struct Shared {
  FactorDesc desc;  // accumulate-side copy of the current factor's descriptor (flush / epilogue)
  double red[kC][kAcc];
  double tot[kAcc];
  double A[36], X[36], D[36];
  double R[9], t[3];  // pose the residuals of the current factor run are evaluated at (accumulate side)
  double RL[9];       // rotation of its linearization point (== R when linearizing)
  int flag;
  // ring control words, one set per probe warp (all monotonic counters)
  uint32_t tail[kP];   // items published by the probe warp
  uint32_t head[kP];   // items consumed (operands read) by the accumulate warp
  uint32_t done[kP];   // factor runs completed by the probe warp (tail is final for run e once done == e + 1)
  uint32_t ack[kP];    // factor runs the accumulate warp has finished draining
  double probe_pose[kP][12];   // per probe warp: R (9, row-major) | t (3) of the factor run it is working on (multi-factor launches)
  alignas(8) uint64_t full_x[kSX];   // coordinate stage s has landed (1 arrival + tx bytes)
  alignas(8) uint64_t empty_x[kSX];  // every probe warp is done reading it (kP arrivals)
};

// tiles of CTA c: [c * T / G, (c + 1) * T / G) -- contiguous and balanced; the host uses the same formula for the slots
__host__ __device__ __forceinline__ uint32_t cta_tile_begin(uint32_t c, uint32_t T, uint32_t G) {
  return static_cast<uint32_t>(static_cast<unsigned long long>(c) * T / G);
}

// ---------------------------------------------------------------------------------------------------------------
// Per-factor flush by the accumulate warps: warp butterfly -> cross-warp sum -> fixed slot; the last CTA of the
// factor sums the slots in slot order and runs the epilogue (H_t = X^T A' X, ...).
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void flush_factor(Shared& sh, double (&v)[kAcc], int ctid, double* __restrict__ partials, unsigned int* __restrict__ counters, double* __restrict__ out,
                                             const double* __restrict__ pose_lin, const DoneSignal& sig) {
  constexpr int kCT = kC * 32;
  const int lane = ctid & 31, warp = ctid >> 5;
  const double w = warp_reduce32(v, lane);
  sh.red[warp][lane] = w;
  consumer_barrier();
  const FactorDesc& d = sh.desc;
  const uint32_t slot = blockIdx.x - d.cta_first[MODE];
  bool last;
  if (d.num_slots[MODE] == 1u) {
    // the whole factor lived in this CTA (small factors of a big set): no slot round trip through global memory
    last = true;
    if (ctid < kAcc) {
      double s = sh.red[0][ctid];
#pragma unroll
      for (int k = 1; k < kC; k++) s += sh.red[k][ctid];
      sh.tot[ctid] = s;
    }
    consumer_barrier();
  } else {
    if (warp == 0) {
      double s = sh.red[0][lane];
#pragma unroll
      for (int k = 1; k < kC; k++) s += sh.red[k][lane];
      partials[(static_cast<size_t>(d.slot_begin[MODE]) + slot) * kAcc + lane] = s;
      __threadfence();  // only the writing warp pays for the fence
    }
    consumer_barrier();
    if (ctid == 0) {
      const unsigned int prev = atomicAdd(&counters[d.out_index], 1u);
      sh.flag = (prev == d.num_slots[MODE] - 1u) ? 1 : 0;
    }
    consumer_barrier();
    last = sh.flag != 0;
    if (!last) return;
    // ---- last CTA of this factor ----
    __threadfence();
    {
      double s = 0.0;
      for (uint32_t sl = warp; sl < d.num_slots[MODE]; sl += kC) s += __ldcg(&partials[(static_cast<size_t>(d.slot_begin[MODE]) + sl) * kAcc + lane]);
      sh.red[warp][lane] = s;
    }
    consumer_barrier();
    if (ctid < kAcc) {
      double s = sh.red[0][ctid];
#pragma unroll
      for (int k = 1; k < kC; k++) s += sh.red[k][ctid];
      sh.tot[ctid] = s;
    }
    if (ctid == 0) counters[d.out_index] = 0u;  // re-arm for the next launch
    consumer_barrier();
  }

  if (MODE == MODE_ERROR) {
    if (ctid == 0) {
      out[d.out_index] = sh.tot[27];
      __threadfence_system();  // `out` may be mapped host memory
      signal_done(sig);
    }
    consumer_barrier();
    return;
  }
  epilogue_build(sh.A, sh.X, sh.D, sh.tot, sh.R, sh.t, ctid);
  consumer_barrier();
  double* rec = out + static_cast<size_t>(d.out_index) * B2_LINEARIZED_DOUBLES;
  epilogue_store(rec, sh.A, sh.X, sh.D, sh.tot, ctid);
  if (ctid >= 100 && ctid < 116) {
    // remember the linearization point with the factor (error-only launches of ANY set read it back)
    d.lin_pose[ctid - 100] = pose_lin[ctid - 100];
  }
  if (sig.n_peers > 1) {
    // multi-GPU exchange fused into the epilogue: the finished record goes into the same slot of every peer's buffer.
    // All peers' stores are issued before the single system-scope fence below (they travel over NVLink concurrently).
    __syncwarp();
    consumer_barrier();  // record complete in this GPU's buffer (visible CTA-wide)
    for (int i = ctid; i < B2_LINEARIZED_DOUBLES * (sig.n_peers - 1); i += kCT) {
      int p = i / B2_LINEARIZED_DOUBLES;
      const int k = i - p * B2_LINEARIZED_DOUBLES;
      if (p >= sig.my_rank) p++;
      sig.peer_out[p][static_cast<size_t>(d.out_index) * B2_LINEARIZED_DOUBLES + k] = rec[k];
    }
  }
  __threadfence_system();  // `out` may be mapped host memory (zero-copy host API) / peer memory
  consumer_barrier();
  if (ctid == 0) signal_done(sig);  // every writer of this record fenced before the barrier
  (void)kCT;
}

// ---------------------------------------------------------------------------------------------------------------
// The kernel.  KIND: 0 = VGICP (voxel hash probe), 1 = GICP (kd-tree 1-NN).  MODE: linearize / error-only.
// SINGLE: the launch covers exactly one factor and its pose is the by-value parameter `pose` (linearize: the
// linearization point; error: the evaluation point); otherwise poses are read from poses_lin / poses_eval.
// ---------------------------------------------------------------------------------------------------------------
template <typename PT, typename CT, int KIND, int MODE, bool SINGLE>
__global__ void __launch_bounds__(kThreads, 1)
factor_kernel(const FactorDesc* __restrict__ descs, const uint32_t* __restrict__ tile_factor, uint32_t num_tiles, const double* __restrict__ poses_lin,
              const double* __restrict__ poses_eval, double* __restrict__ partials, unsigned int* __restrict__ counters, double* __restrict__ out,
              const __grid_constant__ DoneSignal sig, const __grid_constant__ PoseArg pose, const uint32_t* __restrict__ /*frozen_flags: kd-tree factors only*/) {
  using L = Layout<PT, CT, MODE>;
  using XS = XyzStage<PT, MODE>;
  constexpr uint32_t kSXp = XStages<PT>::value;  // coordinate stages of this instantiation
  __shared__ Shared sh;
  extern __shared__ __align__(128) unsigned char dyn_smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kP) {
    sh.tail[tid] = 0u;
    sh.head[tid] = 0u;
    sh.done[tid] = 0u;
    sh.ack[tid] = 0u;
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kSX; s++) {
      mbar_init(&sh.full_x[s], 1u);
      mbar_init(&sh.empty_x[s], kP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t G = gridDim.x;
  const uint32_t tile_lo = cta_tile_begin(blockIdx.x, num_tiles, G);
  const uint32_t tile_hi = cta_tile_begin(blockIdx.x + 1, num_tiles, G);

  if (warp >= kP + kC) {
    // ================================================= PRODUCER warp =================================================
    // aux warp 0, one lane: streams the coordinate planes (+ frozen correspondences) of the CTA's tiles, running ahead of the
    // probe warps as far as the stages allow -- across factor boundaries, too.  (setmaxnreg works on warpgroups: the other
    // three warps of this group only give their registers back.)
    set_role_registers<B2_V2_REGS_AUX>();
#if B2_V2_WARM_L2
    if (warp != kP + kC) {
      // The other three warps of this group stream the GATHERED arrays of the CTA's first factor -- bucket table and target
      // records, a few MB that every point probes at random -- into L2, each CTA its share, with coalesced prefetches: the
      // random first touches of the probe / operand gathers then find L2 lines instead of paying a DRAM round trip each.
      const FactorDesc* __restrict__ dg = descs + (SINGLE ? 0u : __ldg(tile_factor + tile_lo));
      const uint32_t share = dg->num_slots[MODE], me = blockIdx.x - dg->cta_first[MODE];
      const int w = warp - (kP + kC) - 1;  // 0..2
      auto warm = [&](const char* ptr, size_t bytes) {
        const size_t lines = (bytes + 127) / 128;
        const size_t per = (lines + share - 1) / share;
        const size_t lo = static_cast<size_t>(me) * per, hi = min(lines, lo + per);
        for (size_t l = lo + static_cast<size_t>(w) * 32 + lane; l < hi; l += 96) prefetch_l2(ptr + l * 128);
      };
      if (me < share) {
        if (KIND == 0) warm(reinterpret_cast<const char*>(dg->buckets), (static_cast<size_t>(dg->bucket_mask) + 1) * kGroup * sizeof(VoxelBucket));
#if B2_V2_WARM_L2 > 1
        if (KIND == 0) warm(reinterpret_cast<const char*>(dg->records), static_cast<size_t>(dg->num_records) * kRecordDoubles * sizeof(double));
#endif
      }
      return;
    }
    if (lane != 0) return;
#else
    if (warp != kP + kC || lane != 0) return;
#endif
    uint32_t j = 0u;  // CTA tiles issued so far: stage = j % kSXp, use count = j / kSXp
    uint32_t tile = tile_lo;
#ifdef B2_V2_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
#endif
    while (tile < tile_hi) {
      const FactorDesc* __restrict__ dg = descs + (SINGLE ? 0u : __ldg(tile_factor + tile));
      const uint32_t n_pad = dg->n_pad;
      const uint32_t f_tile_begin = dg->tile_begin, f_num_tiles = dg->num_tiles, perm_stride = dg->perm_stride;
      const uint32_t run_end = min(tile_hi, f_tile_begin + f_num_tiles);
      const PT* __restrict__ px = static_cast<const PT*>(dg->pts);
      const int32_t* __restrict__ corr = dg->corr;
      uint32_t pt = static_cast<uint32_t>(static_cast<unsigned long long>(tile - f_tile_begin) * perm_stride % f_num_tiles);
      for (; tile < run_end; tile++, j++) {
        const uint32_t base = pt * kTile;
        const uint32_t cnt = base >= n_pad ? 0u : min(static_cast<uint32_t>(kTile), n_pad - base);  // multiple of 32
        const uint32_t s = j % kSXp;
        {
          B2_T0(tw);
          if (j >= kSXp) mbar_wait(&sh.empty_x[s], ((j / kSXp) - 1u) & 1u);
          B2_TACC(0, tw);
        }
        fence_proxy_async();  // the stage was read through the generic proxy
        const uint32_t dst = smem_u32(dyn_smem + L::kXyzOff + s * XS::kBytes);
        mbar_expect_tx(&sh.full_x[s], cnt * (3u * static_cast<uint32_t>(sizeof(PT)) + (MODE == MODE_ERROR ? 4u : 0u)));
        if (cnt) {
          const uint32_t bytes = cnt * static_cast<uint32_t>(sizeof(PT));
#pragma unroll
          for (int a = 0; a < 3; a++) bulk_g2s(dst + a * XS::kPlane, px + static_cast<size_t>(a) * n_pad + base, bytes, &sh.full_x[s]);
          if (MODE == MODE_ERROR) bulk_g2s(dst + XS::kCorrOff, corr + base, cnt * 4u, &sh.full_x[s]);
        }
        pt += perm_stride;
        if (pt >= f_num_tiles) pt -= f_num_tiles;
      }
    }
#ifdef B2_V2_TIMING
    tacc[7] = clock64() - t_begin;
    for (int k = 0; k < 8; k++) g_warp_cycles[(blockIdx.x * 32 + warp) * 8 + k] = tacc[k];
#endif
    return;
  }

  if (warp < kP) {
    // ================================================= PROBE warps =================================================
    set_role_registers<B2_V2_REGS_PRODUCER>();
    const int p = warp;
    double2* const ring = reinterpret_cast<double2*>(dyn_smem + L::kRingOff + static_cast<size_t>(p) * kRingBytes1);
    unsigned char* const bkt = dyn_smem + L::kBktOff + static_cast<size_t>(p) * kSB * kBucketStageBytes;
    const uint32_t bkt_s = smem_u32(bkt);
    uint32_t tail = 0u, run = 0u;
    uint32_t j = 0u;  // CTA tiles this warp has MATCHED so far (all runs): bucket stage = j % kSB; phase A runs kSB - 1 tiles ahead
    uint32_t head_seen = 0u;
    uint32_t tile = tile_lo;
#ifdef B2_V2_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
#endif
    while (tile < tile_hi) {
      const FactorDesc* __restrict__ dg = descs + (SINGLE ? 0u : __ldg(tile_factor + tile));
      const uint32_t n = dg->n;
      const uint32_t f_tile_begin = dg->tile_begin;
      const uint32_t run_end = min(tile_hi, f_tile_begin + dg->num_tiles);
      const uint32_t out_index = dg->out_index;
      int32_t* __restrict__ corr = dg->corr;
      const VoxelBucket* __restrict__ buckets = dg->buckets;
      const uint32_t bucket_mask = dg->bucket_mask;
      const double inv_leaf = dg->inv_leaf;
      if (!SINGLE) {
        __syncwarp();
        if (lane < 12) {
          const double* pe = (MODE == MODE_ERROR ? poses_eval : poses_lin) + static_cast<size_t>(out_index) * 16;
          sh.probe_pose[p][lane] = __ldg(pe + (lane < 9 ? (lane / 3) * 4 + lane % 3 : (lane - 9) * 4 + 3));
        }
        __syncwarp();
      }
      // pose the correspondences are searched at: uniform-register operands (SINGLE) or this warp's shared-memory slot
      auto Rm = [&](int i) -> double { return SINGLE ? pose.m[(i / 3) * 4 + (i % 3)] : sh.probe_pose[p][i]; };
      auto tv = [&](int i) -> double { return SINGLE ? pose.m[i * 4 + 3] : sh.probe_pose[p][9 + i]; };
      const KdTreeView tview{dg->nodes, dg->leaf_pts, static_cast<int>(dg->leaf_f32)};
      const double max_sq = dg->max_sq;

      // Virtual tile v of the factor (the CTA owns a contiguous range of them) is physical tile (v * S) mod n_tiles with
      // S ~ 0.618 n_tiles coprime to n_tiles: every CTA samples the (Morton-ordered) cloud quasi-uniformly.
      const uint32_t f_num_tiles = dg->num_tiles, perm_stride = dg->perm_stride;
      auto next_tile = [&](uint32_t pt) {
        pt += perm_stride;
        return pt >= f_num_tiles ? pt - f_num_tiles : pt;
      };
      const uint32_t K = run_end - tile;  // tiles of this run
      uint32_t pt_a = static_cast<uint32_t>(static_cast<unsigned long long>(tile - f_tile_begin) * perm_stride % f_num_tiles), pt_b = pt_a;

      // Phase A(jj): the tile's coordinates (TMA stage jj % kSXp) -> R p -> voxel coordinate -> hash; R p goes into bucket stage
      // jj % kSB, the bucket group is requested into the same slot (cp.async), the coordinate stage goes back to the producer.
      auto phase_a = [&](uint32_t jj) {
        const uint32_t sx = jj % kSXp;
        {
          B2_T0(tw);
          mbar_wait(&sh.full_x[sx], (jj / kSXp) & 1u);
          B2_TACC(1, tw);
        }
        B2_T0(t_a);
        const unsigned char* xs = dyn_smem + L::kXyzOff + sx * XS::kBytes;
        double2* st = reinterpret_cast<double2*>(bkt + (jj % kSB) * kBucketStageBytes);
        const uint32_t st_s = bkt_s + (jj % kSB) * kBucketStageBytes;
#pragma unroll
        for (int q = 0; q < kPPL; q++) {
          const int lw = lane + 32 * q;             // position inside the warp tile
          const int li = p * kWarpPoints + lw;      // position inside the CTA tile
          const double x = static_cast<double>(reinterpret_cast<const PT*>(xs)[li]);
          const double y = static_cast<double>(reinterpret_cast<const PT*>(xs + XS::kPlane)[li]);
          const double z = static_cast<double>(reinterpret_cast<const PT*>(xs + 2 * XS::kPlane)[li]);
          // u = R p : coefficient sums in index order, each operation individually rounded (bit-parity with the CPU float64 path)
          const double u0 = __dadd_rn(__dadd_rn(__dmul_rn(Rm(0), x), __dmul_rn(Rm(1), y)), __dmul_rn(Rm(2), z));
          const double u1 = __dadd_rn(__dadd_rn(__dmul_rn(Rm(3), x), __dmul_rn(Rm(4), y)), __dmul_rn(Rm(5), z));
          const double u2 = __dadd_rn(__dadd_rn(__dmul_rn(Rm(6), x), __dmul_rn(Rm(7), y)), __dmul_rn(Rm(8), z));
          int aux = -1;
          if (MODE == MODE_ERROR) aux = reinterpret_cast<const int*>(xs + XS::kCorrOff)[li];  // the frozen correspondence
          st[lw] = make_double2(u0, u1);
          st[kWarpPoints + lw] = make_double2(u2, __longlong_as_double(static_cast<long long>(aux)));
          if (MODE == MODE_LINEARIZE && KIND == 0) {
            const int cx = voxel_coord1(__dadd_rn(u0, tv(0)), inv_leaf);
            const int cy = voxel_coord1(__dadd_rn(u1, tv(1)), inv_leaf);
            const int cz = voxel_coord1(__dadd_rn(u2, tv(2)), inv_leaf);
            const uint32_t g = voxel_hash(cx, cy, cz) & bucket_mask;
#ifndef B2_V2_DEBUG_NO_PROBE
            const VoxelBucket* gp = buckets + static_cast<size_t>(g) * kGroup;
#pragma unroll
            for (int k = 0; k < kGroup; k++) cp_async16(st_s + (2u + k) * (kWarpPoints * 16u) + lw * 16u, gp + k);
#endif
          }
        }
        cp_async_commit();
        // every lane has read its coordinates: the stage goes back to the producer (one arrival per probe warp)
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty_x[sx]);
        B2_TACC(2, t_a);
      };

      // prologue: kSB - 1 tiles ahead (empty commit groups keep the group count uniform for short runs)
#pragma unroll 1
      for (uint32_t a = 0; a < static_cast<uint32_t>(kSB - 1); a++) {
        if (a < K) {
          phase_a(j + a);
          pt_a = next_tile(pt_a);
        } else {
          cp_async_commit();
        }
      }

#pragma unroll 1
      for (uint32_t k = 0; k < K; k++, tile++, j++) {
        if (k + kSB - 1 < K) {
          phase_a(j + kSB - 1);
          pt_a = next_tile(pt_a);
        } else {
          cp_async_commit();
        }
        // ---- Phase B(j): the tile whose bucket groups were requested kSB - 1 tiles ago ----
        {
          B2_T0(tw);
          cp_async_wait<kSB - 1>();  // all but the newest kSB - 1 groups have landed: this lane's own requests of tile j included
          B2_TACC(4, tw);
        }
        B2_T0(t_b);
        const double2* st = reinterpret_cast<const double2*>(bkt + (j % kSB) * kBucketStageBytes);
        const uint32_t base = pt_b * kTile + p * kWarpPoints + lane;
        pt_b = next_tile(pt_b);
        double u[kPPL][3];
        int id[kPPL];
        bool ok[kPPL];
        uint32_t mask[kPPL], cnt = 0u;
#pragma unroll
        for (int q = 0; q < kPPL; q++) {
          const int lw = lane + 32 * q;
          ok[q] = base + 32 * q < n;
          const double2 q0 = st[lw], q1 = st[kWarpPoints + lw];
          u[q][0] = q0.x, u[q][1] = q0.y, u[q][2] = q1.x;
          id[q] = -1;
          if (MODE == MODE_ERROR) id[q] = ok[q] ? static_cast<int>(__double_as_longlong(q1.y)) : -1;
          if (MODE == MODE_LINEARIZE) {
            if (KIND == 0) {
              // the same individually rounded operations as phase A: identical voxel coordinate
              const int cx = voxel_coord1(__dadd_rn(u[q][0], tv(0)), inv_leaf);
              const int cy = voxel_coord1(__dadd_rn(u[q][1], tv(1)), inv_leaf);
              const int cz = voxel_coord1(__dadd_rn(u[q][2], tv(2)), inv_leaf);
#ifdef B2_V2_DEBUG_NO_PROBE
              id[q] = static_cast<int>((voxel_hash(cx, cy, cz) & bucket_mask) & 0xffffu);  // measurement aid: no table access, every point "hits" some record
#else
              BucketGroup grp;
#pragma unroll
              for (int k = 0; k < kGroup; k++) grp.b[k] = reinterpret_cast<const int4*>(st)[(2 + k) * kWarpPoints + lw];
              id[q] = match_group(grp, cx, cy, cz);
              if (id[q] == -2) {  // rare: the home group is full and does not hold the key -- walk on (synchronous loads)
                uint32_t g = voxel_hash(cx, cy, cz) & bucket_mask;
                do {
                  g = (g + 1) & bucket_mask;
                  id[q] = match_group(load_group(buckets, g), cx, cy, cz);
                } while (id[q] == -2);
              }
#endif
            } else {
              double sq;
              id[q] = kdtree_nn1_warp(tview, __dadd_rn(u[q][0], tv(0)), __dadd_rn(u[q][1], tv(1)), __dadd_rn(u[q][2], tv(2)), ok[q], max_sq, &sq);
            }
            if (ok[q]) corr[base + 32 * q] = id[q];
          }
          mask[q] = __ballot_sync(0xffffffffu, ok[q] && id[q] >= 0);
          cnt += __popc(mask[q]);
        }
        B2_TACC(5, t_b);
        if (cnt != 0u) {
          B2_T0(t_ring);
          // room in the ring
          if (tail + cnt - head_seen > static_cast<uint32_t>(kRing)) {
            head_seen = __shfl_sync(0xffffffffu, ld_acquire(&sh.head[p]), 0);
            if (tail + cnt - head_seen > static_cast<uint32_t>(kRing)) {
              Backoff bo(B2_V2_BACKOFF_MIN, B2_V2_BACKOFF_MAX);
              do {
                bo.wait();
                head_seen = __shfl_sync(0xffffffffu, ld_acquire(&sh.head[p]), 0);
              } while (tail + cnt - head_seen > static_cast<uint32_t>(kRing));
            }
          }
#pragma unroll
          for (int q = 0; q < kPPL; q++) {
            if ((mask[q] >> lane) & 1u) {
              const uint32_t slot = (tail + __popc(mask[q] & ((1u << lane) - 1u))) & (kRing - 1);
              const unsigned long long bits = static_cast<unsigned long long>(base + 32 * q) | (static_cast<unsigned long long>(static_cast<uint32_t>(id[q])) << 32);
              ring[slot] = make_double2(u[q][0], u[q][1]);
              ring[kRing + slot] = make_double2(u[q][2], __longlong_as_double(static_cast<long long>(bits)));
            }
            tail += __popc(mask[q]);
          }
          __syncwarp();
          if (lane == 0) st_publish(&sh.tail[p], tail);
          B2_TACC(3, t_ring);
        }
      }
      // end of this CTA's run of the factor: publish, then wait until the accumulate warp has drained the ring
      cp_async_wait<0>();
      run++;
      __syncwarp();
      if (lane == 0) st_release(&sh.done[p], run);
      {
        B2_T0(tw);
        Backoff bo(128u, 512u);
        while (ld_acquire(&sh.ack[p]) != run) bo.wait();
        B2_TACC(6, tw);
      }
      head_seen = tail;
      __syncwarp();
    }
#ifdef B2_V2_TIMING
    tacc[7] = clock64() - t_begin;
    if (lane == 0)
      for (int k = 0; k < 8; k++) g_warp_cycles[(blockIdx.x * 32 + warp) * 8 + k] = tacc[k];
#endif
  } else {
    // =============================================== ACCUMULATE warps ===============================================
    set_role_registers<B2_V2_REGS_CONSUMER>();
    const int cw = warp - kP;        // accumulate warp index
    const int ctid = tid - kP * 32;  // thread index within the accumulate group
    unsigned char* const opbuf = dyn_smem + L::kOpOff + static_cast<size_t>(cw) * 2u * kOperandBufBytes;
    const uint32_t opbuf_s = smem_u32(opbuf);
    uint32_t head[kRingsPerConsumer];  // items taken (fetched) from each of this warp's rings
#pragma unroll
    for (int r = 0; r < kRingsPerConsumer; r++) head[r] = 0u;
    uint32_t run = 0u;
    uint32_t tile = tile_lo;
    double acc[kAcc];
#ifdef B2_V2_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
#endif
    while (tile < tile_hi) {
      const uint32_t f = SINGLE ? 0u : __ldg(tile_factor + tile);
      consumer_barrier();  // previous flush is done with sh.desc
      if (ctid < static_cast<int>(sizeof(FactorDesc) / 4)) reinterpret_cast<uint32_t*>(&sh.desc)[ctid] = __ldg(reinterpret_cast<const uint32_t*>(descs + f) + ctid);
      consumer_barrier();
      const FactorDesc& d = sh.desc;
      const double* pe = SINGLE ? pose.m : ((MODE == MODE_ERROR ? poses_eval : poses_lin) + static_cast<size_t>(d.out_index) * 16);
      if (ctid < 21) {
        const double* pl = (MODE == MODE_ERROR) ? d.lin_pose : pe;
        if (ctid < 9)
          sh.R[ctid] = pe[(ctid / 3) * 4 + ctid % 3];
        else if (ctid < 12)
          sh.t[ctid - 9] = pe[(ctid - 9) * 4 + 3];
        else
          sh.RL[ctid - 12] = pl[((ctid - 12) / 3) * 4 + (ctid - 12) % 3];
      }
      consumer_barrier();
      // pose operands of the arithmetic: uniform registers where the launch allows it, else registers
      constexpr bool kConstRL = SINGLE && MODE == MODE_LINEARIZE;
      double RLr[kConstRL ? 1 : 9], tr[SINGLE ? 1 : 3];
      if (!kConstRL) {
#pragma unroll
        for (int k = 0; k < 9; k++) RLr[k] = sh.RL[k];
      }
      if (!SINGLE) {
#pragma unroll
        for (int k = 0; k < 3; k++) tr[k] = sh.t[k];
      }
      auto rl = [&](int i) -> double { return kConstRL ? pose.m[(i / 3) * 4 + (i % 3)] : RLr[kConstRL ? 0 : i]; };
      auto tt = [&](int i) -> double { return SINGLE ? pose.m[i * 4 + 3] : tr[SINGLE ? 0 : i]; };
#pragma unroll
      for (int k = 0; k < kAcc; k++) acc[k] = 0.0;
      const uint32_t run_end = min(tile_hi, d.tile_begin + d.num_tiles);
      run++;

      // One batch = up to 32 consecutive hits of one ring (always 32 except for a probe warp's last batch of the run).
      // `cur` is being computed, `nxt` has been fetched: its gathered operands -- target record (5 x 16 B) and source
      // covariance (6 cells of 8 B) per hit -- are on their way into the other half of this warp's operand buffer (cp.async:
      // no register, no scoreboard).  Strict rotation over the warp's rings keeps the batch sequence a function of the data.
      const double* __restrict__ records = d.records;
      const CT* __restrict__ cv = static_cast<const CT*>(d.covs);
      const size_t n_pad = d.n_pad;
      struct Bat {
        uint32_t hd, nb;
        int r;
        bool fin, valid;
      };
      constexpr uint32_t kAllFinished = (1u << kRingsPerConsumer) - 1u;
      uint32_t finished = 0u;  // rings whose final batch of this run has been FETCHED
      int rot = 0;             // next ring in the strict rotation
      uint32_t buf = 0u;
      auto fetch = [&](int r, bool blocking, Bat& bt, uint32_t which) -> bool {
        const int p = cw + r * kC;
        uint32_t hd = 0u;
#pragma unroll
        for (int k = 0; k < kRingsPerConsumer; k++)
          if (k == r) hd = head[k];
        uint32_t nb = 0u;
        bool fin = false;
        Backoff bo(B2_V2_BACKOFF_MIN, B2_V2_BACKOFF_MAX);
        while (true) {
          const uint32_t dn = ld_acquire(&sh.done[p]);
          const uint32_t tl = ld_acquire(&sh.tail[p]);
          const uint32_t avail = tl - hd;
          if (avail >= kBatch) {
            nb = kBatch;
            break;
          }
          if (dn == run) {  // the probe warp finished this run: `tl` is final
            nb = avail;
            fin = true;
            break;
          }
          if (!blocking) return false;
          bo.wait();
        }
        bt.hd = hd, bt.nb = nb, bt.r = r, bt.fin = fin, bt.valid = true;
#pragma unroll
        for (int k = 0; k < kRingsPerConsumer; k++)
          if (k == r) head[k] = hd + nb;
        if (fin) finished |= 1u << r;
        if (static_cast<uint32_t>(lane) < nb) {
          const double2* rg = reinterpret_cast<const double2*>(dyn_smem + L::kRingOff + static_cast<size_t>(p) * kRingBytes1);
          const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(rg[kRing + ((hd + lane) & (kRing - 1))].y));
          const char* rec = reinterpret_cast<const char*>(records + static_cast<size_t>(bits >> 32) * kRecordDoubles);
          const CT* cp = cv + static_cast<uint32_t>(bits);
          const uint32_t dst = opbuf_s + which * kOperandBufBytes + lane * 16u;
#pragma unroll
          for (int k = 0; k < 5; k++) cp_async16(dst + k * 512u, rec + 16 * k);
#pragma unroll
          for (int k = 0; k < 6; k++) cp_async_small<static_cast<int>(sizeof(CT))>(dst + (5u + (k >> 1)) * 512u + (k & 1) * 8u, cp + static_cast<size_t>(k) * n_pad);
        }
        cp_async_commit();
        return true;
      };
      auto next_ring = [&](int from) -> int {  // first unfinished ring at or after `from` in the rotation (-1: none)
#pragma unroll
        for (int k = 0; k < kRingsPerConsumer; k++) {
          const int r = (from + k) % kRingsPerConsumer;
          if (!(finished & (1u << r))) return r;
        }
        return -1;
      };

      Bat cur{0u, 0u, 0, false, false}, nxt{0u, 0u, 0, false, false};
      while (true) {
        if (!cur.valid) {
          const int r = next_ring(rot);
          if (r < 0) break;
          B2_T0(tw);
          fetch(r, true, cur, buf);
          B2_TACC(0, tw);
          rot = (r + 1) % kRingsPerConsumer;
        }
        // get the following batch on its way before computing this one
        nxt.valid = false;
        {
          const int r = next_ring(rot);
          if (r >= 0 && fetch(r, false, nxt, buf ^ 1u)) rot = (r + 1) % kRingsPerConsumer;
        }
        {
          B2_T0(t_cp);
          if (nxt.valid)
            cp_async_wait<1>();
          else
            cp_async_wait<0>();
          B2_TACC(1, t_cp);
        }
        B2_T0(t_comp);
#ifdef B2_V2_TIMING
        tacc[5]++;
        if (nxt.valid) tacc[4]++;
#endif
        {
          const int p = cw + cur.r * kC;
          const bool valid = static_cast<uint32_t>(lane) < cur.nb;
          if (valid) {  // uniform for full batches
            const double2* rg = reinterpret_cast<const double2*>(dyn_smem + L::kRingOff + static_cast<size_t>(p) * kRingBytes1) + ((cur.hd + lane) & (kRing - 1));
            const double2 q0 = rg[0], q1 = rg[kRing];
#ifndef B2_V2_DEBUG_NO_ACCUM
            const double2* ob = reinterpret_cast<const double2*>(opbuf + buf * kOperandBufBytes) + lane;
            TargetRec T;
            T.r01 = ob[0];
            T.r23 = ob[32];
            T.r45 = ob[64];
            T.r67 = ob[96];
            T.r89 = ob[128];
            const double2 c0 = ob[160], c1 = ob[192], c2 = ob[224];
            auto cell = [](double v) -> double {  // an 8-byte cell holds a double, or a float in its low half
              return sizeof(CT) == 8 ? v : static_cast<double>(__int_as_float(__double2loint(v)));
            };
            SourceCov A;
            A.a00 = cell(c0.x);
            A.a01 = cell(c0.y);
            A.a02 = cell(c1.x);
            A.a11 = cell(c1.y);
            A.a12 = cell(c2.x);
            A.a22 = cell(c2.y);
            accumulate_point_f<MODE>(acc, rl, tt, q0.x, q0.y, q1.x, T, A);
#else
            acc[28] += q0.x * 0.0 + 1.0 + q1.x * 0.0;  // measurement aid: probe-side throughput only
#endif
          }
          // ring slots read: hand them back to the probe warp
          __syncwarp();
          if (lane == 0) {
            st_publish(&sh.head[p], cur.hd + cur.nb);
            if (cur.fin) st_release(&sh.ack[p], run);
          }
        }
        B2_TACC(2, t_comp);
        cur = nxt;
        buf ^= 1u;
      }
      B2_T0(t_fl);
      flush_factor<MODE>(sh, acc, ctid, partials, counters, out, pe, sig);
      B2_TACC(3, t_fl);
      tile = run_end;
    }
#ifdef B2_V2_TIMING
    tacc[7] = clock64() - t_begin;
    if (lane == 0)
      for (int k = 0; k < 8; k++) g_warp_cycles[(blockIdx.x * 32 + warp) * 8 + k] = tacc[k];
#endif
  }
}

}  // namespace B2_V2_NAMESPACE
}  // namespace b2
