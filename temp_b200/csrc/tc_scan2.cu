// temp_b200 -- gru_scan_tm_kernel: the chain-partitioned GRU scan with W_hh resident in TENSOR MEMORY, one or two
// independent partition pipelines per CTA and the state handed from step to step through DISTRIBUTED SHARED MEMORY
// (round 2; replaces gru_scan_tc_kernel for scans that use ONE recurrent cell).
//
// What the recurrence is (reference models/RRGCN.py:77-89, models/DynamicRGCN.py:156-174): per window step,
//     h0 = decay(state[prev_row[r]]);  gh = h0 . W_hh^T + b_hh;  gates with the precomputed gi;  state[r] = h' (+ te)
// where prev_row links a packed row only to the row of the same entity in the same batch item one step earlier, so the
// planner's chain partitions (entity-id ranges, <= 48 rows per step) are independent dependency chains.
//
// Why this shape (round-1 kernel: one <= 96-row tile in flight per 4-CTA cluster, W_hh as a 128 KB shared-memory image,
// 8.7 % / 10 % of the HBM roofline, the step chain = cluster barrier -> L2 gather -> MMA -> gates -> release fence):
//   * a step of ONE chain is latency, not work -- throughput comes from running several chains at once.  A CTA runs
//     kPipes pipelines (16 / kPipes warps each); a pipeline owns its operand tile, its staging tiles, its TMEM
//     accumulator and its barriers, issues its own tcgen05.mma batch and never synchronises with the other pipeline
//     (named barriers + mbarriers, no __syncthreads and no cluster barrier in the loop).
//   * W_hh lives in TMEM as the A operand (tcgen05.mma "TS" form): 128 lanes (r|z|n rows of this CTA's 32 hidden
//     columns, 32 pad rows) x 128 k columns, hi and lo tf32 parts = 256 of the 512 columns.  An MMA then reads only its
//     N x 8 activation slice from shared memory, which is what makes N <= 48 batches run at the tensor-pipe floor
//     (34.7 cycles per MMA, tools/probes/mma_probe.cu), and the shared memory the weight image used to take holds the
//     pipelines' tiles.  The slice is fetched by bulk copies during the first tile step (which multiplies nothing).
//   * the state never goes through L2 between the steps of a partition: every CTA stores its 32 columns of the new
//     state into the fp32 staging tile of all four CTAs with st.async; the consumer's mbarrier counts the bytes (no
//     fence, no cluster barrier).  One staging tile per pipeline: a tile step converts it into the operand first, then
//     reports it vacant to the four CTAs; senders wait for the four reports of the step before they store (so a fast CTA
//     never writes a tile a slow peer still reads, within a partition or across a partition switch).
//   * gate exchange without a second tile: the accumulator holds gate g on TMEM lane quadrant g (rows on columns); the
//     three gate warps park their quadrant in the k-atoms of the dead operand image OTHER than this CTA's own.
//
// Layout conventions are those of tc_common.cuh: D^T[feature, row] ("features on TMEM lanes"), B = activations K-major
// SWIZZLE_128B in shared memory, 3xTF32 operand split.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr unsigned kFull = 0xffffffffu;
constexpr int kD = 128;
constexpr int kCluster = 4;                       // CTAs per cluster = blocks of 32 hidden columns
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;             // 512 -> 128 registers per thread
// kR = max packed rows per partition step = UMMA N: 48 (the widest N at the tensor-pipe floor: latency regime, few
// partitions) or 64 (chained scans with more partitions than pipelines: 25 % fewer tile steps for 20 % more MMA time each).
// Per pipeline 96 KB: chained = operand tile (hi | lo image, 2 x kR x 512 bytes) at +0, ONE fp32 staging tile at +64 KB;
// plain (kR = 48 only) = two operand tiles of 48 KB that alternate.
constexpr int kRowsPlain = 48;
constexpr int kRowsMax = 64;
constexpr int kPipeSmem = 96 * 1024;
constexpr int kStageOff = 64 * 1024;
constexpr int kMaxPipes = 2;
constexpr int kWExtra = 32 * 1024;                // with 48 KB of each pipeline's operand region: room for the 128 KB W_hh slice
constexpr int kSmem = kMaxPipes * kPipeSmem + kWExtra + 1024;
static_assert(2 * 3 * 16384 + kWExtra == 4 * kWChunkBytes, "the W_hh slice lands in the operand regions + the extra block");
static_assert(kSmem <= 227 * 1024, "shared memory per CTA");
constexpr int kTmemCols = 512;
constexpr int kColWLo = 128;                      // W_hh hi parts at columns [0, 128), lo parts at [128, 256)
constexpr int kColD = 256, kColDStride = 64;      // accumulator of pipeline p: columns [256 + 64 p, + kR)
static_assert(kColD + (kMaxPipes - 1) * kColDStride + kRowsMax <= kTmemCols, "TMEM budget");

struct Bars {
  uint64_t w_full;                // the W_hh slice has landed in shared memory (bulk copies, bytes counted)
  uint64_t mma_done[kMaxPipes];
  uint64_t ready[kMaxPipes];      // the staging tile of a pipeline is complete (bytes counted: st.async from the four CTAs)
  uint64_t vacant[kMaxPipes];     // all four CTAs have finished reading their staging tile of the current tile step
  uint32_t tmem_base;
  float* push[TEMP_MAX_PUSH_PEERS];
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}
__device__ __forceinline__ float decay_factor(float dt, const float* wb, float inv_temperature) {
  if (wb != nullptr) return expf(-fmaxf(fmaf(__ldg(wb), dt, __ldg(wb + 1)), 0.f));
  return expf(-dt * inv_temperature);
}
// ex2.approx / rcp.approx based gate math, absolute error ~1e-7 (as gru_scan_tc_kernel)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Development-only phase timeline (tools/probe_timeline2.py builds a separate library with -DTEMP_TIMELINE): lane 0 of
// every warp stores clock64() per phase slot; slot 63 holds %globaltimer at kernel entry.
#ifdef TEMP_TIMELINE
__device__ unsigned long long* g_timeline2 = nullptr;
constexpr int kTlWarps = 16, kTlSlots = 64;
__device__ __forceinline__ void tl_mark(int slot) {
  if (g_timeline2 != nullptr && (threadIdx.x & 31) == 0 && slot < 63)
    g_timeline2[(static_cast<size_t>(blockIdx.x) * kTlWarps + (threadIdx.x >> 5)) * kTlSlots + slot] = clock64();
}
__device__ __forceinline__ void tl_start() {
  if (g_timeline2 != nullptr && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_timeline2[(static_cast<size_t>(blockIdx.x) * kTlWarps + (threadIdx.x >> 5)) * kTlSlots + 63] = t;
  }
  tl_mark(0);
}
#define TL(slot) tl_mark(slot)
#define TLS(k) do { if (t >= 2 && t < 9) tl_mark(3 + 8 * (t - 2) + (k)); } while (0)   // (first partition of a pipeline)
#define TL_START() tl_start()
#else
#define TL(slot)
#define TLS(k)
#define TL_START()
#endif

__device__ __forceinline__ void st_async_f32x4(uint32_t cluster_addr, float4 v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

// kPipes in {1, 2}: independent partition pipelines per CTA, 16 / kPipes warps each (the launcher takes one pipeline while
// that still gives every chain partition its own, else two).
// kChained (a chain-partition table is given): the steps of a partition hand the state over through DISTRIBUTED SHARED
// MEMORY -- each CTA stores its 32 hidden columns of the new state (fp32) straight into the staging tile of all four CTAs
// with st.async (completion counted in bytes on the consumer's `ready` mbarrier): no release fence, no L2 round trip.
// ONE staging tile per pipeline: a tile step converts it into the tf32 hi / lo operand first thing (h0 is read back from
// the operand: hi + lo is the fp32 value exactly), then tells all four CTAs that its staging tile is vacant (`vacant`
// mbarrier, four arrivals per tile step); the gate phase waits for that before it sends the next state.  The operand keeps
// the PREVIOUS step's row order; a row reads its accumulator column / h0 values through prev_row, and the decay factor -- a
// per-row scalar -- is applied to the gathered values after the MMA (W . (c h) = c (W . h)).  The global state store only
// serves the callers.
// !kChained (one step, plain row tiles, state read from global memory through prev_row): the operand is gathered by the
// CTA itself; tiles are independent, nothing is exchanged.
template <int kPipes, bool kChained, int kRows>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    gru_scan_tm_kernel(const TempGruScanArgs P, const int n_parts, const void* __restrict__ w_packed) {
  static_assert(kRows == 48 || (kRows == 64 && kChained), "tile rows");
  constexpr int kAtomBytes = kRows * 128;           // one k-atom block (32 k) of an operand image
  constexpr int kImage = 4 * kAtomBytes;            // hi (or lo) image of a tile = one fp32 staging tile: kRows x 512 bytes
  constexpr int kBufBytes = 2 * kImage;             // one operand tile: hi image, lo image
  static_assert(kAtomBytes % 1024 == 0, "k-atom blocks must keep the 1024-byte swizzle period");
  static_assert(kChained ? (kBufBytes <= kStageOff && kStageOff + kImage <= kPipeSmem) : 2 * kBufBytes <= kPipeSmem, "pipeline memory");
  constexpr int kPW = kWarps / kPipes;              // warps per pipeline (16 or 8)
  constexpr int kPT = kPW * 32;
  constexpr int kH = kPW / 4;                       // warps per TMEM lane quadrant inside a pipeline
  constexpr int kRP = 4 * kPW;                      // gate phase: rows per pass (a thread = one row x 4 hidden columns)
  constexpr int kUg = (kRows + kRP - 1) / kRP;      // passes: 1 or 2
  constexpr int kU = kRows / kPW;                   // gather phase (!kChained): rows per warp, a lane = 4 of 128 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  __shared__ Bars S;

  const int tid = threadIdx.x, lane = tid & 31;
  // the broadcast makes the warp index (and everything derived from it: operand addresses, descriptors) provably
  // warp-uniform, so the MMA issue below runs on uniform registers instead of a per-operand R2UR election loop
  const int warp = __shfl_sync(kFull, tid >> 5, 0);
  const int pipe = warp / kPW, gw = warp % kPW;
  const int quad = warp & 3;                                        // TMEM lane quadrant this warp can access
  const int cb = blockIdx.x & (kCluster - 1), jb = 32 * cb;         // == %cluster_ctarank for a 1-D grid
  const int cid = blockIdx.x / kCluster, n_clusters = gridDim.x / kCluster;
  const int bar_id = 1 + pipe;                                      // named barrier of this pipeline
  const int rsub = lane >> 3, cq = lane & 7;                        // gate phase: row inside the warp's 4, column quad

  if (tid == 0) {
    mbar_init(&S.w_full, 1);
    for (int i = 0; i < kMaxPipes; ++i) {
      mbar_init(&S.mma_done[i], 1);
      mbar_init(&S.ready[i], 1);
      mbar_init(&S.vacant[i], kCluster);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&S.tmem_base, kTmemCols);
  if (P.push_bufs != nullptr && tid < P.push_world) S.push[tid] = P.push_bufs[tid] + P.push_offset;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(kFull, S.tmem_base, 0);
  pdl_launch_dependents();
  TL_START();
  cluster_sync_all();   // every CTA's barriers exist before a peer may signal them
  TL(1);

  // ---- this CTA's W_hh slice -> tensor memory (parameters: not produced by the predecessor kernels) -----------------
  // The slice (4 k-chunks x (hi image, lo image) x 16 KB, contiguous in temp_pack_gru_weights' layout) is fetched by the
  // copy engine into shared memory that the FIRST tile step of a chained pipeline does not touch -- a partition's first
  // step has no previous state, hence no operand tile and no MMA -- so the 128 KB L2 read overlaps that step instead of
  // standing in front of it; install_w() (all warps, once) then moves it to TMEM: warp (quad, kq = warp >> 2) writes
  // TMEM lanes 32 quad .. + 31 (feature rows m), k columns 32 kq .. + 31 of the hi and lo parts.
  // 16 KB piece g = 2 kq + img: pieces 0-2 -> pipeline 0's operand region, 3-5 -> pipeline 1's, 6-7 -> the extra block.
  auto w_piece = [&](int g) -> uint32_t {
    const uint32_t base = smem_u32(smem);
    return g < 3 ? base + g * 16384 : g < 6 ? base + kPipeSmem + (g - 3) * 16384 : base + kMaxPipes * kPipeSmem + (g - 6) * 16384;
  };
  if (w_packed != nullptr && tid == 0) {
    const uint8_t* slice = static_cast<const uint8_t*>(w_packed) + static_cast<size_t>(cb) * 4 * kWChunkBytes;
    mbar_expect_tx(&S.w_full, 4 * kWChunkBytes);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(w_piece(g)),
                   "l"(slice + g * 16384), "r"(16384), "r"(smem_u32(&S.w_full))
                   : "memory");
    }
  }
  auto install_w = [&]() {
    if (w_packed != nullptr) {
      const int m = 32 * quad + lane, kq = warp >> 2;
      mbar_wait(&S.w_full, 0);
#pragma unroll
      for (int img = 0; img < 2; ++img) {
        const uint32_t rowp = w_piece(2 * kq + img) + (m >> 3) * 1024 + (m & 7) * 128;
        float v[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = lds_f32x4(rowp + ((c ^ (m & 7)) << 4));
          v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        }
        tmem_st32(tbase + (static_cast<uint32_t>(32 * quad) << 16) + img * kColWLo + kq * 32, v);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();   // every warp's part is in TMEM; the shared-memory copy may now be overwritten by operand tiles
    tc_fence_after();
  };
  bool w_installed = false;
  if (!kChained) {     // plain row tiles read a previous state (and multiply) in their first tile step already
    install_w();
    w_installed = true;
  }

  // ---- this pipeline's tile steps: its partitions vc, vc + NV, ... ; all steps of a partition, then the next --------
  const int vc = pipe * n_clusters + cid, NV = kPipes * n_clusters;
  const uint32_t s_buf0 = smem_u32(smem + pipe * kPipeSmem);   // (plain) operand tile b: hi image at s_buf0 + b * kBufBytes, lo kImage later
  const uint32_t s_stage = s_buf0 + kStageOff;                 // (chained) fp32 [row][128] state of the previous step
  const uint32_t dcol = tbase + kColD + pipe * kColDStride;

  // lane l < n_steps keeps the packed-row range of a partition at step l: one load per partition, fetched one partition
  // ahead, so that walking the tile steps never waits for the table
  auto load_ranges = [&](int pt) -> int2 {
    int2 r = make_int2(0, 0);
    if (pt < n_parts && lane < P.n_steps) {
      if (kChained) {
        r = __ldg(reinterpret_cast<const int2*>(P.parts) + static_cast<size_t>(pt) * P.part_stride + P.steps[lane].part_col);
      } else {   // single step without a partition table: plain row tiles
        r.x = P.steps[0].row0 + pt * kRows;
        r.y = min(r.x + kRows, P.steps[0].row1);
      }
    }
    return r;
  };
  auto first_step = [&](int2 r, int from) -> int {   // first step >= from at which the partition has rows, -1: none
    unsigned m = __ballot_sync(kFull, r.y > r.x);
    m = from < 32 ? ((m >> from) << from) : 0u;
    return m != 0u ? __ffs(m) - 1 : -1;
  };
  // next tile step after (pt, st): further steps of the partition, then the pipeline's next partitions
  auto advance = [&](int& pt, int& st, int2& r, int2& r_next) {
    st = first_step(r, st + 1);
    while (st < 0 && pt < n_parts) {
      pt += NV;
      r = r_next;
      r_next = load_ranges(pt + NV);
      st = first_step(r, 0);
    }
  };
  // what a tile step needs from the plan, fetched one tile step ahead: per gate-phase row the previous-state row (-1:
  // none) and the time gap; the time-embedding rows of the first / last row; (!kChained) the gather phase's rows on lanes
  struct Pre {
    int prv[kUg];
    float dt[kUg];
    int tr0, tr1;
    int gprv;     // !kChained: lanes 0 .. kU-1 = previous-state row of this warp's u-th gather row
  };
  auto load_pre = [&](const TempGruArgs& q, int g0, int g1, Pre& f) {
#pragma unroll
    for (int u = 0; u < kUg; ++u) {
      const int r = g0 + 4 * gw + rsub + kRP * u;
      f.prv[u] = -1;
      f.dt[u] = 0.f;
      if (r < g1 && q.prev_row != nullptr) {
        f.prv[u] = __ldg(q.prev_row + r);
        if (q.dt != nullptr) f.dt[u] = __ldg(q.dt + r);
      }
    }
    f.tr0 = f.tr1 = q.row_time_scalar;
    if (q.time_embed != nullptr && q.row_time != nullptr) {
      f.tr0 = __ldg(q.row_time + g0);
      f.tr1 = __ldg(q.row_time + g1 - 1);
    }
    f.gprv = -1;
    if (!kChained && lane < kU) {
      const int r = g0 + gw + kPW * lane;
      if (r < g1 && q.prev_row != nullptr) f.gprv = __ldg(q.prev_row + r);
    }
  };

  int part = vc, s = -1;
  int2 rgs = load_ranges(part), rgs_next = load_ranges(part + NV);
  advance(part, s, rgs, rgs_next);
  bool have = part < n_parts;
  Pre cur;
  if (have) load_pre(P.steps[s], __shfl_sync(kFull, rgs.x, s), __shfl_sync(kFull, rgs.y, s), cur);
  TL(2);
  pdl_wait();   // gi (and, for a single step, the previous state) come from the predecessor kernels

  uint32_t t = 0;                    // tile steps done by this pipeline (operand tile of step t: t & 1)
  uint32_t mma_par = 0, ready_par = 0, vacant_par = 0;   // phase parities of S.mma_done / S.ready / S.vacant of this pipeline
  int chain_part = -1, prev_rb = 0, prev_rows = 0;   // the previous tile step (same partition <=> its state is the operand)
#pragma unroll 1
  while (have) {
    const TempGruArgs& p = P.steps[s];
    const int rb = __shfl_sync(kFull, rgs.x, s), r1 = __shfl_sync(kFull, rgs.y, s);
    int n_part = part, n_s = s;
    int2 n_rgs = rgs, n_rgs_next = rgs_next;
    advance(n_part, n_s, n_rgs, n_rgs_next);
    const bool more = n_part < n_parts;
    const bool cont_next = kChained && more && n_part == part;   // the next tile step consumes this one's state
    const uint32_t b = t & 1;
    // chained: ONE operand tile (and one staging tile); plain: two operand tiles alternate
    const uint32_t s_bhi = kChained ? s_buf0 : s_buf0 + b * kBufBytes, s_blo = s_bhi + kImage;
    const uint32_t ex_base = s_bhi;   // gate g parks in k-atom (cb + 1 + g) & 3 of the (dead) hi image, plain [row][32] rows
    bool has_prev;
    int nmma;
    TLS(0);

    // ---- everything that does not depend on the previous step is requested first: input gates, biases, time embedding --
    const int j4 = jb + 4 * cq;
    float4 gr[kUg], gz[kUg], gn[kUg];
#pragma unroll
    for (int u = 0; u < kUg; ++u) {
      const int r = rb + 4 * gw + rsub + kRP * u;
      gr[u] = gz[u] = gn[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < r1) {
        const float4* gi = reinterpret_cast<const float4*>(p.gi + static_cast<size_t>(r) * p.gi_ld + p.gi_off + j4);
        gr[u] = ld_dep_f32x4(gi);              // (written by the predecessor kernel: see ld_dep_* in tc_common.cuh)
        gz[u] = ld_dep_f32x4(gi + kD / 4);
        gn[u] = ld_dep_f32x4(gi + 2 * (kD / 4));
      }
    }
    const float4 br = __ldg(reinterpret_cast<const float4*>(p.b_hh + j4));
    const float4 bz = __ldg(reinterpret_cast<const float4*>(p.b_hh + kD + j4));
    const float4 bn = __ldg(reinterpret_cast<const float4*>(p.b_hh + 2 * kD + j4));
    // rows of one partition step belong to one snapshot instance (one time-embedding row); plain tiles may mix
    float4 te_uni = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool te_rows = p.time_embed != nullptr && cur.tr0 != cur.tr1;
    if (p.time_embed != nullptr && !te_rows)
      te_uni = __ldg(reinterpret_cast<const float4*>(p.time_embed + static_cast<size_t>(cur.tr0) * kD + j4));

    // ---- the previous state -> shared-memory operand (tf32 hi / lo parts; the decay is applied after the MMA), and
    // gh^T = W_hh . h0^T : 16 k-steps x 3 split passes, A from tensor memory -------------------------------------------
    auto issue_atoms = [&](int ka0, int ka1, uint32_t s_hi, uint32_t s_lo) {   // (one elected lane) k-atoms [ka0, ka1)
      const uint32_t idesc = umma_idesc_tf32(128, nmma);
#pragma unroll
      for (int ka = ka0; ka < ka1; ++ka) {
        const uint32_t bh0 = umma_desc_lo(s_hi + ka * kAtomBytes), bl0 = umma_desc_lo(s_lo + ka * kAtomBytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a_hi = tbase + ka * 32 + ks * 8, a_lo = a_hi + kColWLo;
          const uint32_t bh = bh0 + 2 * ks, bl = bl0 + 2 * ks;          // +32 bytes per k-step inside the atom
          umma_tf32_ts(dcol, a_lo, bh, idesc, (ka | ks) != 0 ? 1u : 0u);  // small terms first
          umma_tf32_ts(dcol, a_hi, bl, idesc, 1u);
          umma_tf32_ts(dcol, a_hi, bh, idesc, 1u);
        }
      }
    };
    if (kChained) {
      has_prev = chain_part == part;
      nmma = (prev_rows + 15) & ~15;
      if (has_prev) {
        mbar_wait(&S.ready[pipe], ready_par);   // the staging tile is complete (all 4 CTAs' columns)
        ready_par ^= 1u;
        TLS(1);
        // Two stages of two k-atoms each: while the issuing warp feeds the tensor pipe with the first half (the issue loop
        // is throttled to the pipe's rate, ~35 cycles per MMA), the other warps convert the second half.  A load covers two
        // rows x 256 bytes (lane >> 4 = row of the pair, lane & 15 = 16-byte chunk = 2 atoms x 8 chunks).
        const int rsel = lane >> 4, asel = (lane & 15) >> 3, cch = lane & 7;
        auto convert = [&](int stage, int first, int stride, int count) {
#pragma unroll
          for (int j = 0; j < count; ++j) {
            const int q = first + stride * j;          // row pair
            if (2 * q >= nmma) continue;               // warp-uniform: beyond the rows this step's MMA reads
            const int i = 2 * q + rsel, ka = 2 * stage + asel;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < prev_rows) v = lds_f32x4(s_stage + i * 512 + ka * 128 + cch * 16);
            float4 hi, lo;
            split_tf32(v.x, hi.x, lo.x);
            split_tf32(v.y, hi.y, lo.y);
            split_tf32(v.z, hi.z, lo.z);
            split_tf32(v.w, hi.w, lo.w);
            const uint32_t off = static_cast<uint32_t>(ka) * kAtomBytes + (i >> 3) * 1024u + (i & 7) * 128u + ((cch ^ (i & 7)) << 4);
            sts_f32x4(s_bhi + off, hi);
            sts_f32x4(s_blo + off, lo);
          }
        };
        constexpr int kPairs = kRows / 2;
        convert(0, gw, kPW, (kPairs + kPW - 1) / kPW);                                   // stage 0: all warps
        fence_proxy_async();
        bar_named(bar_id, kPT);
        if (gw == 0) {   // warp-uniform: the issuing warp
          tc_fence_after();
          if (elect_one()) issue_atoms(0, 2, s_bhi, s_blo);
          __syncwarp();
        } else {
          convert(1, gw - 1, kPW - 1, (kPairs + kPW - 2) / (kPW - 1));                    // stage 1: everyone else
          fence_proxy_async();
        }
        bar_named(bar_id, kPT);
      }
    } else {
      float4 v[kU];
      nmma = (r1 - rb + 15) & ~15;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int pr = __shfl_sync(kFull, cur.gprv, u);
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pr >= 0) v[u] = ld_dep_f32x4(reinterpret_cast<const float4*>(p.state + static_cast<size_t>(pr) * kD) + lane);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = gw + kPW * u;
        if (i >= nmma) continue;   // warp-uniform: beyond the rows this step's MMA reads
        float4 hi, lo;
        split_tf32(v[u].x, hi.x, lo.x);
        split_tf32(v[u].y, hi.y, lo.y);
        split_tf32(v[u].z, hi.z, lo.z);
        split_tf32(v[u].w, hi.w, lo.w);
        const uint32_t off = static_cast<uint32_t>(lane >> 3) * kAtomBytes + (i >> 3) * 1024u + (i & 7) * 128u +
                             (((lane & 7) ^ (i & 7)) << 4);
        sts_f32x4(s_bhi + off, hi);
        sts_f32x4(s_blo + off, lo);
      }
      fence_proxy_async();
      has_prev = bar_red_or(bar_id, kPT, __any_sync(kFull, lane < kU && cur.gprv >= 0));
    }
    TLS(2);

    // ---- the staging tile has been consumed (or was never read: a partition's first step): arm `ready` for the bytes this
    // step's rows will bring, then tell all four CTAs that this CTA's staging tile is vacant ----------------------------
    if (kChained && gw == 0) {   // warp-uniform
      if (elect_one()) {
        if (cont_next) mbar_expect_tx(&S.ready[pipe], static_cast<uint32_t>(r1 - rb) * 512u);
        const uint32_t vb = smem_u32(&S.vacant[pipe]);
#pragma unroll
        for (int k = 0; k < kCluster; ++k) mbar_arrive_remote_relaxed(mapa_u32(vb, k));
      }
      __syncwarp();
    }

    // ---- the (rest of the) MMA batch ------------------------------------------------------------------------------
    if (has_prev && gw == 0) {   // warp-uniform
      tc_fence_after();
      if (elect_one()) {
        if (kChained)
          issue_atoms(2, 4, s_bhi, s_blo);
        else
          issue_atoms(0, 4, s_bhi, s_blo);
        umma_commit(&S.mma_done[pipe]);
      }
      __syncwarp();
    }
    TLS(3);

    // ---- while the MMA runs: h0 of this thread's (row, 4 columns); the NEXT tile step's indices ------------------------
    float4 h0[kUg];
    int nrow[kUg];   // operand / accumulator row of this thread's row: prev_row relative to the previous step (-1: no state)
#pragma unroll
    for (int u = 0; u < kUg; ++u) {
      const int i = 4 * gw + rsub + kRP * u;
      h0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      nrow[u] = -1;
      // (a partition's first tile step has no state to read: the planner links a row only to the previous step's rows)
      if (kChained && !has_prev && rb + i < r1 && cur.prv[u] >= 0) __trap();
      if (rb + i < r1 && has_prev && cur.prv[u] >= 0) {
        const int n = kChained ? cur.prv[u] - prev_rb : i;
        if (kChained && (n < 0 || n >= prev_rows)) __trap();   // a state row outside the previous step of this partition
        nrow[u] = n;
        {   // from the operand images: hi + lo is the fp32 value exactly (k-atom cb = this CTA's own hidden columns)
          const uint32_t off = static_cast<uint32_t>(cb) * kAtomBytes + (n >> 3) * 1024u + (n & 7) * 128u + ((cq ^ (n & 7)) << 4);
          const float4 a = lds_f32x4(s_bhi + off), c = lds_f32x4(s_blo + off);
          h0[u] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
        }
      }
    }
    Pre nxt;
    if (more) load_pre(P.steps[n_s], __shfl_sync(kFull, n_rgs.x, n_s), __shfl_sync(kFull, n_rgs.y, n_s), nxt);

    // ---- accumulator (gate = lane quadrant, rows on columns) -> exchange rows, one barrier --------------------------
    if (has_prev) {
      mbar_wait(&S.mma_done[pipe], mma_par);
      mma_par ^= 1;
      tc_fence_after();
      TLS(4);
      if (quad < 3) {
        const int h = gw >> 2;   // the kH warps of a quadrant take the 16-column chunks round robin
        const uint32_t ta = dcol + (static_cast<uint32_t>(32 * quad) << 16);
        const uint32_t exw = ex_base + ((cb + 1 + quad) & 3) * kAtomBytes + lane * 4;
#pragma unroll
        for (int c = 0; c < kRows / 16; ++c) {
          if (c % kH == h && 16 * c < nmma) {   // warp-uniform
            float v[16];
            tmem_ld16(ta + 16 * c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) sts_f32(exw + (16 * c + i) * 128, v[i]);
          }
        }
      }
      tc_fence_before();
      bar_named(bar_id, kPT);
    }
    TLS(5);

    // ---- gates, state store, hand-over to the next step (+ fused all-gather peer stores) ------------------------------
    uint32_t peer_stage[kCluster], peer_bar[kCluster];
    if (kChained) {   // every CTA of the cluster has converted its staging tile of this tile step: it may be overwritten
      mbar_wait(&S.vacant[pipe], vacant_par);
      vacant_par ^= 1u;
    }
    if (cont_next) {
#pragma unroll
      for (int k = 0; k < kCluster; ++k) {
        peer_stage[k] = mapa_u32(s_stage, k);
        peer_bar[k] = mapa_u32(smem_u32(&S.ready[pipe]), k);
      }
    }
#pragma unroll
    for (int u = 0; u < kUg; ++u) {
      const int i = 4 * gw + rsub + kRP * u;
      const int r = rb + i;
      if (r < r1) {
        float4 hr = br, hz = bz, hn = bn;
        float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nrow[u] >= 0) {
          const float dec = p.dt != nullptr ? decay_factor(cur.dt[u], p.decay_wb, p.inv_temperature) : 1.f;
          const uint32_t ea = ex_base + nrow[u] * 128 + cq * 16;
          const float4 xr = lds_f32x4(ea + ((cb + 1) & 3) * kAtomBytes);
          const float4 xz = lds_f32x4(ea + ((cb + 2) & 3) * kAtomBytes);
          const float4 xn = lds_f32x4(ea + ((cb + 3) & 3) * kAtomBytes);
          hr.x = fmaf(dec, xr.x, hr.x); hr.y = fmaf(dec, xr.y, hr.y); hr.z = fmaf(dec, xr.z, hr.z); hr.w = fmaf(dec, xr.w, hr.w);
          hz.x = fmaf(dec, xz.x, hz.x); hz.y = fmaf(dec, xz.y, hz.y); hz.z = fmaf(dec, xz.z, hz.z); hz.w = fmaf(dec, xz.w, hz.w);
          hn.x = fmaf(dec, xn.x, hn.x); hn.y = fmaf(dec, xn.y, hn.y); hn.z = fmaf(dec, xn.z, hn.z); hn.w = fmaf(dec, xn.w, hn.w);
          hp = make_float4(dec * h0[u].x, dec * h0[u].y, dec * h0[u].z, dec * h0[u].w);
        }
        // torch.nn.GRU, gate order r, z, n
        float4 hy;
        {
          const float rg_ = fast_sigmoid(gr[u].x + hr.x), zg = fast_sigmoid(gz[u].x + hz.x);
          const float ng = fast_tanh(gn[u].x + rg_ * hn.x);
          hy.x = (1.f - zg) * ng + zg * hp.x;
        }
        {
          const float rg_ = fast_sigmoid(gr[u].y + hr.y), zg = fast_sigmoid(gz[u].y + hz.y);
          const float ng = fast_tanh(gn[u].y + rg_ * hn.y);
          hy.y = (1.f - zg) * ng + zg * hp.y;
        }
        {
          const float rg_ = fast_sigmoid(gr[u].z + hr.z), zg = fast_sigmoid(gz[u].z + hz.z);
          const float ng = fast_tanh(gn[u].z + rg_ * hn.z);
          hy.z = (1.f - zg) * ng + zg * hp.z;
        }
        {
          const float rg_ = fast_sigmoid(gr[u].w + hr.w), zg = fast_sigmoid(gz[u].w + hz.w);
          const float ng = fast_tanh(gn[u].w + rg_ * hn.w);
          hy.w = (1.f - zg) * ng + zg * hp.w;
        }
        float4 te = te_uni;
        if (te_rows) te = __ldg(reinterpret_cast<const float4*>(p.time_embed + static_cast<size_t>(__ldg(p.row_time + r)) * kD + j4));
        hy.x += te.x; hy.y += te.y; hy.z += te.z; hy.w += te.w;
        float4* o = reinterpret_cast<float4*>(p.out + static_cast<size_t>(r) * kD + j4);
        if (p.accumulate) {
          const float4 a = __ldcg(o);
          hy.x += a.x; hy.y += a.y; hy.z += a.z; hy.w += a.w;
        }
        if (cont_next) {   // this CTA's 32 columns of the next step's staging tile, in all four CTAs (row order of THIS step)
          const uint32_t off = static_cast<uint32_t>(i) * 512u + cb * 128u + cq * 16u;
#pragma unroll
          for (int k = 0; k < kCluster; ++k) st_async_f32x4(peer_stage[k] + off, hy, peer_bar[k]);
        }
        *o = hy;
        if (p.push != 0 && (P.push_bufs != nullptr || P.push_multicast != nullptr)) {   // NVLink stores into every peer's slab
          const size_t po = static_cast<size_t>(r - P.push_row0) * kD + j4;
          if (P.push_multicast != nullptr) {   // one 16-byte store, replicated to every GPU by the NVSwitch (NVLS)
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.push_multicast + P.push_offset + po),
                         "f"(hy.x), "f"(hy.y), "f"(hy.z), "f"(hy.w)
                         : "memory");
          }
          // (with a multicast address push_bufs lists only the buffers OUTSIDE the multicast group, e.g. a host buffer)
          for (int k = 0; k < P.push_world; ++k) *reinterpret_cast<float4*>(S.push[k] + po) = hy;
        }
      }
    }
    TLS(6);
    chain_part = part;
    prev_rb = rb;
    prev_rows = r1 - rb;
    t += 1;
    cur = nxt;
    part = n_part;
    s = n_s;
    rgs = n_rgs;
    rgs_next = n_rgs_next;
    have = more;
    if (!w_installed) {   // (pipeline-uniform; both pipelines meet here after their first tile step)
      install_w();
      w_installed = true;
    }
  }
  if (!w_installed) install_w();   // a pipeline without any tile step still takes part

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while a peer may still signal its barriers / write its operand tiles
  if (warp == 0) tmem_dealloc(tbase, kTmemCols);
}

}  // namespace

#ifdef TEMP_TIMELINE
extern "C" int temp_debug_timeline2(void* device_buffer) {  // [ctas][16 warps][64 slots] u64, or null to disable
  unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
  return cudaMemcpyToSymbol(g_timeline2, &p, sizeof(p)) == cudaSuccess ? 0 : -2;
}
#endif

namespace temp_internal {

// One recurrent cell (torch GRU equations) for every step that reads a previous state, d == 128, chain partitions of at
// most 48 rows per step (TempGruScanArgs.part_rows) or a single step of plain row tiles.
bool tc_scan2_supported(const TempGruScanArgs* a) {
  static const bool disabled = getenv("TEMP_SCAN_V1") != nullptr;
  if (disabled || a->n_steps <= 0) return false;
  if (a->parts == nullptr && a->n_steps != 1) return false;
  if (a->parts != nullptr && (a->part_rows <= 0 || a->part_rows > kRowsMax)) return false;
  const void* w = nullptr;
  for (int s = 0; s < a->n_steps; ++s) {
    const TempGruArgs& g = a->steps[s];
    if (g.d != kD || g.cell_type != TEMP_CELL_TORCH_GRU) return false;
    if (a->parts != nullptr && (g.part_col < 0 || g.part_col >= a->part_stride)) return false;
    // chained steps hand the state over on chip: what a step reads through prev_row must be what its predecessor wrote
    // (an accumulating step -- the Bi centre step of the backward cell -- adds to what an EARLIER launch wrote)
    if (a->parts != nullptr && ((g.prev_row != nullptr && g.state != a->steps[0].out) || g.out != a->steps[0].out)) return false;
    if (g.prev_row != nullptr) {
      if (g.whh_packed == nullptr) return false;
      if (w != nullptr && g.whh_packed != w) return false;
      w = g.whh_packed;
    }
  }
  if (a->push_bufs != nullptr && (a->push_world < 0 || a->push_world > TEMP_MAX_PUSH_PEERS)) return false;
  if (a->push_bufs != nullptr && a->push_world == 0 && a->push_multicast == nullptr) return false;
  return true;
}

template <int kPipes, bool kChained, int kRows>
int launch_scan2_t(const TempGruScanArgs* a, int n_parts, int clusters, const void* w, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gru_scan_tm_kernel<kPipes, kChained, kRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return cuda_fail(e, "gru_scan_tm_kernel");
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cfg.gridDim = dim3(kCluster * clusters);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gru_scan_tm_kernel<kPipes, kChained, kRows>, *a, n_parts, w);
  if (e != cudaSuccess) return cuda_fail(e, "gru_scan_tm_kernel launch");
  return TEMP_OK;
}

int tc_launch_scan2(const TempGruScanArgs* a, cudaStream_t st) {
  int n_parts = a->n_parts;
  if (a->parts == nullptr) n_parts = (a->steps[0].row1 - a->steps[0].row0 + kRowsPlain - 1) / kRowsPlain;
  if (n_parts <= 0) return TEMP_OK;
  const void* w = nullptr;
  for (int s = 0; s < a->n_steps; ++s)
    if (a->steps[s].prev_row != nullptr) w = a->steps[s].whh_packed;
  static int max_clusters = 0;   // co-resident 4-CTA clusters (one CTA per SM)
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.gridDim = dim3(kCluster * 64);
    cudaFuncSetAttribute(gru_scan_tm_kernel<2, true, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, gru_scan_tm_kernel<2, true, 48>, &cfg);
    if (e != cudaSuccess || max_clusters <= 0) {
      cudaGetLastError();
      max_clusters = 32;
    }
  }
  static const int force = getenv("TEMP_SCAN_PIPES") != nullptr ? atoi(getenv("TEMP_SCAN_PIPES")) : 0;   // development knob
  // spread first: as many clusters as there are partitions; one pipeline per CTA while that gives every partition its
  // own (16 warps per tile = the shortest step chain), two when partitions queue anyway
  const int clusters = n_parts < max_clusters ? n_parts : max_clusters;
  int pipes = n_parts <= clusters ? 1 : 2;
  if (force == 1 || force == 2) pipes = force;
  if (a->parts != nullptr) {
    if (a->part_rows <= 48)
      return pipes == 1 ? launch_scan2_t<1, true, 48>(a, n_parts, clusters, w, st) : launch_scan2_t<2, true, 48>(a, n_parts, clusters, w, st);
    return pipes == 1 ? launch_scan2_t<1, true, 64>(a, n_parts, clusters, w, st) : launch_scan2_t<2, true, 64>(a, n_parts, clusters, w, st);
  }
  return pipes == 1 ? launch_scan2_t<1, false, 48>(a, n_parts, clusters, w, st) : launch_scan2_t<2, false, 48>(a, n_parts, clusters, w, st);
}

}  // namespace temp_internal
