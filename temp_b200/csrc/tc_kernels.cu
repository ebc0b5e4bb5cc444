// temp_b200 -- tcgen05 (5th-gen tensor core) kernels for embed_size == hidden_size == 128, the width of every
// shipped TeMP configuration (grid/*/config_*.json).  fp32 parity is kept with a 3xTF32 operand split
// (tc_common.cuh); everything that is not a dense contraction (the CSR gather with the in-register 1x1
// relation projection, norms, bias / activation / time embedding, GRU gates) stays fp32 SIMT.
//
// "Features on TMEM lanes": every GEMM computes D^T[feature, row] so that a warp's 32 lanes hold 32
// consecutive features of one packed row -- global activations are read and written as coalesced 128-byte
// lines and the thread that gathered agg[row][feature] is the thread that later reads D^T[feature][row].
//
//   rgcn_gather_kernel   : the CSR-by-destination aggregation with the in-register 1x1 relation projection as its own
//        HBM / L2-bound launch (compact work lists, a warp per row, a thread block per high in-degree row).
//   rgcn_layer_tc_kernel : one CTA per 128 packed rows.
//        workers (8 warps) : self-loop operand rows -> smem (hi/lo, SWIZZLE_128B) ; aggregate rows prefetched ;
//                            epilogue 1 (agg + D1 + bias, act, time embedding, h_out) ; chain operand X written
//                            in place over the self-loop operand ; chain epilogues (gi / q|k|v + bias)
//        TMA warp          : packed weight chunks (32 KB = 128 features x 32 k, hi + lo) through a 3-stage ring
//        MMA warp          : D1 = W_loop^T . x^T, then per 128 chain features D2 = W_chain . X^T (TMEM ring of 3)
//   gru_scan_tc_kernel   : persistent chain-partitioned scan over the GRU steps of a window; CTA = (block of 32 hidden
//        columns) of a 4-CTA cluster, tiles of <= 96 rows per partition step; its W_hh slice (r|z|n rows, hi/lo) stays
//        in shared memory for the whole scan; steps are separated by the hardware cluster barrier.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <utility>

#include <type_traits>

#include "internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr unsigned kFull = 0xffffffffu;
constexpr int kD = 128;
constexpr int kKAtoms = kD / kAtomK;  // 4

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float decay_factor(float dt, const float* wb, float inv_temperature) {
  if (wb != nullptr) return expf(-fmaxf(fmaf(__ldg(wb), dt, __ldg(wb + 1)), 0.f));
  return expf(-dt * inv_temperature);
}

// Development-only phase timeline (tools/probe_timeline.py builds a separate library with -DTEMP_TIMELINE):
// lane 0 of every warp stores clock64() per phase slot; slot 63 holds %globaltimer at kernel entry.
#ifdef TEMP_TIMELINE
__device__ unsigned long long* g_timeline = nullptr;
constexpr int kTlWarps = 20, kTlSlots = 64;
__device__ __forceinline__ void tl_mark(int slot) {
  if (g_timeline != nullptr && (threadIdx.x & 31) == 0)
    g_timeline[(static_cast<size_t>(blockIdx.x) * kTlWarps + (threadIdx.x >> 5)) * kTlSlots + slot] = clock64();
}
__device__ __forceinline__ void tl_start() {
  if (g_timeline != nullptr && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_timeline[(static_cast<size_t>(blockIdx.x) * kTlWarps + (threadIdx.x >> 5)) * kTlSlots + 63] = t;
  }
  tl_mark(0);
}
#define TL(slot) tl_mark(slot)
#define TL_START() tl_start()
#else
#define TL(slot)
#define TL_START()
#endif

// ------------------------------------------------------------------------------------------------
// weight packing (once per parameter version)
// ------------------------------------------------------------------------------------------------
// w [128, n] row-major (k x n).  chunk (mb, ka) at (mb * 4 + ka) * 32 KB: element (feature m, k) of the
// 128-feature block mb, k-atom ka; hi image then lo image.
__global__ void pack_weights_kernel(const float* __restrict__ w, int n, uint8_t* __restrict__ out) {
  const int total = kD * n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx / n, col = idx - k * n;
    const int mb = col >> 7, m = col & 127, ka = k >> 5, kk = k & 31;
    float hi, lo;
    split_tf32(w[idx], hi, lo);
    uint8_t* chunk = out + static_cast<size_t>(mb * kKAtoms + ka) * kWChunkBytes;
    const uint32_t off = sw128_off(m, kk);
    *reinterpret_cast<float*>(chunk + off) = hi;
    *reinterpret_cast<float*>(chunk + 128 * 128 + off) = lo;
  }
}

// whh_t [128, 384] row-major.  Block cb (32 hidden columns): 4 chunks of 32 KB; feature row m = 32 * gate + jj
// <-> column gate * 128 + 32 * cb + jj; rows 96..127 are zero.
__global__ void pack_gru_kernel(const float* __restrict__ whh_t, uint8_t* __restrict__ out) {
  const int total = 4 * kD * 128;  // cb, k, m
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cb = idx / (kD * 128), rem = idx - cb * (kD * 128);
    const int k = rem >> 7, m = rem & 127;
    const int gate = m >> 5, jj = m & 31;
    float v = 0.f;
    if (gate < 3) v = whh_t[static_cast<size_t>(k) * (3 * kD) + gate * kD + 32 * cb + jj];
    float hi, lo;
    split_tf32(v, hi, lo);
    uint8_t* chunk = out + static_cast<size_t>(cb * kKAtoms + (k >> 5)) * kWChunkBytes;
    const uint32_t off = sw128_off(m, k & 31);
    *reinterpret_cast<float*>(chunk + off) = hi;
    *reinterpret_cast<float*>(chunk + 128 * 128 + off) = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// RGCN aggregation (1x1 relation blocks, d == 128): the CSR-by-destination gather
// ------------------------------------------------------------------------------------------------
// agg[v] = norm_v * sum_{e: dst_e = v} norm_v * W[rel_e] (.) x[src_e]      (RGCN.py:91-104, diagonal blocks)
// One warp per destination row, a lane per 4 channels (one 512-byte feature row per load instruction), edge
// indices fetched lane-parallel, 2 feature-row pairs in flight per warp (40 warps per SM), edges summed in edge-id order (deterministic, no
// atomics).  Rows without in-edges are not written (the consumer tests row_ptr).  No shared memory, 40 resident
// warps per SM: this is the HBM / L2-bound part of the layer, kept out of the one-CTA-per-SM tensor-core kernel.
constexpr int kGatherWarps = 8;

// sum_{e in [e0, e1)} (x[src_e] * W[rel_e]) * nrm for this lane's 4 channels, in edge order
// (the first chunk's indices are plan data: they are fetched BEFORE pdl_wait(), the feature rows x after it)
__device__ __forceinline__ float4 gather_edges(const TempRgcnLayerArgs& p, int e0, int e1, float nrm, int lane) {
  const float4* x4 = reinterpret_cast<const float4*>(p.x) + lane;
  const float4* w4 = reinterpret_cast<const float4*>(p.weight) + lane;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int s0 = 0, rl0 = 0;
  if (e0 + lane < e1) {
    s0 = __ldg(p.e_src + e0 + lane);
    rl0 = __ldg(p.e_rel + e0 + lane);
  }
  pdl_wait();
  for (int base = e0; base < e1; base += 32) {
    const int cnt = min(32, e1 - base);
    int s = s0, rl = rl0;
    if (base != e0 && lane < cnt) {
      s = __ldg(p.e_src + base + lane);
      rl = __ldg(p.e_rel + base + lane);
    }
#pragma unroll 1
    for (int u0 = 0; u0 < cnt; u0 += 2) {
      float4 hv[2], wv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int su = __shfl_sync(kFull, s, (u0 + u) & 31), ru = __shfl_sync(kFull, rl, (u0 + u) & 31);
        if (u0 + u < cnt) {
          hv[u] = ld_dep_f32x4(x4 + static_cast<size_t>(su) * (kD / 4));   // (the previous layer's output)
          wv[u] = __ldg(w4 + static_cast<size_t>(ru) * (kD / 4));
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u0 + u < cnt) {  // msg = (h * w) * norm_e, summed in edge order (RGCN.py:92-97)
          a.x += (hv[u].x * wv[u].x) * nrm;
          a.y += (hv[u].y * wv[u].y) * nrm;
          a.z += (hv[u].z * wv[u].z) * nrm;
          a.w += (hv[u].w * wv[u].w) * nrm;
        }
      }
    }
  }
  return a;
}

__global__ void __launch_bounds__(kGatherWarps * 32, 5) rgcn_gather_kernel(const TempRgcnLayerArgs p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  if (p.agg_lists != 0 && static_cast<int>(blockIdx.x) < p.n_agg_heavy) {
    // ---- a high in-degree row: the block's 8 warps sum contiguous edge chunks, partials added in chunk order ----
    __shared__ float4 part[kGatherWarps][32];
    const int t = lane < 3 ? __ldg(p.agg_heavy + 3 * static_cast<size_t>(blockIdx.x) + lane) : 0;
    const int r = __shfl_sync(kFull, t, 0), p0 = __shfl_sync(kFull, t, 1), p1 = __shfl_sync(kFull, t, 2);
    const float nrm = __ldg(p.norm + r);
    const int chunk = (p1 - p0 + kGatherWarps - 1) / kGatherWarps;
    const int e0 = min(p0 + warp * chunk, p1), e1 = min(e0 + chunk, p1);
    part[warp][lane] = gather_edges(p, e0, e1, nrm, lane);
    __syncthreads();
    if (warp == 0) {
      float4 a = part[0][lane];
#pragma unroll
      for (int w = 1; w < kGatherWarps; ++w) {
        const float4 b = part[w][lane];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      a.x *= nrm; a.y *= nrm; a.z *= nrm; a.w *= nrm;  // apply_func (RGCN.py:103-104)
      reinterpret_cast<float4*>(p.agg_scratch + static_cast<size_t>(r) * kD)[lane] = a;
    }
    return;
  }
  int r, p0, p1;
  if (p.agg_lists != 0) {  // compact work list: rows with in-edges only, CSR range inline
    const int item = (static_cast<int>(blockIdx.x) - p.n_agg_heavy) * kGatherWarps + warp;
    if (item >= p.n_agg_rows) return;
    const int t = lane < 3 ? __ldg(p.agg_rows + 3 * static_cast<size_t>(item) + lane) : 0;
    r = __shfl_sync(kFull, t, 0);
    p0 = __shfl_sync(kFull, t, 1);
    p1 = __shfl_sync(kFull, t, 2);
  } else {
    r = p.row0 + blockIdx.x * kGatherWarps + warp;
    if (r >= p.row1) return;
    p0 = __ldg(p.row_ptr + r);
    p1 = __ldg(p.row_ptr + r + 1);
    if (p1 <= p0) return;
  }
  const float nrm = __ldg(p.norm + r);
  float4 a = gather_edges(p, p0, p1, nrm, lane);
  a.x *= nrm; a.y *= nrm; a.z *= nrm; a.w *= nrm;  // apply_func (RGCN.py:103-104)
  reinterpret_cast<float4*>(p.agg_scratch + static_cast<size_t>(r) * kD)[lane] = a;
}

// The same aggregation with the neighbour rows STREAMED INTO SHARED MEMORY BY THE TMA ENGINE (cp.async.bulk, one 512-byte
// row per copy, completion counted in bytes on an mbarrier) -- the north star's formulation.  A warp takes four light
// destination rows at a time; their (<= 32) edges are spread over the lanes, every lane issues the bulk copy of its own
// source row into the warp's 16-slot ring, so up to 16 rows (8 KB) per warp are in flight without holding registers;
// the reduction then walks the slots in edge order: the arithmetic (and therefore every bit of the result) is that of
// rgcn_gather_kernel.  High in-degree rows keep the thread-block path.  Which of the two runs is a measured choice
// (DESIGN.md section 8): TEMP_GATHER=ldg|bulk overrides it.
constexpr int kBulkSlots = 16;                  // 512-byte row slots per warp
constexpr int kBulkRows = 4;                    // destination rows per warp batch
constexpr int kBulkSmem = kGatherWarps * kBulkSlots * 512;   // 64 KB per CTA -> 3 CTAs (24 warps) per SM

__global__ void __launch_bounds__(kGatherWarps * 32, 3) rgcn_gather_bulk_kernel(const TempRgcnLayerArgs p) {
  extern __shared__ __align__(128) uint8_t ring_raw[];
  __shared__ uint64_t bars[kGatherWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  if (static_cast<int>(blockIdx.x) < p.n_agg_heavy) {
    // ---- a high in-degree row: the block's 8 warps sum contiguous edge chunks, partials added in chunk order ----
    __shared__ float4 part[kGatherWarps][32];
    const int t = lane < 3 ? __ldg(p.agg_heavy + 3 * static_cast<size_t>(blockIdx.x) + lane) : 0;
    const int r = __shfl_sync(kFull, t, 0), p0 = __shfl_sync(kFull, t, 1), p1 = __shfl_sync(kFull, t, 2);
    const float nrm = __ldg(p.norm + r);
    const int chunk = (p1 - p0 + kGatherWarps - 1) / kGatherWarps;
    const int e0 = min(p0 + warp * chunk, p1), e1 = min(e0 + chunk, p1);
    part[warp][lane] = gather_edges(p, e0, e1, nrm, lane);
    __syncthreads();
    if (warp == 0) {
      float4 a = part[0][lane];
#pragma unroll
      for (int w = 1; w < kGatherWarps; ++w) {
        const float4 b = part[w][lane];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      a.x *= nrm; a.y *= nrm; a.z *= nrm; a.w *= nrm;
      reinterpret_cast<float4*>(p.agg_scratch + static_cast<size_t>(r) * kD)[lane] = a;
    }
    return;
  }
  if (lane == 0) {
    mbar_init(&bars[warp], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int batch = (static_cast<int>(blockIdx.x) - p.n_agg_heavy) * kGatherWarps + warp;
  const int item0 = batch * kBulkRows;
  if (item0 >= p.n_agg_rows) return;
  const int n_rows = min(kBulkRows, p.n_agg_rows - item0);
  // lane k < n_rows: work-list entry k of the batch (row, first edge, end edge) and the row's norm
  int row = 0, e0 = 0, deg = 0;
  float nrm = 0.f;
  if (lane < n_rows) {
    const int* it = p.agg_rows + 3 * static_cast<size_t>(item0 + lane);
    row = __ldg(it);
    e0 = __ldg(it + 1);
    deg = __ldg(it + 2) - e0;
    nrm = __ldg(p.norm + row);
  }
  int off[kBulkRows + 1];
  off[0] = 0;
#pragma unroll
  for (int k = 0; k < kBulkRows; ++k) off[k + 1] = off[k] + __shfl_sync(kFull, deg, k);
  const int total = off[kBulkRows];            // <= 4 * 8 = 32 edges: one per lane
  // lane q < total: edge q of the batch -> (row slot, source row, relation)
  int k_of = 0, src = 0, rel = 0;
  if (lane < total) {
#pragma unroll
    for (int k = 1; k < kBulkRows; ++k) k_of += lane >= off[k] ? 1 : 0;
  }
  const int e0_of = __shfl_sync(kFull, e0, k_of);
  int off_of = 0;
#pragma unroll
  for (int k = 1; k < kBulkRows; ++k) off_of = k_of >= k ? off[k] : off_of;
  if (lane < total) {
    const int e = e0_of + (lane - off_of);
    src = __ldg(p.e_src + e);
    rel = __ldg(p.e_rel + e);
  }
  const uint32_t ring = smem_u32(ring_raw) + warp * (kBulkSlots * 512);
  const float4* w4 = reinterpret_cast<const float4*>(p.weight) + lane;
  float4 acc[kBulkRows];
#pragma unroll
  for (int k = 0; k < kBulkRows; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  pdl_wait();   // the source rows are the previous layer's output
  uint32_t phase = 0;
  for (int base = 0; base < total; base += kBulkSlots) {
    const int cnt = min(kBulkSlots, total - base);
    fence_proxy_async();   // the slots' previous readers (generic proxy) before the engine's writes
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&bars[warp], static_cast<uint32_t>(cnt) * 512u);
    __syncwarp();
    if (lane >= base && lane < base + cnt) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       ring + (lane - base) * 512),
                   "l"(p.x + static_cast<size_t>(src) * kD), "r"(512), "r"(smem_u32(&bars[warp]))
                   : "memory");
    }
    mbar_wait(&bars[warp], phase);
    phase ^= 1;
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const int q = base + j;
      const int kq = __shfl_sync(kFull, k_of, q), rq = __shfl_sync(kFull, rel, q);
      const float nq = __shfl_sync(kFull, nrm, kq);
      float4 hv;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(hv.x), "=f"(hv.y), "=f"(hv.z), "=f"(hv.w) : "r"(ring + j * 512 + lane * 16));
      const float4 wv = __ldg(w4 + static_cast<size_t>(rq) * (kD / 4));
      // msg = (h * w) * norm_e, summed in edge order (RGCN.py:92-97)
      const float mx = (hv.x * wv.x) * nq, my = (hv.y * wv.y) * nq, mz = (hv.z * wv.z) * nq, mw = (hv.w * wv.w) * nq;
#pragma unroll
      for (int k = 0; k < kBulkRows; ++k) {
        if (kq == k) {   // warp-uniform
          acc[k].x += mx; acc[k].y += my; acc[k].z += mz; acc[k].w += mw;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kBulkRows; ++k) {
    const int rk = __shfl_sync(kFull, row, k);
    const float nk = __shfl_sync(kFull, nrm, k);
    if (k < n_rows) {
      float4 a = acc[k];
      a.x *= nk; a.y *= nk; a.z *= nk; a.w *= nk;  // apply_func (RGCN.py:103-104)
      reinterpret_cast<float4*>(p.agg_scratch + static_cast<size_t>(rk) * kD)[lane] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fused RGCN layer, tcgen05
// ------------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;
constexpr int kWorkerWarps = 8;
constexpr int kWorkers = kWorkerWarps * 32;
// Registers are allocated per 4 warps: 3 warpgroups of 168.  The control warpgroup (TMA warp 8, MMA warp 9, two idle
// warps) hands registers to the two worker warpgroups with setmaxnreg (88 / 208; 56 made the MMA issue loop spill).
constexpr int kLayerThreads = kWorkers + 128;
constexpr int kStages = 3;
constexpr int kBImage = kTileRows * kD * 4;             // 64 KB: one hi (or lo) image of the activation tile
constexpr int kLayerSmem = 2 * kBImage + kStages * kWChunkBytes + 1024;

struct LayerBars {
  uint64_t w_full[kStages], w_empty[kStages];
  uint64_t b_ready, d1_full, x_ready;
  uint64_t d2_full[3], d2_empty[3];
  uint32_t tmem_base;
  float* chain_peer[TEMP_MAX_PUSH_PEERS];   // snapshot-sharded forward: peer-mapped bases of the chained output
};

__global__ void __launch_bounds__(kLayerThreads, 1) rgcn_layer_tc_kernel(const TempRgcnLayerArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + kBImage;
  uint8_t* ring = smem + 2 * kBImage;
  __shared__ LayerBars S;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the same value, provably warp-uniform: the control-warp branches (and with them the MMA issue loop) then run on the
  // uniform datapath instead of a per-operand R2UR election loop (13 instructions per tcgen05.mma in round 1's SASS)
  const int warp_u = __shfl_sync(kFull, tid >> 5, 0);
  const int rbase = p.row0 + blockIdx.x * kTileRows;
  const int n_mb = p.chain_w_packed != nullptr ? p.chain_n >> 7 : 0;
  pdl_launch_dependents();
  TL_START();

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&S.w_full[i], 1);
      mbar_init(&S.w_empty[i], 1);
    }
    mbar_init(&S.b_ready, kWorkers);
    mbar_init(&S.d1_full, 1);
    mbar_init(&S.x_ready, kWorkers);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&S.d2_full[i], 1);
      mbar_init(&S.d2_empty[i], kWorkers);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&S.tmem_base, 512);
  if (p.chain_peers != nullptr && tid < p.chain_world) S.chain_peer[tid] = p.chain_peers[tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;
  TL(1);

  if (warp_u >= kWorkerWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == 8 && lane == 0) {
      // ===== TMA producer: GEMM1 weight chunks, then the chain's, in the order the MMA warp consumes them =====
      {
      const uint8_t* w1 = static_cast<const uint8_t*>(p.terms[0].w_packed);
      const uint8_t* wc = static_cast<const uint8_t*>(p.chain_w_packed);
      const int total = kKAtoms * (1 + n_mb);
      for (int i = 0; i < total; ++i) {
        const int st = i % kStages;
        if (i >= kStages) mbar_wait(&S.w_empty[st], ((i / kStages) - 1) & 1);
        const uint8_t* src = i < kKAtoms ? w1 + static_cast<size_t>(i) * kWChunkBytes
                                         : wc + static_cast<size_t>(i - kKAtoms) * kWChunkBytes;
        mbar_expect_tx(&S.w_full[st], kWChunkBytes);
        bulk_g2s(ring + st * kWChunkBytes, src, kWChunkBytes, &S.w_full[st]);
      }
    }
    } else if (warp_u == 9) {
      // ===== MMA issuer: the whole warp walks the pipeline (waits, stage arithmetic: convergent, on the uniform datapath);
      // only the tcgen05.mma / commit instructions are issued by the elected lane =====
      const bool leader = elect_one();
      const uint32_t tbase = __shfl_sync(kFull, S.tmem_base, 0);
      const uint32_t idesc = umma_idesc_tf32(128, kTileRows);
      const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo), rg = smem_u32(ring);
      int i = 0;
      mbar_wait(&S.b_ready, 0);
      tc_fence_after();
      TL(2);
#pragma unroll
      for (int ka = 0; ka < kKAtoms; ++ka, ++i) {   // (unrolled: the accumulate flag of each MMA is an immediate)
        const int st = i % kStages;
        mbar_wait(&S.w_full[st], (i / kStages) & 1);
        tc_fence_after();
        if (leader) {
          umma_katom_3x(tbase, rg + st * kWChunkBytes, bh + ka * (kTileRows * 128), bl + ka * (kTileRows * 128), idesc, ka == 0);
          umma_commit(&S.w_empty[st]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(&S.d1_full);
      __syncwarp();
      TL(3);
      if (n_mb > 0) {
        mbar_wait(&S.x_ready, 0);
        tc_fence_after();
        TL(5);
        for (int mb = 0; mb < n_mb; ++mb) {
          const int slot = mb % 3;
          if (mb >= 3) {
            mbar_wait(&S.d2_empty[slot], ((mb / 3) - 1) & 1);
            tc_fence_after();
          }
#pragma unroll
          for (int ka = 0; ka < kKAtoms; ++ka, ++i) {
            const int st = i % kStages;
            mbar_wait(&S.w_full[st], (i / kStages) & 1);
            tc_fence_after();
            if (leader) {
              umma_katom_3x(tbase + 128 + 128 * slot, rg + st * kWChunkBytes, bh + ka * (kTileRows * 128),
                            bl + ka * (kTileRows * 128), idesc, ka == 0);
              umma_commit(&S.w_empty[st]);
            }
            __syncwarp();
          }
          if (leader) umma_commit(&S.d2_full[slot]);
          __syncwarp();
          TL(6 + (mb < 3 ? mb : 3));
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ===== workers: warp (q, hf) owns feature block q (TMEM lane quadrant) of rows [64 hf, 64 hf + 64) =====
    const int q = warp & 3, hf = warp >> 2;
    const int f = 32 * q + lane;
    const int R0 = rbase + 64 * hf;
    const uint32_t my_off = static_cast<uint32_t>(q) * (kTileRows * 128);  // k-atom q of the activation tile
    const uint32_t sb_hi = smem_u32(b_hi) + my_off, sb_lo = smem_u32(b_lo) + my_off;
    const uint32_t lane_base = tbase + (static_cast<uint32_t>(32 * q) << 16);

    // ---- 1. self-loop operand rows -> shared memory (hi / lo) ---------------------------------------------
    // (any thread mapping works here: warp w stages rows 16 w .. 16 w + 15, a lane per 4 features, so that one load
    // instruction fetches a whole 512-byte row and one 16-byte store fills a swizzle chunk)
    {
      const TempDenseTerm& tm = p.terms[0];
      const int rr = rbase + 16 * warp + (lane & 15);
      const int idx = rr < p.row1 ? (tm.a_index != nullptr ? __ldg(tm.a_index + rr) : rr) : -1;
      pdl_wait();  // everything above (barriers, TMEM, packed-weight prefetch, plan indices) overlapped the predecessor
      float4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int sr = __shfl_sync(kFull, idx, i);
        v[i] = sr >= 0 ? ld_dep_f32x4(reinterpret_cast<const float4*>(tm.a + static_cast<size_t>(sr) * kD) + lane)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // row 16 w + i: 8-row group 2 w + (i >> 3), row in group i & 7; k-atom lane >> 3, 16-byte chunk (lane & 7) ^ (i & 7)
      const uint32_t base_off = static_cast<uint32_t>(lane >> 3) * (kTileRows * 128) + static_cast<uint32_t>(2 * warp) * 1024u;
      const uint32_t s_hi = smem_u32(b_hi) + base_off, s_lo = smem_u32(b_lo) + base_off;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float4 hi, lo;
        split_tf32(v[i].x, hi.x, lo.x);
        split_tf32(v[i].y, hi.y, lo.y);
        split_tf32(v[i].z, hi.z, lo.z);
        split_tf32(v[i].w, hi.w, lo.w);
        const uint32_t off = static_cast<uint32_t>(i >> 3) * 1024u + static_cast<uint32_t>(i & 7) * 128u +
                             (static_cast<uint32_t>((lane & 7) ^ (i & 7)) << 4);
        sts_f32x4(s_hi + off, hi);
        sts_f32x4(s_lo + off, lo);
      }
      fence_proxy_async();
      mbar_arrive(&S.b_ready);
      TL(2);
    }

    // ---- 2. the aggregate agg[r] = norm_r * sum_e (x[src_e] * W[rel_e]) * norm_r was written by rgcn_gather_kernel
    // (rows with in-edges only): bit j of `has` = row R0 + j has in-edges.
    unsigned long long has = 0ull;
    if (p.row_ptr != nullptr) {
      const int ra = min(R0 + lane, p.row1), rb = min(R0 + 32 + lane, p.row1);
      const int pa0 = __ldg(p.row_ptr + ra), pa1 = __ldg(p.row_ptr + min(ra + 1, p.row1));
      const int pb0 = __ldg(p.row_ptr + rb), pb1 = __ldg(p.row_ptr + min(rb + 1, p.row1));
      has = static_cast<unsigned long long>(__ballot_sync(kFull, pa1 > pa0)) |
            (static_cast<unsigned long long>(__ballot_sync(kFull, pb1 > pb0)) << 32);
    }

    TL(3);
    // ---- 3. epilogue 1: out = act(agg (+x) + x . W_loop + bias) ; h_out ; chain operand X ---------------
    const float bias = p.h_bias != nullptr ? __ldg(p.h_bias + f) : 0.f;
    const bool need_te = (p.te_out | p.te_chain) != 0;
    int rtA = p.row_time_scalar, rtB = p.row_time_scalar;
    if (need_te && p.row_time != nullptr) {
      rtA = __ldg(p.row_time + min(R0 + lane, p.row1 - 1));
      rtB = __ldg(p.row_time + min(R0 + 32 + lane, p.row1 - 1));
    }
    // Rows are packed snapshot by snapshot, so the 64 rows of this thread see one or two time-embedding rows almost
    // always: both are fetched up front and a row picks one by its index (fast path, no per-row shuffle / branch / load).
    // Spans with three or more snapshots (tiny graphs) take the general walk that reloads on change.
    const int t_first = need_te ? __shfl_sync(kFull, rtA, 0) : -1, t_last = need_te ? __shfl_sync(kFull, rtB, 31) : -1;
    float te_a = 0.f, te_b = 0.f;
    int te_split = 64;      // rows [0, te_split) of the span use te_a, the rest te_b
    bool te_fast = true;
    if (need_te) {
      te_a = __ldg(p.time_embed + static_cast<size_t>(t_first) * kD + f);
      te_b = __ldg(p.time_embed + static_cast<size_t>(t_last) * kD + f);
      const unsigned long long is_first = static_cast<unsigned long long>(__ballot_sync(kFull, rtA == t_first)) |
                                          (static_cast<unsigned long long>(__ballot_sync(kFull, rtB == t_first)) << 32);
      const bool two = __all_sync(kFull, (rtA == t_first || rtA == t_last) && (rtB == t_first || rtB == t_last));
      te_split = __popcll(is_first);
      te_fast = two && is_first == (te_split >= 64 ? ~0ull : ((1ull << te_split) - 1ull));
    }
    // the aggregate rows of this thread's 64 rows: all loads in flight while the self-loop MMA completes
    const float* agg_col = p.agg_scratch + static_cast<size_t>(R0) * kD + f;
    float ag[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) ag[j] = ((has >> j) & 1ull) ? ld_dep_f32(agg_col + j * kD) : 0.f;
    mbar_wait(&S.d1_full, 0);
    tc_fence_after();
    TL(4);
    int cur_trow = t_first;
    float cur_te = te_a;
    // What a row needs from the argument block, hoisted and made opaque (the empty asm): left to itself the compiler re-reads
    // the flags from the constant bank and rebuilds the 64-bit row address inside a reconvergence region for every one of
    // the 64 unrolled rows (~35 instructions per row; found on the 64-row kernel of tc_wide.cu, same pattern here).
    float* hp = p.h_out + static_cast<size_t>(R0) * kD + f;
    int n_store = p.h_out != nullptr ? max(0, min(64, p.row1 - R0)) : 0;      // rows of this thread that reach h_out
    int n_keep = max(0, min(64, p.row1 - R0));                                // rows that are not operand padding
    unsigned fl = (p.activation == TEMP_ACT_RELU ? 1u : 0u) | (p.residual ? 2u : 0u) | (p.te_out ? 4u : 0u) | (p.te_chain ? 8u : 0u) |
                  (n_mb > 0 ? 16u : 0u);
    asm volatile("" : "+l"(hp), "+r"(n_store), "+r"(n_keep), "+r"(fl));
    const bool f_relu = fl & 1u, f_res = fl & 2u, f_teo = fl & 4u, f_tec = fl & 8u, f_chain = fl & 16u;
    auto epilogue1 = [&](auto fast_tag) {
      constexpr bool kFast = decltype(fast_tag)::value;
#pragma unroll
      for (int c = 0; c < 2; ++c) {          // 32 rows per TMEM load (two load latencies per thread instead of four)
        float v[32];
        tmem_ld32(lane_base + 64 * hf + 32 * c, v);
        const int rt = c == 0 ? rtA : rtB;
        const uint32_t grp = (static_cast<uint32_t>(8 * hf + 4 * c)) * 1024u;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t off = grp + sw128_off(i, lane);
          float te;
          if (kFast) {
            te = 32 * c + i < te_split ? te_a : te_b;
          } else {
            const int trow = __shfl_sync(kFull, rt, i);  // warp-uniform
            if (trow != cur_trow) {
              cur_trow = trow;
              cur_te = __ldg(p.time_embed + static_cast<size_t>(trow) * kD + f);
            }
            te = cur_te;
          }
          float val = ag[32 * c + i];
          if (f_res) val += lds_f32(sb_hi + off) + lds_f32(sb_lo + off);
          val += v[i];
          val += bias;
          if (f_relu) val = fmaxf(val, 0.f);
          const float with_te = val + te;
          if (32 * c + i < n_store) hp[(32 * c + i) * kD] = f_teo ? with_te : val;
          if (f_chain) {
            const float xx = 32 * c + i < n_keep ? (f_tec ? with_te : val) : 0.f;
            float hi, lo;
            split_tf32(xx, hi, lo);
            sts_f32(sb_hi + off, hi);
            sts_f32(sb_lo + off, lo);
          }
        }
      }
    };
    if (te_fast)
      epilogue1(std::true_type{});
    else
      epilogue1(std::false_type{});

    TL(5);
    // ---- 4. chain epilogues: chain_out[r, 128 mb + f] = D2 + chain_b -----------------------------------
    if (n_mb > 0) {
      fence_proxy_async();
      mbar_arrive(&S.x_ready);
      // snapshot-sharded forward: the rank that scans the chain partition of a row receives the row's chained output
      int ownA = 0, ownB = 0;
      const bool use_peers = p.chain_peers != nullptr;
      if (use_peers) {
        ownA = __ldg(p.chain_owner + min(R0 + lane, p.row1 - 1));
        ownB = __ldg(p.chain_owner + min(R0 + 32 + lane, p.row1 - 1));
      }
      for (int mb = 0; mb < n_mb; ++mb) {
        const int slot = mb % 3;
        const float cbias = p.chain_b != nullptr ? __ldg(p.chain_b + 128 * mb + f) : 0.f;
        mbar_wait(&S.d2_full[slot], (mb / 3) & 1);
        tc_fence_after();
        TL(6 + 2 * (mb < 3 ? mb : 3));
        float* orow = p.chain_out + static_cast<size_t>(R0) * p.chain_ld + 128 * mb + f;
        size_t chain_pitch = static_cast<size_t>(p.chain_ld) * sizeof(float);
        asm volatile("" : "+l"(orow), "+l"(chain_pitch));      // (running pointer: see epilogue 1)
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32(lane_base + 128 + 128 * slot + 64 * hf + 32 * c, v);
          tmem_ld_wait();
          if (!use_peers) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (32 * c + i < n_keep) *orow = v[i] + cbias;
              orow = reinterpret_cast<float*>(reinterpret_cast<char*>(orow) + chain_pitch);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int own = __shfl_sync(kFull, c == 0 ? ownA : ownB, i);
              if (R0 + 32 * c + i < p.row1)
                S.chain_peer[own][static_cast<size_t>(R0 + 32 * c + i) * p.chain_ld + 128 * mb + f] = v[i] + cbias;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&S.d2_empty[slot]);
        TL(7 + 2 * (mb < 3 ? mb : 3));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------------------------------------
// chain-partitioned GRU scan, tcgen05
// ------------------------------------------------------------------------------------------------
// The recurrence couples a packed row only with the row of the SAME entity in the SAME batch item at the previous
// step (prev_row), never with other rows.  The planner therefore cuts every batch item into "chain partitions"
// (entity-id ranges): per scan step a partition owns one contiguous row range of at most kScanN rows and no
// dependency ever crosses a partition.  One cluster of 4 CTAs (one per block of 32 hidden columns, its W_hh
// slice resident in shared memory) walks all steps of a partition; steps are separated by a hardware cluster
// barrier (release / acquire, ~0.2 us) instead of a grid-wide barrier, and different partitions never synchronise.
constexpr int kScanN = 96;                     // max packed rows per partition step
constexpr int kScanHalf = kScanN / 2;          // rows per worker group = UMMA N of one MMA batch
constexpr int kScanCluster = 4;                // CTAs per cluster = blocks of 32 hidden columns
// 16 warps = 4 per SM sub-partition = 128 registers per thread: worker group 0 = warps 0..7 (tile rows gw + 8u),
// worker group 1 = warps 8..14 (tile rows 48 + gw + 7u), warp 15 = control (TMEM allocation, bulk copies, MMA issue).
// (A 17th warp capped the kernel at 96 registers and the spills sat on the per-step critical path.)
constexpr int kScanCtlWarp = 15;
constexpr int kScanThreads = 16 * 32;
constexpr int kScanCtlTid = kScanCtlWarp * 32;
constexpr int kScanU = 7;                      // row slots per worker thread and step (group 0 uses 6 of them)
constexpr int kScanAImage = kKAtoms * kWChunkBytes;   // 128 KB: this CTA's W_hh slice (r|z|n|pad rows, hi + lo)
constexpr int kScanBImage = kScanN * kD * 4;   // 48 KB per hi / lo
constexpr int kScanSmem = kScanAImage + 2 * kScanBImage + 1024;
static_assert(kScanHalf == 48, "the worker groups' row maps (8 x 6 and 7 x 7 slots) assume 48-row half tiles");
static_assert(kScanHalf * 32 * 4 == kScanHalf * 128, "the gate exchange rows alias operand rows one to one");

struct ScanBars {
  uint64_t w_full, mma_done[2];
  uint32_t tmem_base;
  float* push[TEMP_MAX_PUSH_PEERS];  // peer-mapped buffer bases of the fused all-gather (offset applied)
};

// Publishing a step: every thread's state stores are ordered before ONE thread's cluster-scope release fence by a CTA
// barrier (cumulativity), then all threads arrive relaxed -- one memory barrier per CTA and step instead of one per warp.
__device__ __forceinline__ void cluster_publish() {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("fence.acq_rel.cluster;" ::: "memory");
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int group) {  // the 8 (group 0) or 7 (group 1) warps of one worker group
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(group == 0 ? 256 : 224) : "memory");
}

// fast gate math: ex2.approx / rcp.approx based, absolute error ~1e-7 (the parity bar is 1e-4 relative)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// What the gather of a step needs, fetched one step ahead (while the previous step's MMA runs) so that the serial
// chain per step is only: cluster barrier -> L2 gather of the previous state -> MMA -> gates -> state write.
struct ScanPre {
  int prv;      // lanes 0..kScanU-1: previous-state row of this thread's u-th tile row, -1 = zero state
  float dt;     // lanes 0..kScanU-1: its time gap (the decay factor is computed at use)
};

// tile row of slot u of worker warp gw of group g: g * 48 + gw + stride * u (stride 8 / 7), valid while gw + stride * u < 48
__device__ __forceinline__ void scan_prefetch(const TempGruArgs& p, int rb, int r1, int row0, int stride, int lim, int lane,
                                              ScanPre& f) {
  f.prv = -1;
  f.dt = 0.f;
  if (lane < kScanU) {
    const int i = row0 + stride * lane;
    const int r = rb + i;
    if (i < lim && r < r1 && p.prev_row != nullptr) {
      f.prv = __ldg(p.prev_row + r);
      if (p.dt != nullptr) f.dt = __ldg(p.dt + r);
    }
  }
}

__global__ void __cluster_dims__(kScanCluster, 1, 1) __launch_bounds__(kScanThreads, 1)
    gru_scan_tc_kernel(const TempGruScanArgs P, const int n_parts) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* a_img = smem;
  uint8_t* b_hi = a_img + kScanAImage;
  uint8_t* b_lo = b_hi + kScanBImage;
  __shared__ ScanBars S;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < kScanCtlWarp;
  const int grp = warp >> 3, gw = warp & 7;                 // (the control warp computes with grp 1, gw 7: unused)
  const int stride = grp == 0 ? 8 : 7;
  const int row0 = grp * kScanHalf + gw;                    // this thread's tile rows: row0 + stride * u, below `lim`
  const int lim = (grp + 1) * kScanHalf;
  const int cb = blockIdx.x & (kScanCluster - 1), jb = 32 * cb;   // == %cluster_ctarank for a 1-D grid
  const int cid = blockIdx.x / kScanCluster, n_clusters = gridDim.x / kScanCluster;
  // Gate exchange buffer of a group: gate g, tile row i -> 32 floats at operand row i of k-atom g of the hi image.
  // Those bytes are dead once the group's MMA batch has completed (the other batch reads other rows only).
  const uint32_t s_bhi = smem_u32(b_hi), s_blo = smem_u32(b_lo);
  const uint32_t ex = s_bhi;            // shared-space byte address of the exchange buffer
  constexpr int kExGate = kScanN * 128; // bytes between gates (= one k-atom block)

  if (tid == 0) {
    mbar_init(&S.w_full, 1);
    mbar_init(&S.mma_done[0], 1);
    mbar_init(&S.mma_done[1], 1);
    fence_mbar_init();
  }
  if (warp == kScanCtlWarp) tmem_alloc(&S.tmem_base, 128);
  if (P.push_bufs != nullptr && tid < P.push_world) S.push[tid] = P.push_bufs[tid] + P.push_offset;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;
  pdl_launch_dependents();
  TL_START();

  // Iteration order of a cluster: its partitions cid + j * n_clusters are taken in rounds of 32 (lane j keeps the row
  // range of partition j of the round at the current step); inside a round the order is STEP-MAJOR: step s of every
  // partition of the round, then step s + 1.  With two or more partitions in a round, consecutive tile steps are
  // independent of each other: the cluster barrier between them is already satisfied when it is reached and the next
  // tile step's previous-state rows are fetched while the current one computes (only with a single partition per round
  // -- the small shapes -- each step waits for its predecessor, as the recurrence demands).
  auto load_ranges = [&](int round, int st) -> int2 {
    int2 rg = make_int2(0, 0);
    const int part = cid + (32 * round + lane) * n_clusters;
    if (part < n_parts && st < P.n_steps) {
      if (P.parts != nullptr) {
        rg = __ldg(reinterpret_cast<const int2*>(P.parts) + static_cast<size_t>(part) * P.part_stride + P.steps[st].part_col);
      } else {  // single step without a partition table: plain row tiles
        rg.x = P.steps[st].row0 + part * kScanN;
        rg.y = min(rg.x + kScanN, P.steps[st].row1);
      }
    }
    return rg;
  };
  // first partition >= j of the round that has rows at this step (-1 when none)
  auto first_from = [&](int2 rg, int j) -> int {
    const unsigned has = j < 32 ? (__ballot_sync(kFull, rg.y > rg.x) & ~((1u << j) - 1u)) : 0u;
    return has != 0u ? __ffs(has) - 1 : -1;
  };
  const int n_rounds = (n_parts - cid + 32 * n_clusters - 1) / (32 * n_clusters);   // cid < n_parts by construction
  // next tile step at or after (round, st, j): advances through steps, then rounds; round == n_rounds when exhausted
  auto seek = [&](int& round, int& st, int& j, int2& rg) {
    while (round < n_rounds) {
      j = first_from(rg, j);
      if (j >= 0) return;
      j = 0;
      if (++st >= P.n_steps) {
        st = 0;
        ++round;
      }
      if (round < n_rounds) rg = load_ranges(round, st);
    }
  };

  const void* cur_w = nullptr;
  uint32_t w_phase = 0, mma_phase[2] = {0, 0};  // the second batch's barrier only completes on steps with two halves
  bool w_pending = false;
  bool arrived = false;  // a cluster-barrier arrive that has not been waited on yet

  int round = 0, s = 0, j = 0;
  int2 ranges = load_ranges(0, 0);
  seek(round, s, j, ranges);
  ScanPre cur;
  if (round < n_rounds && worker)
    scan_prefetch(P.steps[s], __shfl_sync(kFull, ranges.x, j), __shfl_sync(kFull, ranges.y, j), row0, stride, lim, lane, cur);
  float4 pv[kScanU];        // previous-state rows fetched one tile step ahead (valid when have_pv)
  bool have_pv = false;
  pdl_wait();  // gi (and, for a single step, the previous state) come from the predecessor kernels

#pragma unroll 1
  while (round < n_rounds) {
    const TempGruArgs& p = P.steps[s];
    const int rb = __shfl_sync(kFull, ranges.x, j), r1 = __shfl_sync(kFull, ranges.y, j);
    // successor in the tile-step sequence of this cluster
    int n_round = round, n_s = s, n_j = j + 1;
    int2 n_ranges = ranges;
    seek(n_round, n_s, n_j, n_ranges);
    const bool more = n_round < n_rounds;
    // the successor continues THIS partition (its input is what this tile step writes) -> nothing to fetch ahead
    const bool dependent = more && n_round == round && n_j == j;

    if (p.prev_row != nullptr && p.whh_packed != cur_w) {
      // every MMA that read the old image has completed (mma_done is waited on inside each step); a copy that
      // no step consumed (its steps had no previous state at all) is drained before the buffer is refilled
      if (w_pending) {
        if (tid == kScanCtlTid) mbar_wait(&S.w_full, w_phase);
        w_phase ^= 1;
      }
      if (tid == kScanCtlTid) {
        const uint8_t* src = static_cast<const uint8_t*>(p.whh_packed) + static_cast<size_t>(cb) * kScanAImage;
        mbar_expect_tx(&S.w_full, kScanAImage);
        for (int c = 0; c < kKAtoms; ++c) bulk_g2s(a_img + c * kWChunkBytes, src + c * kWChunkBytes, kWChunkBytes, &S.w_full);
      }
      cur_w = p.whh_packed;
      w_pending = true;
    }
    const bool type1 = p.cell_type == TEMP_CELL_TYPE1;
    const int rows = r1 - rb;
    const bool two = rows > kScanHalf;  // the second half tile has rows (cluster-uniform)
    if (j == 0 && round == 0) TL(2 + 6 * (s & 7));
    if (arrived) {  // the other column blocks' state writes of the previous step become visible here; also every
      cluster_wait();  // thread of this CTA is past its reads of the operand tile and of the exchange buffer
      arrived = false;
    }
    if (j == 0 && round == 0) TL(3 + 6 * (s & 7));

    // ---- previous-state rows -> smem operand (hi / lo) ------------------------------------------------------
    // worker (group, gw, lane) gathers tile rows row0 + 8u (lane = 4 feature columns) and later does the gate
    // math for hidden column jb + lane of the same rows.
    int any_prev = 0;
    if (worker) {
      float4 v[kScanU];
      const float dec = p.dt != nullptr ? decay_factor(cur.dt, p.decay_wb, p.inv_temperature) : 1.f;
#pragma unroll
      for (int u = 0; u < kScanU; ++u) {
        const int pr = __shfl_sync(kFull, cur.prv, u);
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pr >= 0) {
          any_prev = 1;
          v[u] = have_pv ? pv[u] : __ldcg(reinterpret_cast<const float4*>(p.state + static_cast<size_t>(pr) * kD + 4 * lane));
        }
      }
#pragma unroll
      for (int u = 0; u < kScanU; ++u) {
        const int i = row0 + stride * u;
        const float dv = __shfl_sync(kFull, dec, u);
        if (i >= lim) continue;  // warp-uniform: slot beyond the group's half tile
        float4 hi, lo;
        split_tf32(v[u].x * dv, hi.x, lo.x);
        split_tf32(v[u].y * dv, hi.y, lo.y);
        split_tf32(v[u].z * dv, hi.z, lo.z);
        split_tf32(v[u].w * dv, hi.w, lo.w);
        const uint32_t off = static_cast<uint32_t>(lane >> 3) * (kScanN * 128) + (i >> 3) * 1024u + (i & 7) * 128u +
                             (((lane & 7) ^ (i & 7)) << 4);
        sts_f32x4(s_bhi + off, hi);
        sts_f32x4(s_blo + off, lo);
      }
      fence_proxy_async();
    }
    any_prev = __syncthreads_or(any_prev);
    if (j == 0 && round == 0) TL(4 + 6 * (s & 7));
    if (any_prev && tid == kScanCtlTid) {
      if (w_pending) mbar_wait(&S.w_full, w_phase);
      tc_fence_after();
      const uint32_t ai = smem_u32(a_img), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
      // ONE MMA batch over the whole tile (N = rows rounded up to 16): every tcgen05.mma re-reads its 4 KB weight
      // slice from shared memory, so fewer, wider MMAs are cheaper than one batch per half tile
      const uint32_t idesc = umma_idesc_tf32(128, min(kScanN, (rows + 15) & ~15));
      for (int ka = 0; ka < kKAtoms; ++ka)
        umma_katom_3x(tbase, ai + ka * kWChunkBytes, bh + ka * (kScanN * 128), bl + ka * (kScanN * 128), idesc, ka == 0);
      umma_commit(&S.mma_done[0]);
      if (two) umma_commit(&S.mma_done[1]);

    }
    if (any_prev && w_pending) {
      w_pending = false;
      w_phase ^= 1;
    }
    if (warp == kScanCtlWarp && more) {
      // The control warp's lanes pull the NEXT tile step's input-gate lines (3 x 128 B per row for this CTA's hidden
      // columns) into L2 now: on HBM-resident shapes the gi loads of all CTAs otherwise hit DRAM in one burst inside
      // the MMA window of the next tile step (measured 9.5 k cycles per window at x16 against 4.3 k at x1).
      const TempGruArgs& np_ = P.steps[n_s];
      const int nb = __shfl_sync(kFull, n_ranges.x, n_j), n1 = __shfl_sync(kFull, n_ranges.y, n_j);
      const int gates = np_.cell_type == TEMP_CELL_TYPE1 ? 1 : 3;
      const int lines = (n1 - nb) * gates;
      for (int i = lane; i < lines; i += 32) {
        const int r = nb + i / gates, g = i - (i / gates) * gates;
        const float* a = np_.gi + static_cast<size_t>(r) * np_.gi_ld + np_.gi_off + g * kD + jb;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
    }
    // ---- while the MMA runs: this thread's h0 values and input-gate pre-activations; the NEXT step's indices ----
    const bool active = worker && (grp == 0 || two);  // this group's half tile has rows
    float h0[kScanU], gi_r[kScanU], gi_z[kScanU], gi_n[kScanU];
    float br = 0.f, bz = 0.f, bn = 0.f, te_uni = 0.f;
    bool te_rows = false;
    ScanPre nxt;
    if (worker) {
      const int j = jb + lane;
#pragma unroll
      for (int u = 0; u < kScanU; ++u) {
        h0[u] = 0.f;
        if (any_prev && row0 + stride * u < lim) {
          const uint32_t off = static_cast<uint32_t>(cb) * (kScanN * 128) + sw128_off(row0 + stride * u, lane);
          h0[u] = lds_f32(s_bhi + off) + lds_f32(s_blo + off);
        }
      }
      br = __ldg(p.b_hh + j);
      bz = __ldg(p.b_hh + kD + j);
      bn = __ldg(p.b_hh + 2 * kD + j);
      // rows of one partition step belong to one snapshot instance: a single time-embedding row (checked)
      int trow0 = p.row_time_scalar, trow1 = p.row_time_scalar;
      if (p.time_embed != nullptr && p.row_time != nullptr) {
        trow0 = __ldg(p.row_time + rb);
        trow1 = __ldg(p.row_time + r1 - 1);
      }
#pragma unroll
      for (int u = 0; u < kScanU; ++u) {
        const int r = rb + row0 + stride * u;
        gi_r[u] = gi_z[u] = gi_n[u] = 0.f;
        if (row0 + stride * u < lim && r < r1) {
          const float* gi = p.gi + static_cast<size_t>(r) * p.gi_ld + p.gi_off + j;
          if (type1) {
            gi_n[u] = ld_dep_f32(gi);
          } else {
            gi_r[u] = ld_dep_f32(gi);
            gi_z[u] = ld_dep_f32(gi + kD);
            gi_n[u] = ld_dep_f32(gi + 2 * kD);
          }
        }
      }
      if (more) {
        const TempGruArgs& np_ = P.steps[n_s];
        scan_prefetch(np_, __shfl_sync(kFull, n_ranges.x, n_j), __shfl_sync(kFull, n_ranges.y, n_j), row0, stride, lim, lane, nxt);
        if (!dependent && np_.prev_row != nullptr) {  // its producer tile step lies at least one cluster barrier back
#pragma unroll
          for (int u = 0; u < kScanU; ++u) {
            const int pr = __shfl_sync(kFull, nxt.prv, u);
            pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pr >= 0) pv[u] = __ldcg(reinterpret_cast<const float4*>(np_.state + static_cast<size_t>(pr) * kD + 4 * lane));
          }
        }
      }
      te_rows = p.time_embed != nullptr && trow0 != trow1;  // tiles of a launch without partition table may mix snapshots
      if (p.time_embed != nullptr && !te_rows) te_uni = __ldg(p.time_embed + static_cast<size_t>(trow0) * kD + j);
    }
    if (any_prev && worker) {
      group_sync(grp);  // every h0 read of this group's operand rows is done before the exchange overwrites them
      if (active) {
        mbar_wait(&S.mma_done[grp], grp == 0 ? mma_phase[0] : mma_phase[1]);
        tc_fence_after();
        if (j == 0 && round == 0) TL(5 + 6 * (s & 7));
        if ((gw & 3) < 3) {  // TMEM lane quadrant = gate; columns = rows of the half tile, 24 per warp
          const int gate = gw & 3, hf = gw >> 2;
          constexpr int kQ = kScanHalf / 2;  // 24
          const uint32_t ta = tbase + (static_cast<uint32_t>(32 * gate) << 16) + grp * kScanHalf + kQ * hf;
          const uint32_t exw = ex + gate * kExGate + (grp * kScanHalf + kQ * hf) * 128 + lane * 4;
          float v[16], w[8];
          tmem_ld16(ta, v);
          tmem_ld8(ta + 16, w);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) sts_f32(exw + i * 128, v[i]);
#pragma unroll
          for (int i = 0; i < 8; ++i) sts_f32(exw + (16 + i) * 128, w[i]);
        }
        tc_fence_before();
      }
      group_sync(grp);
    }
    if (any_prev) {
      mma_phase[0] ^= 1;
      if (two) mma_phase[1] ^= 1;
    }
    if (j == 0 && round == 0) TL(6 + 6 * (s & 7));

    if (active) {
      const int j = jb + lane;
#pragma unroll
      for (int u = 0; u < kScanU; ++u) {
        const int i = row0 + stride * u;
        const int r = rb + i;
        if (i < lim && r < r1) {  // warp-uniform
          float hr = br, hz = bz, hn = bn;
          if (any_prev) {
            const uint32_t ea = ex + i * 128 + lane * 4;
            hr += lds_f32(ea);
            hz += lds_f32(ea + kExGate);
            hn += lds_f32(ea + 2 * kExGate);
          }
          float hy;
          if (type1) {  // GRU_cell.py:22-29
            const float rg = fast_sigmoid(hr), zg = fast_sigmoid(hz);
            const float ng = fast_tanh(gi_n[u] + rg * hn);
            hy = ng + zg * (h0[u] - ng);
          } else {      // torch.nn.GRU, gate order r, z, n
            const float rg = fast_sigmoid(gi_r[u] + hr);
            const float zg = fast_sigmoid(gi_z[u] + hz);
            const float ng = fast_tanh(gi_n[u] + rg * hn);
            hy = (1.f - zg) * ng + zg * h0[u];
          }
          hy += te_rows ? __ldg(p.time_embed + static_cast<size_t>(__ldg(p.row_time + r)) * kD + j) : te_uni;
          float* o = p.out + static_cast<size_t>(r) * kD + j;
          if (p.accumulate) hy += __ldcg(o);
          *o = hy;
          if (p.push != 0 && P.push_bufs != nullptr) {  // fused all-gather: NVLink stores into every peer's slab
            const size_t po = static_cast<size_t>(r - P.push_row0) * kD + j;
            if (P.push_multicast != nullptr) {  // one store, replicated to every GPU by the NVSwitch (NVLS)
              asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(P.push_multicast + P.push_offset + po), "f"(hy)
                           : "memory");
            } else {
              for (int k = 0; k < P.push_world; ++k) S.push[k][po] = hy;
            }
          }
        }
      }
    }
    if (j == 0 && round == 0) TL(7 + 6 * (s & 7));
    // publishes this step's state columns to the cluster (release); the matching wait of the next step also
    // orders the reuse of the operand tile and of the exchange buffer
    cluster_publish();
    arrived = true;
    cur = nxt;
    have_pv = more && !dependent && P.steps[n_s].prev_row != nullptr;
    round = n_round;
    s = n_s;
    j = n_j;
    ranges = n_ranges;
  }
  if (w_pending && tid == kScanCtlTid) mbar_wait(&S.w_full, w_phase);  // drain a copy no step consumed
  if (arrived) cluster_wait();

  tc_fence_before();
  __syncthreads();
  if (warp == kScanCtlWarp) tmem_dealloc(tbase, 128);
}

// cudaLaunchKernelEx with programmatic stream serialization (see pdl_wait in tc_common.cuh)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename K>
int ensure_smem_once(K kernel, int bytes, const char* name, bool& done) {
  if (done) return TEMP_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return temp_internal::cuda_fail(e, name);
  done = true;
  return TEMP_OK;
}

}  // namespace

#ifdef TEMP_TIMELINE
extern "C" int temp_debug_timeline(void* device_buffer) {  // [ctas][20 warps][64 slots] u64, or null to disable
  unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
  return cudaMemcpyToSymbol(g_timeline, &p, sizeof(p)) == cudaSuccess ? 0 : -2;
}
#endif

namespace temp_internal {

// d == 128 with 1x1 relation blocks: the 128-row tile kernel of this file; every other supported shape: tc_wide.cu
static bool wide_shape(const TempRgcnLayerArgs* a) { return a->d != kD || (a->row_ptr != nullptr && a->si != 1); }

// thread blocks of the aggregation launch that precedes the tile kernel (0: no launch)
int tc_gather_grid(const TempRgcnLayerArgs* a) {
  if (a->row_ptr == nullptr || a->row1 <= a->row0) return 0;
  if (a->agg_lists != 0) return a->n_agg_heavy + (a->n_agg_rows + kGatherWarps - 1) / kGatherWarps;
  return (a->row1 - a->row0 + kGatherWarps - 1) / kGatherWarps;
}

// kernel launches of the aggregation (0 or 1)
int tc_gather_launches(const TempRgcnLayerArgs* a) { return tc_gather_grid(a) > 0 ? 1 : 0; }

// the aggregation launch alone (also behind temp_rgcn_gather_fwd).  LDG rows (a warp per destination row, two feature
// rows in flight) or the TMA bulk-copy ring: measured on B200 (tools/bench_gather.py, x16 shapes, L2 flushed) the ring is
// 5 % faster on ICEWS14-shaped rows (55.3 vs 58.3 us) and 33 % slower on GDELT-shaped rows (697 vs 523 us: 64 KB of ring per
// CTA leaves 24 instead of 40 warps per SM for the high in-degree rows) -- LDG is the default, TEMP_GATHER=bulk opts in.
int tc_launch_gather(const TempRgcnLayerArgs* a, cudaStream_t st) {
  if (wide_shape(a)) return tcw_launch_gather(a, st);
  static const char* mode = getenv("TEMP_GATHER");
  const bool bulk = mode != nullptr && strcmp(mode, "bulk") == 0;
  if (bulk && a->agg_lists != 0) {
    static bool configured = false;
    if (int rc = ensure_smem_once(rgcn_gather_bulk_kernel, kBulkSmem, "rgcn_gather_bulk_kernel", configured)) return rc;
    const int batches = (a->n_agg_rows + kBulkRows - 1) / kBulkRows;
    const int grid = a->n_agg_heavy + (batches + kGatherWarps - 1) / kGatherWarps;
    if (grid <= 0) return TEMP_OK;
    cudaError_t e = launch_pdl(rgcn_gather_bulk_kernel, grid, kGatherWarps * 32, kBulkSmem, st, *a);
    if (e != cudaSuccess) return cuda_fail(e, "rgcn_gather_bulk_kernel launch");
    return TEMP_OK;
  }
  const int grid = tc_gather_grid(a);
  if (grid <= 0) return TEMP_OK;
  cudaError_t e = launch_pdl(rgcn_gather_kernel, grid, kGatherWarps * 32, 0, st, *a);
  if (e != cudaSuccess) return cuda_fail(e, "rgcn_gather_kernel launch");
  return TEMP_OK;
}

bool tc_layer_supported(const TempRgcnLayerArgs* a) {
  if (wide_shape(a)) return tcw_layer_supported(a);
  if (a->d != kD || a->n_terms != 1) return false;
  const TempDenseTerm& t = a->terms[0];
  if (t.w_packed == nullptr || t.a_dt != nullptr) return false;
  if (a->row_ptr != nullptr && (a->si != 1 || a->so != 1 || a->agg_scratch == nullptr)) return false;
  if (a->agg_lists != 0 && (a->n_agg_rows < 0 || a->n_agg_heavy < 0 || (a->n_agg_rows > 0 && a->agg_rows == nullptr) ||
                            (a->n_agg_heavy > 0 && a->agg_heavy == nullptr)))
    return false;
  if (a->chain_w != nullptr && (a->chain_w_packed == nullptr || (a->chain_n & 127) != 0)) return false;
  if (a->chain_peers != nullptr && (a->chain_owner == nullptr || a->chain_world <= 0 || a->chain_world > TEMP_MAX_PUSH_PEERS))
    return false;
  return true;
}

int tc_launch_layer(const TempRgcnLayerArgs* a, cudaStream_t st) {
  if (wide_shape(a)) return tcw_launch_layer(a, st);
  static bool configured = false;
  if (int rc = ensure_smem_once(rgcn_layer_tc_kernel, kLayerSmem, "rgcn_layer_tc_kernel", configured)) return rc;
  const int rows = a->row1 - a->row0;
  const int gather_grid = tc_gather_grid(a);
  if (gather_grid > 0) {
    if (int rc = tc_launch_gather(a, st)) return rc;
  }
  const int grid = (rows + kTileRows - 1) / kTileRows;
  cudaError_t e = launch_pdl(rgcn_layer_tc_kernel, grid, kLayerThreads, kLayerSmem, st, *a);
  if (e != cudaSuccess) return cuda_fail(e, "rgcn_layer_tc_kernel launch");
  return TEMP_OK;
}

bool tc_scan_supported(const TempGruScanArgs* a) {
  if (tc_scan2_supported(a)) return true;
  if (a->n_steps <= 0) return false;
  if (a->parts != nullptr && a->part_rows > kScanN) return false;
  if (a->parts == nullptr && a->n_steps != 1) return false;  // several steps need the chain partition table
  for (int s = 0; s < a->n_steps; ++s) {
    const TempGruArgs& g = a->steps[s];
    if (g.d != kD) return false;
    if (g.prev_row != nullptr && g.whh_packed == nullptr) return false;
    if (a->parts != nullptr && (g.part_col < 0 || g.part_col >= a->part_stride)) return false;
  }
  if (a->push_bufs != nullptr && (a->push_world <= 0 || a->push_world > TEMP_MAX_PUSH_PEERS)) return false;
  return true;
}

int tc_launch_scan(const TempGruScanArgs* a, cudaStream_t st) {
  if (tc_scan2_supported(a)) return tc_launch_scan2(a, st);
  static bool configured = false;
  if (int rc = ensure_smem_once(gru_scan_tc_kernel, kScanSmem, "gru_scan_tc_kernel", configured)) return rc;
  int n_parts = a->n_parts;
  if (a->parts == nullptr) n_parts = (a->steps[0].row1 - a->steps[0].row0 + kScanN - 1) / kScanN;
  if (n_parts <= 0) return TEMP_OK;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kScanThreads);
  cfg.dynamicSmemBytes = kScanSmem;
  cfg.stream = st;
  static int max_clusters = 0;  // co-resident clusters of 4 (one CTA per SM: 216 KB of shared memory)
  if (max_clusters == 0) {
    cfg.gridDim = dim3(kScanCluster * 64);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, gru_scan_tc_kernel, &cfg);
    if (e != cudaSuccess || max_clusters <= 0) {
      cudaGetLastError();
      max_clusters = 32;
    }
  }
  const int clusters = n_parts < max_clusters ? n_parts : max_clusters;
  cfg.gridDim = dim3(kScanCluster * clusters);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gru_scan_tc_kernel, *a, n_parts);
  if (e != cudaSuccess) return cuda_fail(e, "gru_scan_tc_kernel launch");
  return TEMP_OK;
}

int tc_pack_weights(const float* w_kn, int k, int n, void* packed, cudaStream_t st) {
  if (k != kD || (n & 127) != 0) return tcw_pack_weights(w_kn, k, n, packed, st);   // (for k == 128, n % 128 == 0 the images coincide)
  if (w_kn == nullptr || packed == nullptr || k != kD || n <= 0 || (n & 127) != 0)
    return fail(TEMP_EINVAL, "temp_pack_weights: k must be 128 and n a positive multiple of 128%s", "");
  pack_weights_kernel<<<(kD * n + 255) / 256, 256, 0, st>>>(w_kn, n, static_cast<uint8_t*>(packed));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack_weights_kernel launch");
  return TEMP_OK;
}

int tc_pack_gru_weights(const float* whh_t, int d, void* packed, cudaStream_t st) {
  if (d != kD) return tcw_pack_gru_weights(whh_t, d, packed, st);
  if (whh_t == nullptr || packed == nullptr) return fail(TEMP_EINVAL, "temp_pack_gru_weights: null pointers%s", "");
  pack_gru_kernel<<<(4 * kD * 128 + 255) / 256, 256, 0, st>>>(whh_t, static_cast<uint8_t*>(packed));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack_gru_kernel launch");
  return TEMP_OK;
}

}  // namespace temp_internal
