// temp_b200 -- tcgen05 RGCN layer for every width the 128-wide kernels of tc_kernels.cu do not take: embed_size ==
// hidden_size = d with d % 4 == 0, d <= 256 (BASELINE config 3 as the reference ships it: n_bases = 100 => d = 200,
// 2x2 relation blocks, models/RGCN.py:25-26), and d == 128 with 2x2 / 4x4 relation blocks.
//
// Same conventions as tc_kernels.cu ("features on TMEM lanes", 3xTF32 operand split, packed weight chunks of 128
// features x 32 k fetched by the TMA engine); what differs is the tiling:
//   * K = d is padded to KA = ceil(d / 32) k-atoms (zero columns), the output features to ceil(d / 128) blocks of 128
//     TMEM lanes (zero weight rows) -- pack_weights_wide_kernel writes the padded image;
//   * a tile is 64 packed rows (UMMA N = 64): hi + lo operand images of 64 x 32 KA floats (112 KB at d = 200) next to
//     the 3-stage weight ring (96 KB);
//   * the self-loop GEMM fills up to two 64-column accumulators (feature blocks), the chained GEMM walks
//     ceil(chain_n / 128) feature blocks through a ring of three.
//
//   rgcn_gather_wide_kernel<S> : the CSR-by-destination aggregation with S x S relation blocks (S = 1, 2, 4) applied in
//        registers -- a lane owns float4 channel groups lane and lane + 32, a float4 holds whole blocks for these S, so
//        the block product needs nothing from another lane.  Arithmetic (and summation order) of rgcn_layer_kernel's
//        aggregation in temp_kernels.cu, to which the parity tests hold it.
//   rgcn_layer_tcw_kernel      : one CTA per 64 packed rows; 8 worker warps (operand staging, both epilogues), a TMA
//        warp (weight chunks), an MMA warp -- the structure of rgcn_layer_tc_kernel.
//   gru_scan_tcw_kernel        : all GRU steps of a window in one cooperative launch; CTA (tile slot, block of 32 hidden
//        columns) keeps its r | z | n slice of W_hh in TENSOR MEMORY for the whole scan (TS-form tcgen05.mma, d <= 224),
//        the state goes through L2, steps are ordered by per-tile completion counters (or a grid barrier).
//   gru_step_tcw_kernel        : one GRU step per launch (224 < d <= 256, single steps of the models whose two layers
//        alternate), the weight slice streamed through the ring.
// Reference statements: models/RGCN.py:53-104 (layer), models/RRGCN.py:77-89 and torch.nn.GRU (step), as cited on the
// entry points in include/temp_b200.h.  Measurements (BASELINE config 3: 0.879 -> 0.264 ms) and bounds: DESIGN.md 3a.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <utility>

#include "internal.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWRows = 64;                       // packed rows per tile = UMMA N
constexpr int kWMaxAtoms = 8;                    // d <= 256
constexpr int kWStages = 3;
constexpr int kWWorkerWarps = 8;
constexpr int kWWorkers = kWWorkerWarps * 32;
constexpr int kWThreads = kWWorkers + 128;       // + control warpgroup (TMA warp 8, MMA warp 9, two idle warps)
constexpr int kWAtomBytes = kWRows * 128;        // 8 KB: one k-atom block of one operand image

__host__ __device__ constexpr int wide_smem_bytes(int ka) { return 2 * ka * kWAtomBytes + kWStages * kWChunkBytes + 1024; }

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

// ------------------------------------------------------------------------------------------------
// weight packing (once per parameter version)
// ------------------------------------------------------------------------------------------------
// w [k, n] row-major.  Chunk (mb, ka) at (mb * KA + ka) * 32 KB holds element (feature m of the 128-feature block mb,
// k of k-atom ka), hi image then lo image; features >= n and k >= the matrix's k are zero.
__global__ void pack_weights_wide_kernel(const float* __restrict__ w, int k_dim, int n, int KA, int n_mb,
                                         uint8_t* __restrict__ out) {
  const int total = n_mb * KA * 128 * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int chunk = idx >> 12, rem = idx & 4095;        // 128 features x 32 k per chunk
    const int kk = rem >> 7, m = rem & 127;               // consecutive threads: consecutive features (coalesced reads)
    const int mb = chunk / KA, ka = chunk - mb * KA;
    const int k = 32 * ka + kk, col = 128 * mb + m;
    const float v = (k < k_dim && col < n) ? w[static_cast<size_t>(k) * n + col] : 0.f;
    float hi, lo;
    split_tf32(v, hi, lo);
    uint8_t* dst = out + static_cast<size_t>(chunk) * kWChunkBytes;
    const uint32_t off = sw128_off(m, kk);
    *reinterpret_cast<float*>(dst + off) = hi;
    *reinterpret_cast<float*>(dst + 128 * 128 + off) = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// aggregation with S x S relation blocks
// ------------------------------------------------------------------------------------------------
// msg[b * S + j] = sum_i x[b * S + i] * W[rel][(b * S + i) * S + j]   (RGCN.py:91-98: bmm over the blocks), accumulated
// as m = fmaf(x_i, w_ij, m) from m = 0 in the order of i -- the statement of rgcn_layer_kernel (temp_kernels.cu).
// `w` points at the first weight of the float4's first block: S * 4 consecutive floats serve the 4 channels.
template <int S>
__device__ __forceinline__ float4 block_msg(const float4 x, const float* __restrict__ w) {
  float4 m;
  if (S == 1) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w));
    m.x = __fmul_rn(x.x, a.x);
    m.y = __fmul_rn(x.y, a.y);
    m.z = __fmul_rn(x.z, a.z);
    m.w = __fmul_rn(x.w, a.w);
  } else if (S == 2) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w));      // block 0: W00 W01 W10 W11
    const float4 b = __ldg(reinterpret_cast<const float4*>(w) + 1);  // block 1
    m.x = __fmaf_rn(x.y, a.z, __fmul_rn(x.x, a.x));
    m.y = __fmaf_rn(x.y, a.w, __fmul_rn(x.x, a.y));
    m.z = __fmaf_rn(x.w, b.z, __fmul_rn(x.z, b.x));
    m.w = __fmaf_rn(x.w, b.w, __fmul_rn(x.z, b.y));
  } else {
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(w));     // W0j
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(w) + 1);
    const float4 r2 = __ldg(reinterpret_cast<const float4*>(w) + 2);
    const float4 r3 = __ldg(reinterpret_cast<const float4*>(w) + 3);
    m.x = __fmaf_rn(x.w, r3.x, __fmaf_rn(x.z, r2.x, __fmaf_rn(x.y, r1.x, __fmul_rn(x.x, r0.x))));
    m.y = __fmaf_rn(x.w, r3.y, __fmaf_rn(x.z, r2.y, __fmaf_rn(x.y, r1.y, __fmul_rn(x.x, r0.y))));
    m.z = __fmaf_rn(x.w, r3.z, __fmaf_rn(x.z, r2.z, __fmaf_rn(x.y, r1.z, __fmul_rn(x.x, r0.z))));
    m.w = __fmaf_rn(x.w, r3.w, __fmaf_rn(x.z, r2.w, __fmaf_rn(x.y, r1.w, __fmul_rn(x.x, r0.w))));
  }
  return m;
}

constexpr int kGWWarps = 8;

// sum over the edges [e0, e1) of msg * nrm for this lane's channel groups lane (a) and lane + 32 (b), in edge order.
// The first chunk's indices are plan data and are fetched BEFORE pdl_wait(); the feature rows after it.
template <int S>
__device__ __forceinline__ void gather_edges_wide(const TempRgcnLayerArgs& p, int e0, int e1, float nrm, int lane, bool oka,
                                                  bool okb, float4& acc_a, float4& acc_b) {
  const int D = p.d;
  const size_t wrow = static_cast<size_t>(D) * S;        // floats per relation: n_bases * S * S
  acc_a = make_float4(0.f, 0.f, 0.f, 0.f);
  acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
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
      float4 xa[2], xb[2];
      int ru[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int su = __shfl_sync(kFull, s, (u0 + u) & 31);
        ru[u] = __shfl_sync(kFull, rl, (u0 + u) & 31);
        xa[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        xb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u0 + u < cnt) {
          const float* xr = p.x + static_cast<size_t>(su) * D;
          if (oka) xa[u] = ld_dep_f32x4(xr + 4 * lane);              // (the previous layer's output)
          if (okb) xb[u] = ld_dep_f32x4(xr + 4 * (lane + 32));
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u0 + u < cnt) {                                          // msg * norm_e, summed in edge order (RGCN.py:92-97)
          const float* wr = p.weight + static_cast<size_t>(ru[u]) * wrow;
          if (oka) {
            const float4 m = block_msg<S>(xa[u], wr + static_cast<size_t>(4 * S) * lane);
            acc_a.x += m.x * nrm; acc_a.y += m.y * nrm; acc_a.z += m.z * nrm; acc_a.w += m.w * nrm;
          }
          if (okb) {
            const float4 m = block_msg<S>(xb[u], wr + static_cast<size_t>(4 * S) * (lane + 32));
            acc_b.x += m.x * nrm; acc_b.y += m.y * nrm; acc_b.z += m.z * nrm; acc_b.w += m.w * nrm;
          }
        }
      }
    }
  }
}

template <int S>
__global__ void __launch_bounds__(kGWWarps * 32, 4) rgcn_gather_wide_kernel(const TempRgcnLayerArgs p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.d, nv = D >> 2;
  const bool oka = lane < nv, okb = lane + 32 < nv;
  pdl_launch_dependents();
  if (p.agg_lists != 0 && static_cast<int>(blockIdx.x) < p.n_agg_heavy) {
    // ---- a high in-degree row: the block's 8 warps sum contiguous edge chunks, partials added in chunk order ----
    __shared__ float4 part[kGWWarps][2][32];
    const int t = lane < 3 ? __ldg(p.agg_heavy + 3 * static_cast<size_t>(blockIdx.x) + lane) : 0;
    const int r = __shfl_sync(kFull, t, 0), p0 = __shfl_sync(kFull, t, 1), p1 = __shfl_sync(kFull, t, 2);
    const float nrm = __ldg(p.norm + r);
    const int chunk = (p1 - p0 + kGWWarps - 1) / kGWWarps;
    const int e0 = min(p0 + warp * chunk, p1), e1 = min(e0 + chunk, p1);
    float4 a, b;
    gather_edges_wide<S>(p, e0, e1, nrm, lane, oka, okb, a, b);
    part[warp][0][lane] = a;
    part[warp][1][lane] = b;
    __syncthreads();
    if (warp < 2) {
      float4 v = part[0][warp][lane];
#pragma unroll
      for (int w = 1; w < kGWWarps; ++w) {
        const float4 q = part[w][warp][lane];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      v.x *= nrm; v.y *= nrm; v.z *= nrm; v.w *= nrm;  // apply_func (RGCN.py:103-104)
      if (warp == 0 ? oka : okb)
        reinterpret_cast<float4*>(p.agg_scratch + static_cast<size_t>(r) * D)[lane + 32 * warp] = v;
    }
    return;
  }
  int r, p0, p1;
  if (p.agg_lists != 0) {  // compact work list: rows with in-edges only, CSR range inline
    const int item = (static_cast<int>(blockIdx.x) - p.n_agg_heavy) * kGWWarps + warp;
    if (item >= p.n_agg_rows) return;
    const int t = lane < 3 ? __ldg(p.agg_rows + 3 * static_cast<size_t>(item) + lane) : 0;
    r = __shfl_sync(kFull, t, 0);
    p0 = __shfl_sync(kFull, t, 1);
    p1 = __shfl_sync(kFull, t, 2);
  } else {
    r = p.row0 + blockIdx.x * kGWWarps + warp;
    if (r >= p.row1) return;
    p0 = __ldg(p.row_ptr + r);
    p1 = __ldg(p.row_ptr + r + 1);
    if (p1 <= p0) return;
  }
  const float nrm = __ldg(p.norm + r);
  float4 a, b;
  gather_edges_wide<S>(p, p0, p1, nrm, lane, oka, okb, a, b);
  float4* dst = reinterpret_cast<float4*>(p.agg_scratch + static_cast<size_t>(r) * D);
  if (oka) {
    a.x *= nrm; a.y *= nrm; a.z *= nrm; a.w *= nrm;  // apply_func (RGCN.py:103-104)
    dst[lane] = a;
  }
  if (okb) {
    b.x *= nrm; b.y *= nrm; b.z *= nrm; b.w *= nrm;
    dst[lane + 32] = b;
  }
}

// ------------------------------------------------------------------------------------------------
// fused RGCN layer, tcgen05, 64-row tiles
// ------------------------------------------------------------------------------------------------
// Development-only phase timeline (tools/probe_timeline_wide.py builds a separate library with -DTEMP_TIMELINE): lane 0 of
// every worker warp stores clock64() at mark m of step s into [cta][warp 0..7][step 0..15][mark 0..9].
#ifdef TEMP_TIMELINE
__device__ unsigned long long* g_timeline_w = nullptr;
__device__ __forceinline__ void tlw(int warp, int step, int mark) {
  if (g_timeline_w != nullptr && (threadIdx.x & 31) == 0 && warp < 8 && step < 16)
    g_timeline_w[((static_cast<size_t>(blockIdx.x) * 8 + warp) * 16 + step) * 10 + mark] = clock64();
}
#define TLW(m) tlw(warp, s, m)
#define TLL(m) tlw(warp, 0, m)
#else
#define TLW(m)
#define TLL(m)
#endif

struct WideBars {
  uint64_t w_full[kWStages], w_empty[kWStages];
  uint64_t b_ready, d1_full, x_ready;
  uint64_t d2_full[3], d2_empty[3];
  uint32_t tmem_base;
};

// TMEM columns: self-loop accumulators of feature blocks 0 / 1 at [0, 64) / [64, 128); chain ring slots at 128 + 64 s.
__global__ void __launch_bounds__(kWThreads, 1) rgcn_layer_tcw_kernel(const TempRgcnLayerArgs p, const int KA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + KA * kWAtomBytes;
  uint8_t* ring = smem + 2 * KA * kWAtomBytes;
  __shared__ WideBars S;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(kFull, tid >> 5, 0);       // provably warp-uniform (see rgcn_layer_tc_kernel)
  const int D = p.d;
  const int rbase = p.row0 + blockIdx.x * kWRows;
  const int n_fb = (D + 127) >> 7;
  // Chain feature blocks (128 output columns each) of THIS CTA: launches with few row tiles spread them over grid.y (every
  // y block rebuilds the tile's layer output, then takes blocks y, y + gridDim.y, ...; block y == 0 stores h_out).  The
  // blocks are independent of each other, so every CTA walks its list from a different start (rotated by the tile index):
  // CTAs that run in lock-step then stream DIFFERENT weight chunks instead of all hitting the same L2 lines at once.
  const int n_mb_all = p.chain_w_packed != nullptr ? (p.chain_n + 127) >> 7 : 0;
  const int gy = gridDim.y, by = blockIdx.y;
  const int n_mb = n_mb_all > by ? (n_mb_all - by + gy - 1) / gy : 0;
  const int rot = n_mb > 0 ? static_cast<int>(blockIdx.x % n_mb) : 0;
  auto chain_block = [&](int k) -> int { int kk = k + rot; if (kk >= n_mb) kk -= n_mb; return by + kk * gy; };
  const int fb_rot = static_cast<int>(blockIdx.x) & (n_fb - 1);      // (n_fb is 1 or 2) same idea for the self-loop blocks
  pdl_launch_dependents();

  if (tid == 0) {
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&S.w_full[i], 1);
      mbar_init(&S.w_empty[i], 1);
    }
    mbar_init(&S.b_ready, kWWorkers);
    mbar_init(&S.d1_full, 1);
    mbar_init(&S.x_ready, kWWorkers);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&S.d2_full[i], 1);
      mbar_init(&S.d2_empty[i], kWWorkers);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&S.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;

  if (warp_u >= kWWorkerWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == 8 && lane == 0) {
      // ===== TMA producer: the self-loop weight chunks, then the chain's, in the order the MMA warp consumes them =====
      const uint8_t* w1 = static_cast<const uint8_t*>(p.terms[0].w_packed);
      const uint8_t* wc = static_cast<const uint8_t*>(p.chain_w_packed);
      const int n1 = KA * n_fb, total = n1 + KA * n_mb;
      for (int i = 0; i < total; ++i) {
        const int st = i % kWStages;
        if (i >= kWStages) mbar_wait(&S.w_empty[st], ((i / kWStages) - 1) & 1);
        const uint8_t* src;
        if (i < n1) {
          const int fb = (i / KA) ^ fb_rot, ka = i % KA;
          src = w1 + static_cast<size_t>(fb * KA + ka) * kWChunkBytes;
        } else {
          const int k = (i - n1) / KA, ka = (i - n1) % KA;
          src = wc + static_cast<size_t>(chain_block(k) * KA + ka) * kWChunkBytes;
        }
        mbar_expect_tx(&S.w_full[st], kWChunkBytes);
        bulk_g2s(ring + st * kWChunkBytes, src, kWChunkBytes, &S.w_full[st]);
      }
    } else if (warp_u == 9) {
      // ===== MMA issuer: the whole warp walks the pipeline, the elected lane issues =====
      const bool leader = elect_one();
      const uint32_t tb = __shfl_sync(kFull, S.tmem_base, 0);
      const uint32_t idesc = umma_idesc_tf32(128, kWRows);
      const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo), rg = smem_u32(ring);
      int i = 0;
      mbar_wait(&S.b_ready, 0);
      tc_fence_after();
      for (int fb = 0; fb < n_fb; ++fb) {
#pragma unroll 1
        for (int ka = 0; ka < KA; ++ka, ++i) {
          const int st = i % kWStages;
          mbar_wait(&S.w_full[st], (i / kWStages) & 1);
          tc_fence_after();
          if (leader) {
            umma_katom_3x(tb + 64 * (fb ^ fb_rot), rg + st * kWChunkBytes, bh + ka * kWAtomBytes, bl + ka * kWAtomBytes, idesc, ka == 0);
            umma_commit(&S.w_empty[st]);
          }
          __syncwarp();
        }
      }
      if (leader) umma_commit(&S.d1_full);
      __syncwarp();
      if (n_mb > 0) {
        mbar_wait(&S.x_ready, 0);
        tc_fence_after();
        for (int mb = 0; mb < n_mb; ++mb) {
          const int slot = mb % 3;
          if (mb >= 3) {
            mbar_wait(&S.d2_empty[slot], ((mb / 3) - 1) & 1);
            tc_fence_after();
          }
#pragma unroll 1
          for (int ka = 0; ka < KA; ++ka, ++i) {
            const int st = i % kWStages;
            mbar_wait(&S.w_full[st], (i / kWStages) & 1);
            tc_fence_after();
            if (leader) {
              umma_katom_3x(tb + 128 + 64 * slot, rg + st * kWChunkBytes, bh + ka * kWAtomBytes, bl + ka * kWAtomBytes, idesc,
                            ka == 0);
              umma_commit(&S.w_empty[st]);
            }
            __syncwarp();
          }
          if (leader) umma_commit(&S.d2_full[slot]);
          __syncwarp();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ===== workers: warp (q, hf) owns TMEM lane quadrant q (features 128 fb + 32 q + lane) of tile rows [32 hf, 32 hf + 32) =====
    const int q = warp & 3, hf = warp >> 2;
    const int R0 = rbase + 32 * hf;
    const uint32_t s_hi = smem_u32(b_hi), s_lo = smem_u32(b_lo);
    const uint32_t lane_base = tbase + (static_cast<uint32_t>(32 * q) << 16);
    TLL(0);

    // ---- 1. self-loop operand rows -> shared memory (hi / lo): warp w stages tile rows 8 w .. 8 w + 7 (one 8-row swizzle
    // group), a lane per float4 channel groups lane and lane + 32; channels >= d and rows >= row1 are zero ----
    {
      const TempDenseTerm& tm = p.terms[0];
      const int rr = rbase + 8 * warp + (lane & 7);
      const int idx = rr < p.row1 ? (tm.a_index != nullptr ? __ldg(tm.a_index + rr) : rr) : -1;
      const int nv = D >> 2;
      const bool oka = lane < nv, okb = lane + 32 < nv;
      const bool ina = lane < 8 * KA, inb = lane + 32 < 8 * KA;      // inside the padded operand at all
      pdl_wait();  // everything above (barriers, TMEM, plan indices) overlapped the predecessor
      TLL(1);
      float4 va[8], vb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int sr = __shfl_sync(kFull, idx, i);
        va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sr >= 0) {
          const float* xr = tm.a + static_cast<size_t>(sr) * D;
          if (oka) va[i] = ld_dep_f32x4(xr + 4 * lane);
          if (okb) vb[i] = ld_dep_f32x4(xr + 4 * (lane + 32));
        }
      }
      // float4 group c4: k-atom c4 >> 3, 16-byte chunk (c4 & 7) ^ (row & 7); tile row 8 w + i: swizzle group w, row in group i
      const uint32_t off_a = static_cast<uint32_t>(lane >> 3) * kWAtomBytes + static_cast<uint32_t>(warp) * 1024u;
      const uint32_t off_b = off_a + 4u * kWAtomBytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t off = static_cast<uint32_t>(i) * 128u + (static_cast<uint32_t>((lane & 7) ^ i) << 4);
        float4 hi, lo;
        if (ina) {
          split_tf32(va[i].x, hi.x, lo.x);
          split_tf32(va[i].y, hi.y, lo.y);
          split_tf32(va[i].z, hi.z, lo.z);
          split_tf32(va[i].w, hi.w, lo.w);
          sts_f32x4(s_hi + off_a + off, hi);
          sts_f32x4(s_lo + off_a + off, lo);
        }
        if (inb) {
          split_tf32(vb[i].x, hi.x, lo.x);
          split_tf32(vb[i].y, hi.y, lo.y);
          split_tf32(vb[i].z, hi.z, lo.z);
          split_tf32(vb[i].w, hi.w, lo.w);
          sts_f32x4(s_hi + off_b + off, hi);
          sts_f32x4(s_lo + off_b + off, lo);
        }
      }
      fence_proxy_async();
      mbar_arrive(&S.b_ready);
    }
    TLL(2);

    // ---- 2. bit j of `has`: row R0 + j has in-edges (its aggregate was written by rgcn_gather_wide_kernel) ----
    unsigned has = 0u;
    if (p.row_ptr != nullptr) {
      const int ra = min(R0 + lane, p.row1);
      has = __ballot_sync(kFull, __ldg(p.row_ptr + min(ra + 1, p.row1)) > __ldg(p.row_ptr + ra));
    }
    const bool need_te = (p.te_out | p.te_chain) != 0;
    int rt = p.row_time_scalar;
    if (need_te && p.row_time != nullptr) rt = __ldg(p.row_time + min(R0 + lane, p.row1 - 1));

    // ---- 3. epilogue 1: out = act(agg (+x) + x . W_loop + bias) ; h_out ; chain operand X in place ----
    // Everything a row needs from the argument block is hoisted into registers / running pointers here: in the first
    // version the row loop re-read the flags from the constant bank, rebuilt 64-bit row addresses and carried the general
    // time-embedding walk in every unrolled iteration -- ~45 instructions and ~215 cycles per row (timeline probe: 13.9 k
    // cycles for the 64 rows of a warp with two feature blocks, the longest phase of a tile without a chain).
    const int rows_valid = max(0, min(32, p.row1 - R0));       // rows of this warp inside [row0, row1)
    const bool relu = p.activation == TEMP_ACT_RELU, residual = p.residual != 0, te_o = p.te_out != 0, te_c = p.te_chain != 0;
    const bool chain = n_mb > 0;
    const bool store_h = p.h_out != nullptr && by == 0;
    // the aggregate values of this thread's (feature, 32 rows) for both feature blocks: in flight while the MMAs complete
    float ag[2][32];
#pragma unroll
    for (int fb = 0; fb < 2; ++fb) {
      const int f = 128 * fb + 32 * q + lane;
      const float* ap = p.agg_scratch + static_cast<size_t>(R0) * D + f;
      const unsigned take = f < D ? has : 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        ag[fb][j] = ((take >> j) & 1u) ? ld_dep_f32(ap) : 0.f;
        ap += D;
      }
    }
    TLL(3);
    mbar_wait(&S.d1_full, 0);
    tc_fence_after();
    TLL(4);
    // time embedding: rows are packed snapshot by snapshot, so the 32 rows of this warp see one or two time-embedding rows
    // almost always -- both are fetched up front and a row picks one by its index (kFast); spans with three or more
    // snapshots (tiny graphs) take the general walk that reloads on change
    int t_first = 0, t_last = 0, te_split = 32;
    bool te_fast = true;
    if (need_te) {
      t_first = __shfl_sync(kFull, rt, 0);
      t_last = __shfl_sync(kFull, rt, 31);
      const unsigned is_first = __ballot_sync(kFull, rt == t_first);
      te_split = __popc(is_first);
      te_fast = __all_sync(kFull, rt == t_first || rt == t_last) && is_first == (te_split >= 32 ? 0xffffffffu : ((1u << te_split) - 1u));
    }
    auto epilogue1 = [&](auto fast_tag, const int fb, const float (&agv)[32]) {
      constexpr bool kFast = decltype(fast_tag)::value;
      const int atom = 4 * fb + q;
      const int f = 128 * fb + 32 * q + lane;
      const bool fok = f < D;
      const float bias = (fok && p.h_bias != nullptr) ? __ldg(p.h_bias + f) : 0.f;
      float te_a = 0.f, te_b = 0.f;
      if (need_te && kFast && fok) {
        te_a = __ldg(p.time_embed + static_cast<size_t>(t_first) * D + f);
        te_b = __ldg(p.time_embed + static_cast<size_t>(t_last) * D + f);
      }
      float v[32];
      tmem_ld32(lane_base + 64 * fb + 32 * hf, v);
      // operand address of (tile row 32 hf + i, k = f): k-atom `atom`, swizzle group 4 hf + (i >> 3), 16-byte chunk (lane >> 2) ^ (i & 7)
      const uint32_t rb = static_cast<uint32_t>(atom) * kWAtomBytes + static_cast<uint32_t>(4 * hf) * 1024u + (static_cast<uint32_t>(lane & 3) << 2);
      const uint32_t o_hi = s_hi + rb, o_lo = s_lo + rb, cx = static_cast<uint32_t>(lane) >> 2;
      // (the empty asm statements make these values opaque: without them the compiler re-derives the 64-bit row address and
      // re-reads the flags from the constant bank in every unrolled iteration rather than keeping them in registers)
      float* hp = p.h_out + static_cast<size_t>(R0) * D + f;
      size_t row_pitch = static_cast<size_t>(D) * sizeof(float);
      asm volatile("" : "+l"(hp), "+l"(row_pitch));
      int n_store = (store_h && fok) ? rows_valid : 0;           // rows of this thread that reach h_out
      int n_keep = fok ? rows_valid : 0;                         // rows of this thread that are not operand padding
      unsigned fl = (relu ? 1u : 0u) | (residual ? 2u : 0u) | (te_o ? 4u : 0u) | (te_c ? 8u : 0u) | (chain ? 16u : 0u);
      asm volatile("" : "+r"(n_store), "+r"(n_keep), "+r"(fl));
      const bool f_relu = fl & 1u, f_res = fl & 2u, f_teo = fl & 4u, f_tec = fl & 8u, f_chain = fl & 16u;
      int cur_trow = -1;
      float cur_te = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint32_t off = static_cast<uint32_t>(i >> 3) * 1024u + static_cast<uint32_t>(i & 7) * 128u + ((cx ^ static_cast<uint32_t>(i & 7)) << 4);
        float te = 0.f;
        if (need_te) {
          if (kFast) {
            te = i < te_split ? te_a : te_b;
          } else {
            const int trow = __shfl_sync(kFull, rt, i);  // warp-uniform
            if (trow != cur_trow) {
              cur_trow = trow;
              cur_te = fok ? __ldg(p.time_embed + static_cast<size_t>(trow) * D + f) : 0.f;
            }
            te = cur_te;
          }
        }
        float val = agv[i];
        if (f_res) val += lds_f32(o_hi + off) + lds_f32(o_lo + off);
        val += v[i];
        val += bias;
        if (f_relu) val = fmaxf(val, 0.f);
        const float with_te = val + te;
        if (i < n_store) *hp = f_teo ? with_te : val;
        hp = reinterpret_cast<float*>(reinterpret_cast<char*>(hp) + row_pitch);
        if (f_chain) {
          const float xx = i < n_keep ? (f_tec ? with_te : val) : 0.f;
          float hi, lo;
          split_tf32(xx, hi, lo);
          sts_f32(o_hi + off, hi);
          sts_f32(o_lo + off, lo);
        }
      }
    };
#pragma unroll
    for (int fb = 0; fb < 2; ++fb) {
      if (fb >= n_fb || 4 * fb + q >= KA) continue;          // warp-uniform: no feature of this warp in the block
      if (te_fast)
        epilogue1(std::true_type{}, fb, ag[fb]);
      else
        epilogue1(std::false_type{}, fb, ag[fb]);
    }

    TLL(5);
    // ---- 4. chain epilogues: chain_out[r, 128 mb + 32 q + lane] = D2 + chain_b ----
    if (n_mb > 0) {
      fence_proxy_async();
      mbar_arrive(&S.x_ready);
      const size_t chain_ld = static_cast<size_t>(p.chain_ld);
      for (int mb = 0; mb < n_mb; ++mb) {
        const int slot = mb % 3;
        const int cf = 128 * chain_block(mb) + 32 * q + lane;
        const bool cok = cf < p.chain_n;
        const float cbias = (cok && p.chain_b != nullptr) ? __ldg(p.chain_b + cf) : 0.f;
        mbar_wait(&S.d2_full[slot], (mb / 3) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(lane_base + 128 + 64 * slot + 32 * hf, v);
        float* orow = p.chain_out + static_cast<size_t>(R0) * chain_ld + cf;
        const int n_out = cok ? rows_valid : 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < n_out) *orow = v[i] + cbias;
          orow += chain_ld;
        }
        tc_fence_before();
        mbar_arrive(&S.d2_empty[slot]);
      }
    }
  }

  TLL(6);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tbase, 512);
}


// ------------------------------------------------------------------------------------------------
// GRU step, tcgen05, 64-row tiles x 32 hidden columns
// ------------------------------------------------------------------------------------------------
// whh_t [d, 3 d] row-major (= weight_hh transposed).  Block cb of 32 hidden columns: KA chunks of 32 KB at (cb * KA + ka);
// feature row m = 32 * gate + jj  <->  column gate * d + 32 * cb + jj (the layout of pack_gru_kernel: the r | z | n
// pre-activations of one hidden column sit on the same lane of TMEM lane quadrants 0, 1, 2); rows 96..127, hidden columns
// >= d and k >= d are zero.
__global__ void pack_gru_wide_kernel(const float* __restrict__ whh_t, int d, int KA, int CB, uint8_t* __restrict__ out) {
  const int total = CB * KA * 4096;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int chunk = idx >> 12, rem = idx & 4095;
    const int kk = rem >> 7, m = rem & 127;
    const int cb = chunk / KA, ka = chunk - cb * KA;
    const int k = 32 * ka + kk, gate = m >> 5, j = 32 * cb + (m & 31);
    const float v = (gate < 3 && k < d && j < d) ? whh_t[static_cast<size_t>(k) * (3 * d) + gate * d + j] : 0.f;
    float hi, lo;
    split_tf32(v, hi, lo);
    uint8_t* dst = out + static_cast<size_t>(chunk) * kWChunkBytes;
    const uint32_t off = sw128_off(m, kk);
    *reinterpret_cast<float*>(dst + off) = hi;
    *reinterpret_cast<float*>(dst + 128 * 128 + off) = lo;
  }
}

struct GruWBars {
  uint64_t w_full[kWStages], w_empty[kWStages];
  uint64_t b_ready, d_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float decay_factor_w(float dt, const float* wb, float inv_temperature) {
  if (wb != nullptr) return expf(-fmaxf(fmaf(__ldg(wb), dt, __ldg(wb + 1)), 0.f));
  return expf(-dt * inv_temperature);
}
// ex2.approx / rcp.approx based gate math, absolute error ~1e-7 (as the 128-wide scans; the parity bar is 1e-4 relative).
// The gate phase is on the serial chain of every step: with expf / tanhf it was 4.4 k of a step's 13 k cycles
// (tools/probe_timeline_wide.py).
__device__ __forceinline__ float sigmoid_w(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_w(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// One GRU step (models/RRGCN.py:77-89, torch.nn.GRU equations) for packed rows [row0, row1): CTA (tile, cb) = 64 rows x the
// 32 hidden columns [32 cb, 32 cb + 32).
//   h0 = state[prev_row[r]] (zero row when prev_row[r] < 0)  ->  hi / lo operand in shared memory;
//   gh^T[r|z|n of the 32 columns, row] = W_hh slice . h0^T on the tensor pipe (KA k-atoms x 12 MMAs, weight chunks streamed
//   by the TMA warp); the decay is a per-row scalar and is applied after the MMA (W . (c h) = c (W . h));
//   the three gate quadrants of the accumulator are parked in shared memory, then thread (warp w, lane) computes the gates
//   of rows 8 w .. 8 w + 7 at hidden column 32 cb + lane (gi / out accesses: 128-byte lines) and stores state (+ te).
// Steps of a window are separate launches chained by programmatic dependent launch: barriers, TMEM, the first weight
// chunks and the plan's prev_row are set up while the previous step still runs.
__global__ void __launch_bounds__(kWThreads, 1) gru_step_tcw_kernel(const TempGruArgs p, const int KA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + KA * kWAtomBytes;
  uint8_t* ring = smem + 2 * KA * kWAtomBytes;
  __shared__ GruWBars S;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(kFull, tid >> 5, 0);
  const int D = p.d;
  const int rbase = p.row0 + blockIdx.x * kWRows;
  const int cb = blockIdx.y;
  const bool rec = p.prev_row != nullptr;          // (launch-uniform) false: no row of the step has a previous state
  pdl_launch_dependents();

  if (tid == 0) {
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&S.w_full[i], 1);
      mbar_init(&S.w_empty[i], 1);
    }
    mbar_init(&S.b_ready, kWWorkers);
    mbar_init(&S.d_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&S.tmem_base, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;

  if (warp_u >= kWWorkerWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (rec && warp == 8 && lane == 0) {
      const uint8_t* w = static_cast<const uint8_t*>(p.whh_packed) + static_cast<size_t>(cb) * KA * kWChunkBytes;
      for (int i = 0; i < KA; ++i) {
        const int st = i % kWStages;
        if (i >= kWStages) mbar_wait(&S.w_empty[st], ((i / kWStages) - 1) & 1);
        mbar_expect_tx(&S.w_full[st], kWChunkBytes);
        bulk_g2s(ring + st * kWChunkBytes, w + static_cast<size_t>(i) * kWChunkBytes, kWChunkBytes, &S.w_full[st]);
      }
    } else if (rec && warp_u == 9) {
      const bool leader = elect_one();
      const uint32_t tb = __shfl_sync(kFull, S.tmem_base, 0);
      const uint32_t idesc = umma_idesc_tf32(128, kWRows);
      const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo), rg = smem_u32(ring);
      mbar_wait(&S.b_ready, 0);
      tc_fence_after();
#pragma unroll 1
      for (int ka = 0; ka < KA; ++ka) {
        const int st = ka % kWStages;
        mbar_wait(&S.w_full[st], (ka / kWStages) & 1);
        tc_fence_after();
        if (leader) {
          umma_katom_3x(tb, rg + st * kWChunkBytes, bh + ka * kWAtomBytes, bl + ka * kWAtomBytes, idesc, ka == 0);
          umma_commit(&S.w_empty[st]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(&S.d_full);
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t s_hi = smem_u32(b_hi), s_lo = smem_u32(b_lo);
    // lanes 0..7: previous-state row and time gap of tile row 8 w + lane (plan data: fetched before the dependency point)
    const int rr = rbase + 8 * warp + (lane & 7);
    int pr = -1;
    float dt = 1.f;          // decay factor of tile row 8 w + (lane & 7)
    if (rec && rr < p.row1) {
      pr = __ldg(p.prev_row + rr);
      if (p.dt != nullptr) dt = decay_factor_w(__ldg(p.dt + rr), p.decay_wb, p.inv_temperature);   // (the factor, once per row)
    }
    const int j = 32 * cb + lane;
    const bool jok = j < D;
    const float br = jok ? __ldg(p.b_hh + j) : 0.f, bz = jok ? __ldg(p.b_hh + D + j) : 0.f, bn = jok ? __ldg(p.b_hh + 2 * D + j) : 0.f;
    int trow = p.row_time_scalar;
    if (p.time_embed != nullptr && p.row_time != nullptr) trow = __ldg(p.row_time + min(rr, p.row1 - 1));
    pdl_wait();   // gi, the previous state and (accumulate) out come from the predecessor kernels

    // ---- 1. previous-state rows -> hi / lo operand (warp w: tile rows 8 w .. 8 w + 7, lanes = float4 groups lane, lane + 32) ----
    if (rec) {
      const int nv = D >> 2;
      const bool oka = lane < nv, okb = lane + 32 < nv;
      const bool ina = lane < 8 * KA, inb = lane + 32 < 8 * KA;
      float4 va[8], vb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int sr = __shfl_sync(kFull, pr, i);
        va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sr >= 0) {
          const float* xr = p.state + static_cast<size_t>(sr) * D;
          if (oka) va[i] = ld_dep_f32x4(xr + 4 * lane);
          if (okb) vb[i] = ld_dep_f32x4(xr + 4 * (lane + 32));
        }
      }
      const uint32_t off_a = static_cast<uint32_t>(lane >> 3) * kWAtomBytes + static_cast<uint32_t>(warp) * 1024u;
      const uint32_t off_b = off_a + 4u * kWAtomBytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t off = static_cast<uint32_t>(i) * 128u + (static_cast<uint32_t>((lane & 7) ^ i) << 4);
        float4 hi, lo;
        if (ina) {
          split_tf32(va[i].x, hi.x, lo.x);
          split_tf32(va[i].y, hi.y, lo.y);
          split_tf32(va[i].z, hi.z, lo.z);
          split_tf32(va[i].w, hi.w, lo.w);
          sts_f32x4(s_hi + off_a + off, hi);
          sts_f32x4(s_lo + off_a + off, lo);
        }
        if (inb) {
          split_tf32(vb[i].x, hi.x, lo.x);
          split_tf32(vb[i].y, hi.y, lo.y);
          split_tf32(vb[i].z, hi.z, lo.z);
          split_tf32(vb[i].w, hi.w, lo.w);
          sts_f32x4(s_hi + off_b + off, hi);
          sts_f32x4(s_lo + off_b + off, lo);
        }
      }
      fence_proxy_async();
      mbar_arrive(&S.b_ready);
    }

    // ---- 2. what the gates need besides gh: input gates, time embedding, accumulate target: in flight during the MMAs ----
    // (running pointers and a row count kept opaque: the unrolled row loops otherwise rebuild 64-bit addresses and re-read the
    // argument block from the constant bank per row -- the layer kernel's epilogue showed ~215 cycles per row that way)
    const float* gp = p.gi + static_cast<size_t>(rbase + 8 * warp) * p.gi_ld + p.gi_off + j;
    float* op = p.out + static_cast<size_t>(rbase + 8 * warp) * D + j;
    size_t gi_pitch = static_cast<size_t>(p.gi_ld) * sizeof(float), out_pitch = static_cast<size_t>(D) * sizeof(float);
    int n_rows = jok ? max(0, min(8, p.row1 - (rbase + 8 * warp))) : 0;      // rows of this thread inside [row0, row1)
    unsigned gfl = (p.time_embed != nullptr ? 1u : 0u) | (p.accumulate ? 2u : 0u);
    asm volatile("" : "+l"(gp), "+l"(op), "+l"(gi_pitch), "+l"(out_pitch), "+r"(n_rows), "+r"(gfl));
    const bool has_te = gfl & 1u, accum = gfl & 2u;
    float gir[8], giz[8], gin[8], tev[8], old[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      gir[i] = giz[i] = gin[i] = tev[i] = old[i] = 0.f;
      const int tr_i = __shfl_sync(kFull, trow, i);          // (lane i holds tile row 8 w + i)
      if (i < n_rows) {
        gir[i] = ld_dep_f32(gp);
        giz[i] = ld_dep_f32(gp + D);
        gin[i] = ld_dep_f32(gp + 2 * D);
        if (has_te) tev[i] = __ldg(p.time_embed + static_cast<size_t>(tr_i) * D + j);
        if (accum) old[i] = ld_dep_f32(reinterpret_cast<const char*>(op) + i * out_pitch);
      }
      gp = reinterpret_cast<const float*>(reinterpret_cast<const char*>(gp) + gi_pitch);
    }

    // ---- 3. accumulator quadrants r | z | n -> exchange rows [gate][tile row][32] in the (now idle) weight ring ----
    const uint32_t ex = smem_u32(ring);
    if (rec) {
      mbar_wait(&S.d_full, 0);
      tc_fence_after();
      if (q < 3) {
        float v[32];
        tmem_ld32(tbase + (static_cast<uint32_t>(32 * q) << 16) + 32 * hf, v);
        const uint32_t exw = ex + static_cast<uint32_t>(q * kWRows + 32 * hf) * 128u + lane * 4u;
#pragma unroll
        for (int i = 0; i < 32; ++i) sts_f32(exw + i * 128u, v[i]);
      }
      tc_fence_before();
      bar_named(1, kWWorkers);
    }

    // ---- 4. gates (torch.nn.GRU, gate order r, z, n: SURVEY Appendix A.3), state store ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = 8 * warp + i;
      const int pri = __shfl_sync(kFull, pr, i);
      const float dti = __shfl_sync(kFull, dt, i);
      if (i >= n_rows) continue;
      float hr = br, hz = bz, hn = bn, hp = 0.f;
      if (pri >= 0) {
        const float dec = dti;
        const uint32_t ea = ex + static_cast<uint32_t>(m) * 128u + lane * 4u;
        hr = fmaf(dec, lds_f32(ea), hr);
        hz = fmaf(dec, lds_f32(ea + kWRows * 128u), hz);
        hn = fmaf(dec, lds_f32(ea + 2u * kWRows * 128u), hn);
        const uint32_t off = static_cast<uint32_t>(cb) * kWAtomBytes + sw128_off(static_cast<uint32_t>(m), static_cast<uint32_t>(lane));
        hp = dec * (lds_f32(s_hi + off) + lds_f32(s_lo + off));     // hi + lo is the fp32 value exactly
      }
      const float rg = sigmoid_w(gir[i] + hr), zg = sigmoid_w(giz[i] + hz);
      const float ng = tanh_w(gin[i] + rg * hn);
      float hy = (1.f - zg) * ng + zg * hp;
      hy += tev[i];
      *reinterpret_cast<float*>(reinterpret_cast<char*>(op) + i * out_pitch) = accum ? old[i] + hy : hy;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tbase, 64);
}


// ------------------------------------------------------------------------------------------------
// all GRU steps of a window in ONE cooperative launch, W_hh in tensor memory (d <= 224)
// ------------------------------------------------------------------------------------------------
// The per-step kernel above re-streams its 32 KB x KA weight slice from L2 for every step and pays a launch hand-over per
// step.  Here CTA (slot, cb) keeps the r | z | n rows of hidden-column block cb in TENSOR MEMORY for the whole scan -- hi
// parts at columns [0, 32 KA), lo parts at [32 KA, 64 KA), the 64-column accumulator behind them (64 KA + 64 <= 512 columns:
// KA <= 7) -- and the MMAs take their A operand from there (tcgen05.mma "TS" form: an MMA then reads only its 64 x 8
// activation slice from shared memory).  Steps are separated by a grid-wide barrier (cooperative launch: all CTAs resident);
// the state goes through L2 (ld.global.cg: rows of consecutive steps share 128-byte lines, L1 may hold a stale copy).
// The slice is (re)installed when the recurrent cell changes (the Bi models: forward cell steps, then backward cell steps).
struct ScanWBars {
  uint64_t w_full[kWStages], w_empty[kWStages];
  uint64_t b_ready, d_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ld_cg_f32x4(const void* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cg_f32(const void* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f32x4_w(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void cta_sync_w() { asm volatile("bar.sync 0, %0;" ::"n"(kWThreads) : "memory"); }

// Both role loops below walk the same (step, tile) sequence and meet at the same CTA-wide barriers (barrier 0, all 384
// threads: end of a weight install, the two halves of the grid barrier, end of a tile step); the roles are separated at the
// top so that the register split (setmaxnreg 208 / 88) holds for each loop.
// tile_stride > 0: per-tile completion counters instead of the grid barrier -- word 2 + s * tile_stride + tile of P.barrier
// counts the column-block CTAs that have stored (step s, tile); a worker warp waits, before it gathers its 8 rows' previous
// states, only for the tiles of the producing step that hold those rows (dependencies point to earlier steps only and every
// CTA walks its tile steps in order, so the waits cannot form a cycle; all CTAs are resident: cooperative launch).
__global__ void __launch_bounds__(kWThreads, 1) gru_scan_tcw_kernel(const TempGruScanArgs P, const int KA, const int T,
                                                                     const int tile_stride) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + KA * kWAtomBytes;
  uint8_t* ring = smem + 2 * KA * kWAtomBytes;
  __shared__ ScanWBars S;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(kFull, tid >> 5, 0);
  const int D = P.steps[0].d;
  const int slot = blockIdx.x / KA, cb = blockIdx.x - slot * KA;
  const int col_lo = 32 * KA, col_d = 64 * KA;     // TMEM columns: W lo parts, accumulator

  if (tid == 0) {
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&S.w_full[i], 1);
      mbar_init(&S.w_empty[i], kWWorkerWarps);
    }
    mbar_init(&S.b_ready, kWWorkers);
    mbar_init(&S.d_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&S.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;
  const uint32_t s_hi = smem_u32(b_hi), s_lo = smem_u32(b_lo), ex = smem_u32(ring);
  pdl_wait();   // (a no-op for the cooperative launch; gi comes from the preceding kernels in stream order)

  if (warp_u >= kWWorkerWarps) {
    // =============================== control warps: weight copies, MMA issue ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const void* cur_w = nullptr;
    uint32_t wchunk = 0, mm = 0;
#pragma unroll 1
    for (int s = 0; s < P.n_steps; ++s) {
      const TempGruArgs& p = P.steps[s];
      const bool rec = p.prev_row != nullptr;
      if (rec && p.whh_packed != cur_w) {
        cur_w = p.whh_packed;
        if (warp == 8 && lane == 0) {
          const uint8_t* w = static_cast<const uint8_t*>(p.whh_packed) + static_cast<size_t>(cb) * KA * kWChunkBytes;
          for (int i = 0; i < KA; ++i) {
            const uint32_t c = wchunk + i, st = c % kWStages;
            if (c >= kWStages) mbar_wait(&S.w_empty[st], ((c / kWStages) - 1) & 1);
            mbar_expect_tx(&S.w_full[st], kWChunkBytes);
            bulk_g2s(ring + st * kWChunkBytes, w + static_cast<size_t>(i) * kWChunkBytes, kWChunkBytes, &S.w_full[st]);
          }
        }
        wchunk += KA;
        tc_fence_before();
        cta_sync_w();
        tc_fence_after();
      }
      const int ntiles = (p.row1 - p.row0 + kWRows - 1) / kWRows;
      bool need_bar = s > 0 && tile_stride == 0;
#pragma unroll 1
      for (int tile = slot; tile < ntiles || need_bar; tile += T) {
        if (need_bar) {       // the grid barrier (its global part is thread 0's, a worker)
          cta_sync_w();
          cta_sync_w();
          need_bar = false;
        }
        if (tile >= ntiles) break;
        if (rec && warp_u == 9) {
          // ===== MMA issuer: gh^T = W_hh slice (tensor memory) . h0^T =====
          const bool leader = elect_one();
          const uint32_t idesc = umma_idesc_tf32(128, kWRows);
          mbar_wait(&S.b_ready, mm & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t dcol = tbase + col_d;
#pragma unroll 1
            for (int ka = 0; ka < KA; ++ka) {
              const uint32_t bh0 = umma_desc_lo(s_hi + ka * kWAtomBytes), bl0 = umma_desc_lo(s_lo + ka * kWAtomBytes);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t a_hi = tbase + ka * 32 + ks * 8, a_lo = a_hi + col_lo;
                const uint32_t bh = bh0 + 2 * ks, bl = bl0 + 2 * ks;          // +32 bytes per k-step inside the atom
                umma_tf32_ts(dcol, a_lo, bh, idesc, (ka | ks) != 0 ? 1u : 0u);  // small terms first
                umma_tf32_ts(dcol, a_hi, bl, idesc, 1u);
                umma_tf32_ts(dcol, a_hi, bh, idesc, 1u);
              }
            }
            umma_commit(&S.d_full);
          }
          __syncwarp();
        }
        if (rec) mm += 1;
        cta_sync_w();
        // (completion counters) every worker's state stores are ordered before this release by the barrier; an otherwise
        // idle warp signals, so that no worker warp sits behind the release's memory barrier at the top of its next step
        if (tile_stride > 0 && warp == 10 && lane == 0)
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.barrier + 2 + static_cast<size_t>(s) * tile_stride + tile) : "memory");
      }
    }
  } else {
    // =============================== workers ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int q = warp & 3, hf = warp >> 2;
    const int j = 32 * cb + lane;
    const bool jok = j < D;
    const void* cur_w = nullptr;
    uint32_t wchunk = 0, mm = 0;
    unsigned n_bar = 0;
#pragma unroll 1
    for (int s = 0; s < P.n_steps; ++s) {
      const TempGruArgs& p = P.steps[s];
      const bool rec = p.prev_row != nullptr;
      if (rec && p.whh_packed != cur_w) {
        // ---- this CTA's W_hh slice -> tensor memory: KA chunks through the ring; warp (q, img = hf) moves feature rows
        // 32 q .. + 31 of the hi (img 0) / lo (img 1) image of each chunk ----
        cur_w = p.whh_packed;
        const int m = 32 * q + lane;
#pragma unroll 1
        for (int i = 0; i < KA; ++i) {
          const uint32_t c = wchunk + i, st = c % kWStages;
          mbar_wait(&S.w_full[st], (c / kWStages) & 1);
          const uint32_t rowp = ex + st * kWChunkBytes + hf * (128 * 128) + (m >> 3) * 1024 + (m & 7) * 128;
          float v[32];
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 x = lds_f32x4_w(rowp + ((c4 ^ (m & 7)) << 4));
            v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w;
          }
          tmem_st32(tbase + (static_cast<uint32_t>(32 * q) << 16) + hf * col_lo + 32 * i, v);
          tmem_st_wait();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.w_empty[st]);
        }
        wchunk += KA;
        tc_fence_before();
        cta_sync_w();   // every warp's part is in TMEM (and the ring is idle again: it doubles as the gate exchange buffer)
        tc_fence_after();
      }

      const int ntiles = (p.row1 - p.row0 + kWRows - 1) / kWRows;
      bool need_bar = s > 0 && tile_stride == 0;
#pragma unroll 1
      for (int tile = slot; tile < ntiles || need_bar; tile += T) {
        const bool has = tile < ntiles;
        const int rbase = p.row0 + tile * kWRows;
        // plan data and parameters of the tile: fetched BEFORE the step's dependency point
        const int rr = rbase + 8 * warp + (lane & 7);
        int pr = -1, trow = p.row_time_scalar;
        float dt = 1.f, br = 0.f, bz = 0.f, bn = 0.f;     // dt: the decay FACTOR of tile row 8 w + (lane & 7)
        if (has) {
          if (rec && rr < p.row1) {
            pr = __ldg(p.prev_row + rr);
            if (p.dt != nullptr) dt = decay_factor_w(__ldg(p.dt + rr), p.decay_wb, p.inv_temperature);
          }
          if (jok) {
            br = __ldg(p.b_hh + j);
            bz = __ldg(p.b_hh + D + j);
            bn = __ldg(p.b_hh + 2 * D + j);
          }
          if (p.time_embed != nullptr && p.row_time != nullptr) trow = __ldg(p.row_time + min(rr, p.row1 - 1));
        }
        if (need_bar) {       // grid-wide: every CTA's state stores of the previous step are visible
          const unsigned target = (++n_bar) * gridDim.x;
          cta_sync_w();     // every thread's state stores are ordered before thread 0's release (cumulativity)
          if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.barrier) : "memory");
            unsigned v, spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(P.barrier) : "memory");
              if (++spins > (1u << 26)) __trap();
            } while (static_cast<int>(v - target) < 0);
          }
          cta_sync_w();
          need_bar = false;
        }
        if (!has) break;

        TLW(0);
        // ---- 0. (completion counters) this warp's 8 rows: wait for the tiles of the producing step that hold their states ----
        if (rec && tile_stride > 0) {
          int mn = (lane < 8 && pr >= 0) ? pr : 0x7fffffff, mx = lane < 8 ? pr : -1;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(kFull, mn, o));
            mx = max(mx, __shfl_xor_sync(kFull, mx, o));
          }
          mn = __shfl_sync(kFull, mn, 0);
          mx = __shfl_sync(kFull, mx, 0);
          if (mx >= 0) {       // warp-uniform
            // the producing step of a row: the latest earlier step of this launch that writes it into the buffer this step reads
            int sa = -1, sb = -1;
            for (int k = s - 1; k >= 0 && (sa < 0 || sb < 0); --k) {
              const TempGruArgs& g = P.steps[k];
              if (g.out != p.state) continue;
              if (sa < 0 && mn >= g.row0 && mn < g.row1) sa = k;
              if (sb < 0 && mx >= g.row0 && mx < g.row1) sb = k;
            }
            const unsigned* flags = P.barrier + 2;
            auto wait_tiles = [&](int k, int lo_row, int hi_row) {   // tiles of step k covering rows [lo_row, hi_row]
              const int r0 = P.steps[k].row0;
              const int t0 = (lo_row - r0) / kWRows, t1 = (hi_row - r0) / kWRows;
              for (int t = t0 + lane; t <= t1; t += 32) {
                const unsigned* f = flags + static_cast<size_t>(k) * tile_stride + t;
                unsigned v, spins = 0;
                do {
                  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(f) : "memory");
                  if (++spins > (1u << 26)) __trap();
                } while (v < static_cast<unsigned>(KA));
              }
            };
            if (sa >= 0 && sa == sb) {
              wait_tiles(sa, mn, mx);
            } else {             // rows of several producing steps (no planner emits this): every tile of each of them
              if (sa >= 0) wait_tiles(sa, P.steps[sa].row0, P.steps[sa].row1 - 1);
              if (sb >= 0) wait_tiles(sb, P.steps[sb].row0, P.steps[sb].row1 - 1);
            }
            __syncwarp();
          }
        }

        TLW(1);
        // ---- 1. previous-state rows -> hi / lo operand ----
        if (rec) {
          const int nv = D >> 2;
          const bool oka = lane < nv, okb = lane + 32 < nv;
          const bool ina = lane < 8 * KA, inb = lane + 32 < 8 * KA;
          float4 va[8], vb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int sr = __shfl_sync(kFull, pr, i);
            va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (sr >= 0) {
              const float* xr = p.state + static_cast<size_t>(sr) * D;
              if (oka) va[i] = ld_cg_f32x4(xr + 4 * lane);
              if (okb) vb[i] = ld_cg_f32x4(xr + 4 * (lane + 32));
            }
          }
          const uint32_t off_a = static_cast<uint32_t>(lane >> 3) * kWAtomBytes + static_cast<uint32_t>(warp) * 1024u;
          const uint32_t off_b = off_a + 4u * kWAtomBytes;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t off = static_cast<uint32_t>(i) * 128u + (static_cast<uint32_t>((lane & 7) ^ i) << 4);
            float4 hi, lo;
            if (ina) {
              split_tf32(va[i].x, hi.x, lo.x);
              split_tf32(va[i].y, hi.y, lo.y);
              split_tf32(va[i].z, hi.z, lo.z);
              split_tf32(va[i].w, hi.w, lo.w);
              sts_f32x4(s_hi + off_a + off, hi);
              sts_f32x4(s_lo + off_a + off, lo);
            }
            if (inb) {
              split_tf32(vb[i].x, hi.x, lo.x);
              split_tf32(vb[i].y, hi.y, lo.y);
              split_tf32(vb[i].z, hi.z, lo.z);
              split_tf32(vb[i].w, hi.w, lo.w);
              sts_f32x4(s_hi + off_b + off, hi);
              sts_f32x4(s_lo + off_b + off, lo);
            }
          }
          fence_proxy_async();
          mbar_arrive(&S.b_ready);
        }
        TLW(2);
        // ---- 2. input gates, time embedding, accumulate target: in flight during the MMAs ----
        // (running pointers and a row count kept opaque: the unrolled row loops otherwise rebuild 64-bit addresses and re-read the
        // argument block from the constant bank per row -- the layer kernel's epilogue showed ~215 cycles per row that way)
        const float* gp = p.gi + static_cast<size_t>(rbase + 8 * warp) * p.gi_ld + p.gi_off + j;
        float* op = p.out + static_cast<size_t>(rbase + 8 * warp) * D + j;
        size_t gi_pitch = static_cast<size_t>(p.gi_ld) * sizeof(float), out_pitch = static_cast<size_t>(D) * sizeof(float);
        int n_rows = jok ? max(0, min(8, p.row1 - (rbase + 8 * warp))) : 0;      // rows of this thread inside [row0, row1)
        unsigned gfl = (p.time_embed != nullptr ? 1u : 0u) | (p.accumulate ? 2u : 0u);
        asm volatile("" : "+l"(gp), "+l"(op), "+l"(gi_pitch), "+l"(out_pitch), "+r"(n_rows), "+r"(gfl));
        const bool has_te = gfl & 1u, accum = gfl & 2u;
        float gir[8], giz[8], gin[8], tev[8], old[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          gir[i] = giz[i] = gin[i] = tev[i] = old[i] = 0.f;
          const int tr_i = __shfl_sync(kFull, trow, i);          // (lane i holds tile row 8 w + i)
          if (i < n_rows) {
            gir[i] = ld_dep_f32(gp);
            giz[i] = ld_dep_f32(gp + D);
            gin[i] = ld_dep_f32(gp + 2 * D);
            if (has_te) tev[i] = __ldg(p.time_embed + static_cast<size_t>(tr_i) * D + j);
            if (accum) old[i] = ld_cg_f32(reinterpret_cast<const char*>(op) + i * out_pitch);
          }
          gp = reinterpret_cast<const float*>(reinterpret_cast<const char*>(gp) + gi_pitch);
        }
        TLW(3);
        // ---- 3. accumulator quadrants r | z | n -> exchange rows [gate][tile row][32] (the idle weight ring) ----
        if (rec) {
          mbar_wait(&S.d_full, mm & 1);
          tc_fence_after();
          TLW(4);
          if (q < 3) {
            float v[32];
            tmem_ld32(tbase + (static_cast<uint32_t>(32 * q) << 16) + col_d + 32 * hf, v);
            const uint32_t exw = ex + static_cast<uint32_t>(q * kWRows + 32 * hf) * 128u + lane * 4u;
#pragma unroll
            for (int i = 0; i < 32; ++i) sts_f32(exw + i * 128u, v[i]);
          }
          tc_fence_before();
          bar_named(1, kWWorkers);
        }
        TLW(5);
        // ---- 4. gates, state store ----
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = 8 * warp + i;
          const int pri = __shfl_sync(kFull, pr, i);
          const float dti = __shfl_sync(kFull, dt, i);
          if (i >= n_rows) continue;
          float hr = br, hz = bz, hn = bn, hp = 0.f;
          if (pri >= 0) {
            const float dec = dti;
            const uint32_t ea = ex + static_cast<uint32_t>(m) * 128u + lane * 4u;
            hr = fmaf(dec, lds_f32(ea), hr);
            hz = fmaf(dec, lds_f32(ea + kWRows * 128u), hz);
            hn = fmaf(dec, lds_f32(ea + 2u * kWRows * 128u), hn);
            const uint32_t off = static_cast<uint32_t>(cb) * kWAtomBytes + sw128_off(static_cast<uint32_t>(m), static_cast<uint32_t>(lane));
            hp = dec * (lds_f32(s_hi + off) + lds_f32(s_lo + off));
          }
          const float rg = sigmoid_w(gir[i] + hr), zg = sigmoid_w(giz[i] + hz);
          const float ng = tanh_w(gin[i] + rg * hn);
          float hy = (1.f - zg) * ng + zg * hp;
          hy += tev[i];
          *reinterpret_cast<float*>(reinterpret_cast<char*>(op) + i * out_pitch) = accum ? old[i] + hy : hy;
        }
        TLW(6);
        if (rec) mm += 1;
        cta_sync_w();   // the operand tile and the exchange rows are rewritten by the next tile step
        TLW(7);
      }
    }
    // self-cleaning barrier words and counters: the last CTA to leave resets them for the next launch
    if (tid == 0) {
      const unsigned done = atomicAdd(P.barrier + 1, 1u);
      if (done == gridDim.x - 1) {
        __threadfence();
        for (int i = 0; i < P.n_steps * tile_stride; ++i) P.barrier[2 + i] = 0u;
        P.barrier[0] = 0u;
        P.barrier[1] = 0u;
        __threadfence();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tbase, 512);
}

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

bool wide_enabled() {
  static const char* mode = getenv("TEMP_WIDE_TC");   // TEMP_WIDE_TC=0: these shapes stay on the fp32 SIMT kernels
  return mode == nullptr || strcmp(mode, "0") != 0;
}

}  // namespace

#ifdef TEMP_TIMELINE
extern "C" int temp_debug_timeline_wide(void* device_buffer) {  // [ctas][8 warps][16 steps][10 marks] u64, or null to disable
  unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
  return cudaMemcpyToSymbol(g_timeline_w, &p, sizeof(p)) == cudaSuccess ? 0 : -2;
}
#endif

namespace temp_internal {

int64_t tcw_packed_bytes(int k, int n) {
  if (!wide_enabled() || k <= 0 || (k & 3) != 0 || k > 32 * kWMaxAtoms || n <= 0) return -1;
  const int KA = (k + 31) / 32, n_mb = (n + 127) / 128;
  return static_cast<int64_t>(n_mb) * KA * kWChunkBytes;
}

int tcw_pack_weights(const float* w_kn, int k, int n, void* packed, cudaStream_t st) {
  if (w_kn == nullptr || packed == nullptr || tcw_packed_bytes(k, n) <= 0)
    return fail(TEMP_EINVAL, "temp_pack_weights: k must be a multiple of 4 in [4, 256] and n positive%s", "");
  const int KA = (k + 31) / 32, n_mb = (n + 127) / 128;
  const int total = n_mb * KA * 4096;
  pack_weights_wide_kernel<<<(total + 255) / 256, 256, 0, st>>>(w_kn, k, n, KA, n_mb, static_cast<uint8_t*>(packed));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack_weights_wide_kernel launch");
  return TEMP_OK;
}

// d % 4 == 0, d <= 256, one un-decayed dense term with its packed image, S x S relation blocks with S in {1, 2, 4}
bool tcw_layer_supported(const TempRgcnLayerArgs* a) {
  if (!wide_enabled()) return false;
  if (a->d <= 0 || (a->d & 3) != 0 || a->d > 32 * kWMaxAtoms || a->n_terms != 1) return false;
  const TempDenseTerm& t = a->terms[0];
  if (t.w_packed == nullptr || t.a_dt != nullptr) return false;
  if (a->row_ptr != nullptr) {
    if (a->si != a->so || (a->si != 1 && a->si != 2 && a->si != 4) || a->agg_scratch == nullptr) return false;
    if (a->agg_lists != 0 && (a->n_agg_rows < 0 || a->n_agg_heavy < 0 || (a->n_agg_rows > 0 && a->agg_rows == nullptr) ||
                              (a->n_agg_heavy > 0 && a->agg_heavy == nullptr)))
      return false;
  }
  if (a->chain_w != nullptr && a->chain_w_packed == nullptr) return false;
  if (a->chain_peers != nullptr) return false;
  return true;
}

int tcw_launch_gather(const TempRgcnLayerArgs* a, cudaStream_t st) {
  const int grid = tc_gather_grid(a);
  if (grid <= 0) return TEMP_OK;
  cudaError_t e;
  if (a->si == 1)
    e = launch_pdl(rgcn_gather_wide_kernel<1>, grid, kGWWarps * 32, 0, st, *a);
  else if (a->si == 2)
    e = launch_pdl(rgcn_gather_wide_kernel<2>, grid, kGWWarps * 32, 0, st, *a);
  else if (a->si == 4)
    e = launch_pdl(rgcn_gather_wide_kernel<4>, grid, kGWWarps * 32, 0, st, *a);
  else
    return fail(TEMP_EUNSUPPORTED, "rgcn_gather_wide_kernel: relation blocks of 1, 2 or 4 channels%s", "");
  if (e != cudaSuccess) return cuda_fail(e, "rgcn_gather_wide_kernel launch");
  return TEMP_OK;
}

int tcw_launch_layer(const TempRgcnLayerArgs* a, cudaStream_t st) {
  const int KA = (a->d + 31) / 32;
  static int configured = 0;      // largest dynamic shared memory size the kernel has been configured for
  const int smem = wide_smem_bytes(KA);
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(rgcn_layer_tcw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wide_smem_bytes(kWMaxAtoms));
    if (e != cudaSuccess) return cuda_fail(e, "rgcn_layer_tcw_kernel");
    configured = wide_smem_bytes(kWMaxAtoms);
  }
  const int rows = a->row1 - a->row0;
  if (tc_gather_grid(a) > 0) {
    if (int rc = tcw_launch_gather(a, st)) return rc;
  }
  const int grid = (rows + kWRows - 1) / kWRows;
  int gy = 1;
  if (a->chain_w_packed != nullptr) {   // few row tiles (the Bi centre step, the last steps of a window): fill the SMs with chain blocks
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int n_mb = (a->chain_n + 127) / 128;
    gy = sms / grid;
    if (gy > n_mb) gy = n_mb;
    if (gy < 1) gy = 1;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, gy);
  cfg.blockDim = dim3(kWThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, rgcn_layer_tcw_kernel, *a, KA);
  if (e != cudaSuccess) return cuda_fail(e, "rgcn_layer_tcw_kernel launch");
  return TEMP_OK;
}

int64_t tcw_packed_gru_bytes(int d) {
  if (!wide_enabled() || d <= 0 || (d & 3) != 0 || d > 32 * kWMaxAtoms) return -1;
  const int KA = (d + 31) / 32;
  return static_cast<int64_t>(KA) * KA * kWChunkBytes;     // ceil(d / 32) column blocks x KA k-atoms
}

int tcw_pack_gru_weights(const float* whh_t, int d, void* packed, cudaStream_t st) {
  if (whh_t == nullptr || packed == nullptr || tcw_packed_gru_bytes(d) <= 0)
    return fail(TEMP_EINVAL, "temp_pack_gru_weights: d must be a multiple of 4 in [4, 256]%s", "");
  const int KA = (d + 31) / 32;
  const int total = KA * KA * 4096;
  pack_gru_wide_kernel<<<(total + 255) / 256, 256, 0, st>>>(whh_t, d, KA, KA, static_cast<uint8_t*>(packed));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack_gru_wide_kernel launch");
  return TEMP_OK;
}

// torch GRU cell, d % 4 == 0, d <= 256, the packed W_hh image when the step reads a previous state
bool tcw_gru_supported(const TempGruArgs* g) {
  if (!wide_enabled()) return false;
  if (g->d <= 0 || (g->d & 3) != 0 || g->d > 32 * kWMaxAtoms || g->cell_type != TEMP_CELL_TORCH_GRU) return false;
  if (g->prev_row != nullptr && (g->whh_packed == nullptr || g->state == nullptr)) return false;
  if (g->push != 0) return false;
  return g->gi != nullptr && g->b_hh != nullptr && g->out != nullptr;
}

int tcw_launch_gru(const TempGruArgs* g, cudaStream_t st) {
  const int rows = g->row1 - g->row0;
  if (rows <= 0) return TEMP_OK;
  const int KA = (g->d + 31) / 32;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gru_step_tcw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wide_smem_bytes(kWMaxAtoms));
    if (e != cudaSuccess) return cuda_fail(e, "gru_step_tcw_kernel");
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((rows + kWRows - 1) / kWRows, KA);
  cfg.blockDim = dim3(kWThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(wide_smem_bytes(KA));
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gru_step_tcw_kernel, *g, KA);
  if (e != cudaSuccess) return cuda_fail(e, "gru_step_tcw_kernel launch");
  return TEMP_OK;
}

// every step of the scan as its own gru_step_tcw_kernel launch (chained by programmatic dependent launch)
bool tcw_scan_supported(const TempGruScanArgs* a) {
  if (a->n_steps <= 0 || a->push_bufs != nullptr || a->push_multicast != nullptr) return false;
  for (int s = 0; s < a->n_steps; ++s)
    if (!tcw_gru_supported(&a->steps[s])) return false;
  return true;
}

// one cooperative gru_scan_tcw_kernel launch when the slice fits tensor memory (d <= 224), else a launch per step
static bool scan_persistent(const TempGruScanArgs* a) {
  static const char* mode = getenv("TEMP_WIDE_SCAN");      // TEMP_WIDE_SCAN=steps: always one launch per step
  if (mode != nullptr && strcmp(mode, "steps") == 0) return false;
  if (a->n_steps < 2 || a->barrier == nullptr || a->steps[0].d > 224) return false;
  for (int s = 1; s < a->n_steps; ++s)
    if (a->steps[s].d != a->steps[0].d) return false;
  return true;
}

int tcw_scan_launches(const TempGruScanArgs* a) {
  if (scan_persistent(a)) return 1;
  int k = 0;
  for (int s = 0; s < a->n_steps; ++s) k += a->steps[s].row1 > a->steps[s].row0 ? 1 : 0;
  return k;
}

int tcw_launch_scan(const TempGruScanArgs* a, cudaStream_t st) {
  if (!scan_persistent(a)) {
    for (int s = 0; s < a->n_steps; ++s)
      if (int rc = tcw_launch_gru(&a->steps[s], st)) return rc;
    return TEMP_OK;
  }
  const int KA = (a->steps[0].d + 31) / 32;
  static bool configured = false;
  static int sms = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gru_scan_tcw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wide_smem_bytes(kWMaxAtoms));
    if (e != cudaSuccess) return cuda_fail(e, "gru_scan_tcw_kernel");
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    configured = true;
  }
  int max_rows = 0;
  for (int s = 0; s < a->n_steps; ++s) {
    const int rows = a->steps[s].row1 - a->steps[s].row0;
    if (rows > max_rows) max_rows = rows;
  }
  if (max_rows <= 0) return TEMP_OK;
  int T = (max_rows + kWRows - 1) / kWRows;
  if (T > sms / KA) T = sms / KA;        // one CTA per SM (shared memory): the whole grid must be resident
  if (T < 1) T = 1;
  // per-tile completion counters when the caller's zeroed words hold them, else grid-wide barriers between the steps
  const int max_tiles = (max_rows + kWRows - 1) / kWRows;
  static const char* sync_mode = getenv("TEMP_WIDE_SCAN");      // TEMP_WIDE_SCAN=grid: always the grid barrier
  int tile_stride = 0;
  if (a->barrier_words >= 2 + a->n_steps * max_tiles && !(sync_mode != nullptr && strcmp(sync_mode, "grid") == 0)) tile_stride = max_tiles;
  int ka = KA;
  void* params[] = {const_cast<TempGruScanArgs*>(a), &ka, &T, &tile_stride};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gru_scan_tcw_kernel), dim3(T * KA), dim3(kWThreads), params,
                                              static_cast<size_t>(wide_smem_bytes(KA)), st);
  if (e != cudaSuccess) return cuda_fail(e, "gru_scan_tcw_kernel launch");
  return TEMP_OK;
}

}  // namespace temp_internal
