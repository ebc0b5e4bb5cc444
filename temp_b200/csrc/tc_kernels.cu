// temp_b200 -- tcgen05 (5th-gen tensor core) kernels for embed_size == hidden_size == 128, the width of every
// shipped TeMP configuration (grid/*/config_*.json).  fp32 parity is kept with a 3xTF32 operand split
// (tc_common.cuh); everything that is not a dense contraction (the CSR gather with the in-register 1x1
// relation projection, norms, bias / activation / time embedding, GRU gates) stays fp32 SIMT.
//
// "Features on TMEM lanes": every GEMM computes D^T[feature, row] so that a warp's 32 lanes hold 32
// consecutive features of one packed row -- global activations are read and written as coalesced 128-byte
// lines and the thread that gathered agg[row][feature] is the thread that later reads D^T[feature][row].
//
//   rgcn_layer_tc_kernel : one CTA per 128 packed rows.
//        workers (8 warps) : self-loop operand rows -> smem (hi/lo, SWIZZLE_128B) ; CSR gather into registers ;
//                            epilogue 1 (agg + D1 + bias, act, time embedding, h_out) ; chain operand X written
//                            in place over the self-loop operand ; chain epilogues (gi / q|k|v + bias)
//        TMA warp          : packed weight chunks (32 KB = 128 features x 32 k, hi + lo) through a 3-stage ring
//        MMA warp          : D1 = W_loop^T . x^T, then per 128 chain features D2 = W_chain . X^T (TMEM ring of 3)
//   gru_scan_tc_kernel   : persistent cooperative scan over the GRU steps of a window; CTA = (block of 32 hidden
//        columns, row tiles of 64); its W_hh slice (r|z|n rows, hi/lo) stays in shared memory for the whole scan.
#include <stdio.h>

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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float decay_factor(float dt, const float* wb, float inv_temperature) {
  if (wb != nullptr) return expf(-fmaxf(fmaf(__ldg(wb), dt, __ldg(wb + 1)), 0.f));
  return expf(-dt * inv_temperature);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

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
// fused RGCN layer, tcgen05
// ------------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;
constexpr int kWorkerWarps = 8;
constexpr int kWorkers = kWorkerWarps * 32;
// Registers are allocated per 4 warps: 3 warpgroups of 168.  The control warpgroup (TMA warp 8, MMA warp 9, two idle
// warps) hands its registers to the two worker warpgroups with setmaxnreg (56 / 224).
constexpr int kLayerThreads = kWorkers + 128;
constexpr int kStages = 3;
constexpr int kBImage = kTileRows * kD * 4;             // 64 KB: one hi (or lo) image of the activation tile
constexpr int kLayerSmem = 2 * kBImage + kStages * kWChunkBytes + 1024;

struct LayerBars {
  uint64_t w_full[kStages], w_empty[kStages];
  uint64_t b_ready, d1_full, x_ready;
  uint64_t d2_full[3], d2_empty[3];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kLayerThreads, 1) rgcn_layer_tc_kernel(const TempRgcnLayerArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + kBImage;
  uint8_t* ring = smem + 2 * kBImage;
  __shared__ LayerBars S;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rbase = p.row0 + blockIdx.x * kTileRows;
  const int n_mb = p.chain_w_packed != nullptr ? p.chain_n >> 7 : 0;

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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;

  if (warp >= kWorkerWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
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
    } else if (warp == 9 && lane == 0) {
      // ===== MMA issuer =====
      {
      const uint32_t idesc = umma_idesc_tf32(128, kTileRows);
      const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo), rg = smem_u32(ring);
      int i = 0;
      mbar_wait(&S.b_ready, 0);
      tc_fence_after();
      for (int ka = 0; ka < kKAtoms; ++ka, ++i) {
        const int st = i % kStages;
        mbar_wait(&S.w_full[st], (i / kStages) & 1);
        tc_fence_after();
        umma_katom_3x(tbase, rg + st * kWChunkBytes, bh + ka * (kTileRows * 128), bl + ka * (kTileRows * 128), idesc, ka == 0);
        umma_commit(&S.w_empty[st]);
      }
      umma_commit(&S.d1_full);
      if (n_mb > 0) {
        mbar_wait(&S.x_ready, 0);
        tc_fence_after();
        for (int mb = 0; mb < n_mb; ++mb) {
          const int slot = mb % 3;
          if (mb >= 3) {
            mbar_wait(&S.d2_empty[slot], ((mb / 3) - 1) & 1);
            tc_fence_after();
          }
          for (int ka = 0; ka < kKAtoms; ++ka, ++i) {
            const int st = i % kStages;
            mbar_wait(&S.w_full[st], (i / kStages) & 1);
            tc_fence_after();
            umma_katom_3x(tbase + 128 + 128 * slot, rg + st * kWChunkBytes, bh + ka * (kTileRows * 128),
                          bl + ka * (kTileRows * 128), idesc, ka == 0);
            umma_commit(&S.w_empty[st]);
          }
          umma_commit(&S.d2_full[slot]);
        }
      }
    }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===== workers: warp (q, hf) owns feature block q (TMEM lane quadrant) of rows [64 hf, 64 hf + 64) =====
    const int q = warp & 3, hf = warp >> 2;
    const int f = 32 * q + lane;
    const int R0 = rbase + 64 * hf;
    const uint32_t my_off = static_cast<uint32_t>(q) * (kTileRows * 128);  // k-atom q of the activation tile

    // ---- 1. self-loop operand rows -> shared memory (hi / lo) ---------------------------------------
    {
      const TempDenseTerm& tm = p.terms[0];
      const int ra = R0 + lane, rb = R0 + 32 + lane;
      const int idxA = ra < p.row1 ? (tm.a_index != nullptr ? __ldg(tm.a_index + ra) : ra) : -1;
      const int idxB = rb < p.row1 ? (tm.a_index != nullptr ? __ldg(tm.a_index + rb) : rb) : -1;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = 16 * b + i;
          const int s = __shfl_sync(kFull, j < 32 ? idxA : idxB, j & 31);
          v[i] = s >= 0 ? __ldg(tm.a + static_cast<size_t>(s) * kD + f) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = 16 * b + i;
          float hi, lo;
          split_tf32(v[i], hi, lo);
          const uint32_t off = my_off + sw128_off(64 * hf + j, lane);
          *reinterpret_cast<float*>(b_hi + off) = hi;
          *reinterpret_cast<float*>(b_lo + off) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(&S.b_ready);
    }

    // ---- 2. aggregation: agg[j] = norm_j * sum_e (x[src_e] * W[rel_e]) * norm_j  (RGCN.py:91-104) -------
    // Edges of the warp's 64 rows are one contiguous CSR range, walked in order (= the reference's
    // summation order); indices are fetched lane-parallel 32 at a time, feature rows 8 edges at a time.
    float agg[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) agg[j] = 0.f;
    if (p.row_ptr != nullptr) {
      const int ra = min(R0 + lane, p.row1), rb = min(R0 + 32 + lane, p.row1);
      const float nA = ra < p.row1 ? __ldg(p.norm + ra) : 0.f;
      const float nB = rb < p.row1 ? __ldg(p.norm + rb) : 0.f;
      const int e_begin = __ldg(p.row_ptr + min(R0, p.row1));
      const int e_end = __ldg(p.row_ptr + min(R0 + 64, p.row1));
      float a8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a8[i] = 0.f;
      int cur_grp = 0;

      auto flush = [&](int g) {  // g is warp-uniform
        float nr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * g + i;
          const float x0 = __shfl_sync(kFull, nA, j & 31), x1 = __shfl_sync(kFull, nB, j & 31);
          nr[i] = j < 32 ? x0 : x1;
        }
        switch (g) {
#define TEMP_FLUSH_CASE(G)                                       \
  case G:                                                        \
    _Pragma("unroll") for (int i = 0; i < 8; ++i) agg[8 * G + i] = a8[i] * nr[i]; \
    break;
          TEMP_FLUSH_CASE(0)
          TEMP_FLUSH_CASE(1)
          TEMP_FLUSH_CASE(2)
          TEMP_FLUSH_CASE(3)
          TEMP_FLUSH_CASE(4)
          TEMP_FLUSH_CASE(5)
          TEMP_FLUSH_CASE(6)
          TEMP_FLUSH_CASE(7)
#undef TEMP_FLUSH_CASE
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) a8[i] = 0.f;
      };

      for (int base = e_begin; base < e_end; base += 32) {
        const int me = base + lane;
        const bool ok = me < e_end;
        const int s = ok ? __ldg(p.e_src + me) : 0;
        const int rl = ok ? __ldg(p.e_rel + me) : 0;
        const int dj = ok ? __ldg(p.e_dst + me) - R0 : 0;  // local destination row, 0..63, non-decreasing
        const float t0 = __shfl_sync(kFull, nA, dj & 31), t1 = __shfl_sync(kFull, nB, dj & 31);
        const float en = dj < 32 ? t0 : t1;                // edge norm = norm of the destination (utils.py:23-28)
        const int cnt = min(32, e_end - base);
        for (int u0 = 0; u0 < cnt; u0 += 8) {
          float xv[8], wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int uu = u0 + u;
            const int su = __shfl_sync(kFull, s, uu & 31), ru = __shfl_sync(kFull, rl, uu & 31);
            const bool v = uu < cnt;
            xv[u] = v ? __ldg(p.x + static_cast<size_t>(su) * kD + f) : 0.f;
            wv[u] = v ? __ldg(p.weight + static_cast<size_t>(ru) * kD + f) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int uu = u0 + u;
            const int dju = __shfl_sync(kFull, dj, uu & 31);
            const float nu = __shfl_sync(kFull, en, uu & 31);
            if (uu < cnt) {
              const float m = (xv[u] * wv[u]) * nu;
              const int grp = dju >> 3, i8 = dju & 7;
              while (cur_grp != grp) {
                flush(cur_grp);
                ++cur_grp;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) a8[i] += (i == i8) ? m : 0.f;
            }
          }
        }
      }
      while (cur_grp < 8) {
        flush(cur_grp);
        ++cur_grp;
      }
    }

    // ---- 3. epilogue 1: out = act(agg (+x) + x . W_loop + bias) ; h_out ; chain operand X ---------------
    const float bias = p.h_bias != nullptr ? __ldg(p.h_bias + f) : 0.f;
    const bool need_te = (p.te_out | p.te_chain) != 0;
    int rtA = p.row_time_scalar, rtB = p.row_time_scalar;
    if (need_te && p.row_time != nullptr) {
      rtA = __ldg(p.row_time + min(R0 + lane, p.row1 - 1));
      rtB = __ldg(p.row_time + min(R0 + 32 + lane, p.row1 - 1));
    }
    mbar_wait(&S.d1_full, 0);
    tc_fence_after();
    int cur_trow = -1;       // rows of a tile mostly share a snapshot, i.e. a time-embedding row: reload on change only
    float cur_te = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(tbase + (static_cast<uint32_t>(32 * q) << 16) + 64 * hf + 32 * c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = 32 * c + i;
        const int r = R0 + j;
        const uint32_t off = my_off + sw128_off(64 * hf + j, lane);
        if (need_te) {
          const int trow = __shfl_sync(kFull, c == 0 ? rtA : rtB, i);  // warp-uniform
          if (trow != cur_trow) {
            cur_trow = trow;
            cur_te = __ldg(p.time_embed + static_cast<size_t>(trow) * kD + f);
          }
        }
        float val = agg[j];
        if (p.residual) val += *reinterpret_cast<const float*>(b_hi + off) + *reinterpret_cast<const float*>(b_lo + off);
        val += v[i];
        val += bias;
        if (p.activation == TEMP_ACT_RELU) val = fmaxf(val, 0.f);
        if (r < p.row1 && p.h_out != nullptr) p.h_out[static_cast<size_t>(r) * kD + f] = p.te_out ? val + cur_te : val;
        if (n_mb > 0) {
          const float xx = r < p.row1 ? (p.te_chain ? val + cur_te : val) : 0.f;
          float hi, lo;
          split_tf32(xx, hi, lo);
          *reinterpret_cast<float*>(b_hi + off) = hi;
          *reinterpret_cast<float*>(b_lo + off) = lo;
        }
      }
    }

    // ---- 4. chain epilogues: chain_out[r, 128 mb + f] = D2 + chain_b -----------------------------------
    if (n_mb > 0) {
      fence_proxy_async();
      mbar_arrive(&S.x_ready);
      for (int mb = 0; mb < n_mb; ++mb) {
        const int slot = mb % 3;
        const float cbias = p.chain_b != nullptr ? __ldg(p.chain_b + 128 * mb + f) : 0.f;
        mbar_wait(&S.d2_full[slot], (mb / 3) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32(tbase + (static_cast<uint32_t>(32 * q) << 16) + 128 + 128 * slot + 64 * hf + 32 * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int r = R0 + 32 * c + i;
            if (r < p.row1) p.chain_out[static_cast<size_t>(r) * p.chain_ld + 128 * mb + f] = v[i] + cbias;
          }
        }
        tc_fence_before();
        mbar_arrive(&S.d2_empty[slot]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------------------------------------
// persistent GRU scan, tcgen05
// ------------------------------------------------------------------------------------------------
constexpr int kScanN = 64;                     // packed rows per tile (UMMA N)
constexpr int kScanThreads = kWorkers + 128;   // 8 worker warps + control warp (TMA + MMA issue) + 3 idle (register
                                               // allocation is per 4 warps anyway)
constexpr int kScanAImage = kKAtoms * kWChunkBytes;   // 128 KB: this CTA's W_hh slice (r|z|n|pad rows, hi + lo)
constexpr int kScanBImage = kScanN * kD * 4;   // 32 KB per hi / lo
constexpr int kScanExBytes = 3 * kScanN * 32 * 4;
constexpr int kScanSmem = kScanAImage + 2 * kScanBImage + kScanExBytes + 1024;

struct ScanBars {
  uint64_t w_full, mma_done;
  uint32_t tmem_base;
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(bar) : "memory");
    } while (static_cast<int>(v - target) < 0);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kScanThreads, 1) gru_scan_tc_kernel(const TempGruScanArgs P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* a_img = smem;
  uint8_t* b_hi = a_img + kScanAImage;
  uint8_t* b_lo = b_hi + kScanBImage;
  float* ex = reinterpret_cast<float*>(b_lo + kScanBImage);  // [3 gates][64 rows][32 hidden]
  __shared__ ScanBars S;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < kWorkerWarps;
  const int cb = blockIdx.x & 3, jb = 32 * cb;
  const int tq = blockIdx.x >> 2, Q = gridDim.x >> 2;

  if (tid == 0) {
    mbar_init(&S.w_full, 1);
    mbar_init(&S.mma_done, 1);
    fence_mbar_init();
  }
  if (warp == kWorkerWarps) tmem_alloc(&S.tmem_base, kScanN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;
  const uint32_t idesc = umma_idesc_tf32(128, kScanN);

  const void* cur_w = nullptr;
  uint32_t w_phase = 0, mma_phase = 0;
  bool w_pending = false;
  unsigned n_bar = 0;

  for (int s = 0; s < P.n_steps; ++s) {
    const TempGruArgs p = P.steps[s];
    if (p.prev_row != nullptr && p.whh_packed != cur_w) {
      // every MMA that read the old image has completed (mma_done is waited on inside each tile)
      if (tid == kWorkers) {
        const uint8_t* src = static_cast<const uint8_t*>(p.whh_packed) + static_cast<size_t>(cb) * kScanAImage;
        mbar_expect_tx(&S.w_full, kScanAImage);
        for (int c = 0; c < kKAtoms; ++c) bulk_g2s(a_img + c * kWChunkBytes, src + c * kWChunkBytes, kWChunkBytes, &S.w_full);
      }
      cur_w = p.whh_packed;
      w_pending = true;
    }
    const bool type1 = p.cell_type == TEMP_CELL_TYPE1;
    bool need_bar = s > 0;
    const int ntiles = (p.row1 - p.row0 + kScanN - 1) / kScanN;
    for (int tile = tq; tile < ntiles || need_bar; tile += Q) {
      const bool has = tile < ntiles;
      const int rb = p.row0 + tile * kScanN;

      // ---- phase A: everything that does not depend on the previous step ---------------------------
      // worker (w, lane): gate math for hidden column jb + lane of rows rb + w + 8u (u = 0..7); the same rows
      // are the ones whose previous state this warp gathers (lane = 4 feature columns).
      float gi_r[8], gi_z[8], gi_n[8], tev[8];
      float br = 0.f, bz = 0.f, bn = 0.f;
      int prv = -1;
      float dec = 1.f;
      if (has && worker) {
        const int j = jb + lane;
        br = __ldg(p.b_hh + j);
        bz = __ldg(p.b_hh + kD + j);
        bn = __ldg(p.b_hh + 2 * kD + j);
        if (lane < 8) {
          const int r = rb + warp + 8 * lane;
          if (r < p.row1 && p.prev_row != nullptr) {
            prv = __ldg(p.prev_row + r);
            if (p.dt != nullptr) dec = decay_factor(__ldg(p.dt + r), p.decay_wb, p.inv_temperature);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = rb + warp + 8 * u;
          gi_r[u] = gi_z[u] = gi_n[u] = tev[u] = 0.f;
          if (r < p.row1) {
            const float* gi = p.gi + static_cast<size_t>(r) * p.gi_ld + p.gi_off + j;
            if (type1) {
              gi_n[u] = __ldg(gi);
            } else {
              gi_r[u] = __ldg(gi);
              gi_z[u] = __ldg(gi + kD);
              gi_n[u] = __ldg(gi + 2 * kD);
            }
            if (p.time_embed != nullptr) {
              const int trow = p.row_time != nullptr ? __ldg(p.row_time + r) : p.row_time_scalar;
              tev[u] = __ldg(p.time_embed + static_cast<size_t>(trow) * kD + j);
            }
          }
        }
      }
      if (need_bar) {
        grid_barrier(P.barrier, (++n_bar) * gridDim.x);
        need_bar = false;
      }
      if (!has) break;

      // ---- phase B: previous-state rows -> smem operand ; gh^T = W_hh . h0^T ; gates ; state write -------
      int any_prev = 0;
      if (worker) {
        float4 v[8];
        float dv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int pr = __shfl_sync(kFull, prv, u);
          dv[u] = __shfl_sync(kFull, dec, u);
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pr >= 0) {
            any_prev = 1;
            v[u] = __ldcg(reinterpret_cast<const float4*>(p.state + static_cast<size_t>(pr) * kD + 4 * lane));
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = warp + 8 * u;
          float4 hi, lo;
          split_tf32(v[u].x * dv[u], hi.x, lo.x);
          split_tf32(v[u].y * dv[u], hi.y, lo.y);
          split_tf32(v[u].z * dv[u], hi.z, lo.z);
          split_tf32(v[u].w * dv[u], hi.w, lo.w);
          const uint32_t off = static_cast<uint32_t>(lane >> 3) * (kScanN * 128) + (i >> 3) * 1024u + (i & 7) * 128u +
                               (((lane & 7) ^ (i & 7)) << 4);
          *reinterpret_cast<float4*>(b_hi + off) = hi;
          *reinterpret_cast<float4*>(b_lo + off) = lo;
        }
        fence_proxy_async();
      }
      any_prev = __syncthreads_or(any_prev);
      if (any_prev) {
        if (tid == kWorkers) {
          if (w_pending) mbar_wait(&S.w_full, w_phase);
          tc_fence_after();
          const uint32_t ai = smem_u32(a_img), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
          for (int ka = 0; ka < kKAtoms; ++ka)
            umma_katom_3x(tbase, ai + ka * kWChunkBytes, bh + ka * (kScanN * 128), bl + ka * (kScanN * 128), idesc, ka == 0);
          umma_commit(&S.mma_done);
        }
        if (w_pending) {
          w_pending = false;
          w_phase ^= 1;
        }
        mbar_wait(&S.mma_done, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if (worker && (warp & 3) < 3) {  // TMEM lane quadrant = gate; columns = rows of the tile
          const int gate = warp & 3, hf = warp >> 2;
          float v[32];
          tmem_ld32(tbase + (static_cast<uint32_t>(32 * gate) << 16) + 32 * hf, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ex[(gate * kScanN + 32 * hf + i) * 32 + lane] = v[i];
        }
        tc_fence_before();
      }
      __syncthreads();

      if (worker) {
        const int j = jb + lane;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = warp + 8 * u;
          const int r = rb + i;
          float hr = br, hz = bz, hn = bn, h0 = 0.f;
          if (any_prev) {
            hr += ex[(0 * kScanN + i) * 32 + lane];
            hz += ex[(1 * kScanN + i) * 32 + lane];
            hn += ex[(2 * kScanN + i) * 32 + lane];
            const uint32_t off = static_cast<uint32_t>(cb) * (kScanN * 128) + sw128_off(i, lane);
            h0 = *reinterpret_cast<const float*>(b_hi + off) + *reinterpret_cast<const float*>(b_lo + off);
          }
          float hy;
          if (type1) {  // GRU_cell.py:22-29
            const float rg = sigmoidf_(hr), zg = sigmoidf_(hz);
            const float ng = tanhf(gi_n[u] + rg * hn);
            hy = ng + zg * (h0 - ng);
          } else {      // torch.nn.GRU, gate order r, z, n
            const float rg = sigmoidf_(gi_r[u] + hr);
            const float zg = sigmoidf_(gi_z[u] + hz);
            const float ng = tanhf(gi_n[u] + rg * hn);
            hy = (1.f - zg) * ng + zg * h0;
          }
          hy += tev[u];
          if (r < p.row1) {
            float* o = p.out + static_cast<size_t>(r) * kD + j;
            *o = p.accumulate ? (__ldcg(o) + hy) : hy;
          }
        }
      }
      __syncthreads();  // the operand tile and the exchange buffer are rewritten by the next tile
    }
    if (w_pending && tid == kWorkers) {
      // the image was requested but no tile of this CTA needed it: drain the copy before a possible reload
      mbar_wait(&S.w_full, w_phase);
    }
    if (w_pending) {
      w_pending = false;
      w_phase ^= 1;
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWorkerWarps) tmem_dealloc(tbase, kScanN);
  // self-cleaning barrier words: the last CTA to leave resets them for the next launch
  if (tid == 0 && P.barrier != nullptr) {
    const unsigned done = atomicAdd(P.barrier + 1, 1u);
    if (done == gridDim.x - 1) {
      P.barrier[0] = 0u;
      P.barrier[1] = 0u;
      __threadfence();
    }
  }
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

namespace temp_internal {

bool tc_layer_supported(const TempRgcnLayerArgs* a) {
  if (a->d != kD || a->n_terms != 1) return false;
  const TempDenseTerm& t = a->terms[0];
  if (t.w_packed == nullptr || t.a_dt != nullptr) return false;
  if (a->row_ptr != nullptr && (a->si != 1 || a->so != 1 || a->e_dst == nullptr)) return false;
  if (a->chain_w != nullptr && (a->chain_w_packed == nullptr || (a->chain_n & 127) != 0)) return false;
  return true;
}

int tc_launch_layer(const TempRgcnLayerArgs* a, cudaStream_t st) {
  static bool configured = false;
  if (int rc = ensure_smem_once(rgcn_layer_tc_kernel, kLayerSmem, "rgcn_layer_tc_kernel", configured)) return rc;
  const int rows = a->row1 - a->row0;
  const int grid = (rows + kTileRows - 1) / kTileRows;
  rgcn_layer_tc_kernel<<<grid, kLayerThreads, kLayerSmem, st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "rgcn_layer_tc_kernel launch");
  return TEMP_OK;
}

bool tc_scan_supported(const TempGruScanArgs* a) {
  for (int s = 0; s < a->n_steps; ++s) {
    const TempGruArgs& g = a->steps[s];
    if (g.d != kD) return false;
    if (g.prev_row != nullptr && g.whh_packed == nullptr) return false;
  }
  return a->n_steps > 0;
}

int tc_launch_scan(const TempGruScanArgs* a, cudaStream_t st) {
  static bool configured = false;
  if (int rc = ensure_smem_once(gru_scan_tc_kernel, kScanSmem, "gru_scan_tc_kernel", configured)) return rc;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  int max_rows = 0;
  for (int s = 0; s < a->n_steps; ++s) max_rows = max(max_rows, a->steps[s].row1 - a->steps[s].row0);
  if (max_rows == 0) return TEMP_OK;
  int want = ((max_rows + kScanN - 1) / kScanN) * 4;
  int cap = sm_count / 4 * 4;  // one CTA per SM (216 KB of shared memory), all co-resident
  int grid = want < cap ? want : cap;
  void* params[] = {const_cast<TempGruScanArgs*>(a)};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gru_scan_tc_kernel), dim3(grid), dim3(kScanThreads), params,
                                              kScanSmem, st);
  if (e != cudaSuccess) return cuda_fail(e, "gru_scan_tc_kernel launch");
  return TEMP_OK;
}

int tc_pack_weights(const float* w_kn, int k, int n, void* packed, cudaStream_t st) {
  if (w_kn == nullptr || packed == nullptr || k != kD || n <= 0 || (n & 127) != 0)
    return fail(TEMP_EINVAL, "temp_pack_weights: k must be 128 and n a positive multiple of 128%s", "");
  pack_weights_kernel<<<(kD * n + 255) / 256, 256, 0, st>>>(w_kn, n, static_cast<uint8_t*>(packed));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack_weights_kernel launch");
  return TEMP_OK;
}

int tc_pack_gru_weights(const float* whh_t, int d, void* packed, cudaStream_t st) {
  if (whh_t == nullptr || packed == nullptr || d != kD) return fail(TEMP_EINVAL, "temp_pack_gru_weights: d must be 128%s", "");
  pack_gru_kernel<<<(4 * kD * 128 + 255) / 256, 256, 0, st>>>(whh_t, static_cast<uint8_t*>(packed));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack_gru_kernel launch");
  return TEMP_OK;
}

}  // namespace temp_internal
