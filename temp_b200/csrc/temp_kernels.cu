// temp_b200 -- sm_100a kernels + C ABI for the TeMP RGCN + GRU/BiGRU/attention forward.
//
// FP32 SIMT path (exact fp32 FFMA arithmetic, deterministic summation order).  See DESIGN.md for
// the data layout and the per-kernel roofline; include/temp_b200.h for the ABI contract and the
// reference call sites each entry point replaces.
//
//   rgcn_layer_kernel : CSR-by-destination gather with the block-diagonal relation projection
//                       applied in registers  ->  self-loop / recurrent dense terms (smem-tiled
//                       GEMM, cp.async double-buffered weights)  ->  bias / activation / time
//                       embedding epilogue  ->  chained GEMM (GRU input gates or attention q|k|v)
//   gru_kernel        : h0 = decay(prev state gather); gh = h0 . W_hh^T; gates; state write
//   attn_kernel       : per-entity multi-head attention over the time slots (online softmax)
#include "temp_b200.h"

#include "internal.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

namespace {
extern thread_local char g_err[512];
}

namespace temp_internal {
int fail(int code, const char* fmt, const char* a, long b) {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return TEMP_ECUDA;
}
}  // namespace temp_internal

namespace {

constexpr int kThreads = 256;
constexpr int kTM = 64;    // packed rows per CTA in the layer kernel
constexpr int kNC = 128;   // output columns per accumulation pass
constexpr int kKC = 32;    // weight rows per cp.async stage
constexpr int kGJ = 32;    // hidden columns per CTA in the GRU kernel

thread_local char g_err[512] = "";

using temp_internal::cuda_fail;
using temp_internal::fail;

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// exp(-dt * inv_temperature)  or  exp(-max(w * dt + b, 0))   (RRGCN.py:83 / RGCN.py:106-107)
__device__ __forceinline__ float decay_factor(float dt, const float* wb, float inv_temperature) {
  if (wb != nullptr) return expf(-fmaxf(fmaf(__ldg(wb), dt, __ldg(wb + 1)), 0.f));
  return expf(-dt * inv_temperature);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Stage kKC x kNC weights  W[k0 + kk][col0 + c]  into ws (row-major [kKC][kNC]); zero outside.
__device__ __forceinline__ void stage_weights(float* ws, const float* __restrict__ w, int ldw, int K, int N,
                                              int k0, int col0) {
#pragma unroll
  for (int q = 0; q < (kKC * kNC / 4) / kThreads; ++q) {
    const int idx = threadIdx.x + q * kThreads;
    const int kk = idx / (kNC / 4);
    const int c4 = idx % (kNC / 4);
    const int k = k0 + kk;
    const int col = col0 + c4 * 4;
    float* dst = ws + kk * kNC + c4 * 4;
    if (k < K && col < N) {
      cp_async16(dst, w + static_cast<size_t>(k) * ldw + col);
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// acc[i][0..3] += As[ty*8 + i][0..K) . W[0..K)[col0 + tx*4 .. +3]      (K padded to Kp in smem)
__device__ __forceinline__ void gemm_accumulate(float4 (&acc)[8], const float* __restrict__ As, int lda, int Kp,
                                                float* Ws, const float* __restrict__ w, int ldw, int K, int N,
                                                int col0) {
  const int tx = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;
  const int nk = Kp / kKC;
  int buf = 0;
  stage_weights(Ws, w, ldw, K, N, 0, col0);
  cp_async_commit();
  for (int kc = 0; kc < nk; ++kc) {
    if (kc + 1 < nk) stage_weights(Ws + (buf ^ 1) * (kKC * kNC), w, ldw, K, N, (kc + 1) * kKC, col0);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* wsb = Ws + buf * (kKC * kNC) + tx * 4;
    const float* asb = As + (ty * 8) * lda + kc * kKC;
#pragma unroll
    for (int kk = 0; kk < kKC; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(asb + i * lda + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 wv = *reinterpret_cast<const float4*>(wsb + (kk + q) * kNC);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = q == 0 ? a[i].x : (q == 1 ? a[i].y : (q == 2 ? a[i].z : a[i].w));
          acc[i].x = fmaf(av, wv.x, acc[i].x);
          acc[i].y = fmaf(av, wv.y, acc[i].y);
          acc[i].z = fmaf(av, wv.z, acc[i].z);
          acc[i].w = fmaf(av, wv.w, acc[i].w);
        }
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// fused RGCN layer
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) rgcn_layer_kernel(const TempRgcnLayerArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.d;
  const int Kp = (D + kKC - 1) / kKC * kKC;
  const int lda = Kp + 4;
  const bool has_chain = p.chain_w != nullptr;
  float* As = smem;
  float* Xs = As + kTM * lda;
  float* Ws = has_chain ? Xs + kTM * lda : Xs;

  const int tx = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;
  const int rbase = p.row0 + blockIdx.x * kTM;
  const int nvec = D >> 2;

  // zero the K padding columns once (loads below only touch columns < D)
  if (Kp > D) {
    const int padw = Kp - D;
    for (int idx = threadIdx.x; idx < kTM * padw; idx += kThreads) {
      const int m = idx / padw, c = D + idx % padw;
      As[m * lda + c] = 0.f;
      if (has_chain) Xs[m * lda + c] = 0.f;
    }
  }

  const int n_chunks = (D + kNC - 1) / kNC;
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int c = chunk * kNC + tx * 4;  // this lane's first output column
    float4 acc[8];

    // ---- 1. aggregation: one warp per destination row, lanes across channels ----------------
    if (p.row_ptr != nullptr) {
      const bool diag = (p.si == 1 && p.so == 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // unrolled: acc[] must stay in registers
        const int r = rbase + ty * 8 + i;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.row1 && c < D) {
          const int p0 = __ldg(p.row_ptr + r), p1 = __ldg(p.row_ptr + r + 1);
          const float nrm = __ldg(p.norm + r);
          if (diag) {
            for (int e = p0; e < p1; e += 4) {
              int s[4], rl[4];
              float4 hv[4], wv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int ee = min(e + u, p1 - 1);
                s[u] = __ldg(p.e_src + ee);
                rl[u] = __ldg(p.e_rel + ee);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                hv[u] = ldg4(p.x + static_cast<size_t>(s[u]) * D + c);
                wv[u] = ldg4(p.weight + static_cast<size_t>(rl[u]) * D + c);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (e + u < p1) {  // msg = (h * w) * norm_e, summed in edge order (RGCN.py:92-97)
                  a.x += (hv[u].x * wv[u].x) * nrm;
                  a.y += (hv[u].y * wv[u].y) * nrm;
                  a.z += (hv[u].z * wv[u].z) * nrm;
                  a.w += (hv[u].w * wv[u].w) * nrm;
                }
              }
            }
          } else {
            const int si = p.si, so = p.so;
            const int wrow = p.n_bases * si * so;
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            for (int e = p0; e < p1; ++e) {
              const float* hs = p.x + static_cast<size_t>(__ldg(p.e_src + e)) * D;
              const float* wr = p.weight + static_cast<size_t>(__ldg(p.e_rel + e)) * wrow;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int co = c + q;
                const int b = co / so, j = co - b * so;
                float m = 0.f;
                for (int ii = 0; ii < si; ++ii)
                  m = fmaf(__ldg(hs + b * si + ii), __ldg(wr + (b * si + ii) * so + j), m);
                av[q] += m * nrm;
              }
            }
            a = make_float4(av[0], av[1], av[2], av[3]);
          }
          a.x *= nrm; a.y *= nrm; a.z *= nrm; a.w *= nrm;  // apply_func (RGCN.py:103-104)
        }
        acc[i] = a;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // ---- 2. dense terms: self loop, recurrent projections -------------------------------------
    for (int t = 0; t < p.n_terms; ++t) {
      const TempDenseTerm& tm = p.terms[t];
      __syncthreads();  // previous readers of As are done
      for (int idx = threadIdx.x; idx < kTM * nvec; idx += kThreads) {
        const int m = idx / nvec, c4 = idx - m * nvec;
        const int r = rbase + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.row1) {
          const int srow = tm.a_index != nullptr ? __ldg(tm.a_index + r) : r;
          if (srow >= 0) {
            v = ldg4(tm.a + static_cast<size_t>(srow) * D + c4 * 4);
            if (tm.a_dt != nullptr) {
              const float f = decay_factor(__ldg(tm.a_dt + r), tm.decay_wb, p.inv_temperature);
              v.x *= f; v.y *= f; v.z *= f; v.w *= f;
            }
          }
        }
        *reinterpret_cast<float4*>(As + m * lda + c4 * 4) = v;
      }
      __syncthreads();
      if (t == 0 && p.residual && c < D) {  // forward_isolated: x + x.W_loop (RGCN.py:83)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(As + (ty * 8 + i) * lda + c);
          acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
        }
      }
      gemm_accumulate(acc, As, lda, Kp, Ws, tm.w, D, D, D, chunk * kNC);
    }

    // ---- 3. epilogue ------------------------------------------------------------------------
    if (c < D) {
      float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.h_bias != nullptr) bias = ldg4(p.h_bias + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = ty * 8 + i;
        const int r = rbase + m;
        if (r >= p.row1) continue;
        float4 v = acc[i];
        v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
        if (p.activation == TEMP_ACT_RELU) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        float4 te = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.te_out || p.te_chain) {
          const int trow = p.row_time != nullptr ? __ldg(p.row_time + r) : p.row_time_scalar;
          te = ldg4(p.time_embed + static_cast<size_t>(trow) * D + c);
        }
        if (p.h_out != nullptr && blockIdx.y == 0) {   // (with grid.y > 1 every column block recomputes the tile; one writes it)
          float4 o = v;
          if (p.te_out) { o.x += te.x; o.y += te.y; o.z += te.z; o.w += te.w; }
          *reinterpret_cast<float4*>(p.h_out + static_cast<size_t>(r) * D + c) = o;
        }
        if (has_chain) {
          float4 o = v;
          if (p.te_chain) { o.x += te.x; o.y += te.y; o.z += te.z; o.w += te.w; }
          *reinterpret_cast<float4*>(Xs + m * lda + c) = o;
        }
      }
    } else if (has_chain) {
      // rows past row1 never get written above; keep chain input finite
    }
  }

  // ---- 4. chained GEMM on the tile that is still in shared memory ------------------------------
  if (has_chain) {
    // rows >= row1 of Xs were never written: zero them so the FMAs stay finite
    for (int idx = threadIdx.x; idx < kTM * nvec; idx += kThreads) {
      const int m = idx / nvec, c4 = idx - m * nvec;
      if (rbase + m >= p.row1) *reinterpret_cast<float4*>(Xs + m * lda + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int NCn = p.chain_n;
    const int chunks2 = (NCn + kNC - 1) / kNC;
    // launches with few row tiles spread the chained GEMM's column chunks over grid.y (each y block rebuilds the small
    // input tile above, then takes every gridDim.y-th chunk)
    for (int ch = blockIdx.y; ch < chunks2; ch += gridDim.y) {
      const int c = ch * kNC + tx * 4;
      float4 acc[8];
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.chain_b != nullptr && c < NCn) b = ldg4(p.chain_b + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = b;
      gemm_accumulate(acc, Xs, lda, Kp, Ws, p.chain_w, NCn, D, NCn, ch * kNC);
      if (c < NCn) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rbase + ty * 8 + i;
          if (r < p.row1) *reinterpret_cast<float4*>(p.chain_out + static_cast<size_t>(r) * p.chain_ld + c) = acc[i];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GRU recurrent half + gates.  Work item = (RPW*8 rows) x (kGJ hidden columns, all three gates)
// ------------------------------------------------------------------------------------------------
// this CTA's [D][3*kGJ] slice of W_hh^T -> shared memory (cp.async; caller commits / waits)
__device__ __forceinline__ void gru_stage_weights(float* Ws, const float* __restrict__ whh_t, int D, int jb) {
  const int per_k = 3 * kGJ / 4;  // float4 per k row
  for (int idx = threadIdx.x; idx < D * per_k; idx += kThreads) {
    const int k = idx / per_k, v = idx - k * per_k;
    const int g = v / (kGJ / 4), c4 = v - g * (kGJ / 4);
    const int col = jb + c4 * 4;
    float* dst = Ws + k * (3 * kGJ) + g * kGJ + c4 * 4;
    if (col < D) cp_async16(dst, whh_t + static_cast<size_t>(k) * (3 * D) + g * D + col);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- one GRU work item, split so that everything that does NOT depend on the recurrence (indices,
// decay factors, the precomputed input gates, time-embedding rows) is in flight BEFORE the step's
// dependency point (kernel start / grid barrier) --------------------------------------------------
constexpr int kGruPre = 4;  // gather elements per thread whose indices are prefetched (TM*D/4 <= 4*256)

template <int RPW>
struct GruPrefetch {
  int pr[kGruPre];        // previous-state row per gather element (-1: zero state)
  float decay[kGruPre];
  float gi[RPW][3];       // input-gate pre-activations of this thread's rows at its hidden column
  float te[RPW];          // time-embedding value to add
  float br, bz, bn;
};

template <int RPW>
__device__ __forceinline__ void gru_phase_a(const TempGruArgs& p, int rbase, int jb, GruPrefetch<RPW>& f) {
  constexpr int TM = RPW * 8;
  const int D = p.d, nvec = D >> 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (TM * nvec <= kGruPre * kThreads) {
#pragma unroll
    for (int u = 0; u < kGruPre; ++u) {
      const int idx = threadIdx.x + u * kThreads;
      const int r = rbase + idx / nvec;
      int pr = -1;
      float dec = 1.f;
      if (idx < TM * nvec && r < p.row1 && p.prev_row != nullptr) {
        pr = __ldg(p.prev_row + r);
        if (p.dt != nullptr) dec = decay_factor(__ldg(p.dt + r), p.decay_wb, p.inv_temperature);
      }
      f.pr[u] = pr;
      f.decay[u] = dec;
    }
  }
  const int j = jb + tx;
  const bool jok = j < D;
  f.br = jok ? __ldg(p.b_hh + j) : 0.f;
  f.bz = jok ? __ldg(p.b_hh + D + j) : 0.f;
  f.bn = jok ? __ldg(p.b_hh + 2 * D + j) : 0.f;
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = rbase + ty * RPW + i;
    f.gi[i][0] = f.gi[i][1] = f.gi[i][2] = 0.f;
    f.te[i] = 0.f;
    if (jok && r < p.row1) {
      const float* gi = p.gi + static_cast<size_t>(r) * p.gi_ld + p.gi_off;
      f.gi[i][0] = __ldg(gi + j);
      if (p.cell_type != TEMP_CELL_TYPE1) {
        f.gi[i][1] = __ldg(gi + D + j);
        f.gi[i][2] = __ldg(gi + 2 * D + j);
      }
      if (p.time_embed != nullptr) {
        const int trow = p.row_time != nullptr ? __ldg(p.row_time + r) : p.row_time_scalar;
        f.te[i] = __ldg(p.time_embed + static_cast<size_t>(trow) * D + j);
      }
    }
  }
}

// previous-state gather (the only loads that depend on the previous step) -> GEMM -> gates -> store
template <int RPW, bool kCoherent>
__device__ __forceinline__ void gru_phase_b(const TempGruArgs& p, float* As, int lda, const float* Ws, int rbase,
                                            int jb, const GruPrefetch<RPW>& f, bool& w_pending) {
  constexpr int TM = RPW * 8;
  const int D = p.d, nvec = D >> 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int any_prev = 0;
  if (TM * nvec <= kGruPre * kThreads) {
    float4 v[kGruPre];
#pragma unroll
    for (int u = 0; u < kGruPre; ++u) {
      const int idx = threadIdx.x + u * kThreads;
      const int c4 = idx % nvec;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f.pr[u] >= 0) {
        any_prev = 1;
        const float4* src = reinterpret_cast<const float4*>(p.state + static_cast<size_t>(f.pr[u]) * D + c4 * 4);
        v[u] = kCoherent ? __ldcg(src) : __ldg(src);
      }
    }
#pragma unroll
    for (int u = 0; u < kGruPre; ++u) {
      const int idx = threadIdx.x + u * kThreads;
      if (idx < TM * nvec) {
        const int m = idx / nvec, c4 = idx - m * nvec;
        const float d = f.decay[u];
        *reinterpret_cast<float4*>(As + m * lda + c4 * 4) = make_float4(v[u].x * d, v[u].y * d, v[u].z * d, v[u].w * d);
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < TM * nvec; idx += kThreads) {
      const int m = idx / nvec, c4 = idx - m * nvec;
      const int r = rbase + m;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < p.row1 && p.prev_row != nullptr) {
        const int pr = __ldg(p.prev_row + r);
        if (pr >= 0) {
          any_prev = 1;
          const float4* src = reinterpret_cast<const float4*>(p.state + static_cast<size_t>(pr) * D + c4 * 4);
          v = kCoherent ? __ldcg(src) : __ldg(src);
          if (p.dt != nullptr) {
            const float d = decay_factor(__ldg(p.dt + r), p.decay_wb, p.inv_temperature);
            v.x *= d; v.y *= d; v.z *= d; v.w *= d;
          }
        }
      }
      *reinterpret_cast<float4*>(As + m * lda + c4 * 4) = v;
    }
  }
  if (w_pending) {
    cp_async_wait<0>();
    w_pending = false;
  }
  any_prev = __syncthreads_or(any_prev);

  float acc[RPW][3];
#pragma unroll
  for (int i = 0; i < RPW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
  if (any_prev) {
    const float* asb = As + (ty * RPW) * lda;
    const float* wsb = Ws + tx;
#pragma unroll 4
    for (int k = 0; k < D; k += 4) {
      float4 a[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) a[i] = *reinterpret_cast<const float4*>(asb + i * lda + k);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float w0 = wsb[(k + q) * (3 * kGJ)];
        const float w1 = wsb[(k + q) * (3 * kGJ) + kGJ];
        const float w2 = wsb[(k + q) * (3 * kGJ) + 2 * kGJ];
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const float av = q == 0 ? a[i].x : (q == 1 ? a[i].y : (q == 2 ? a[i].z : a[i].w));
          acc[i][0] = fmaf(av, w0, acc[i][0]);
          acc[i][1] = fmaf(av, w1, acc[i][1]);
          acc[i][2] = fmaf(av, w2, acc[i][2]);
        }
      }
    }
  }

  const int j = jb + tx;
  if (j < D) {
    float hy[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int m = ty * RPW + i;
      const float h0 = As[m * lda + j];
      const float hr = acc[i][0] + f.br, hz = acc[i][1] + f.bz, hn = acc[i][2] + f.bn;
      if (p.cell_type == TEMP_CELL_TYPE1) {  // GRU_cell.py:22-29
        const float rg = sigmoidf_(hr), zg = sigmoidf_(hz);
        const float ng = tanhf(f.gi[i][0] + rg * hn);
        hy[i] = ng + zg * (h0 - ng);
      } else {  // torch.nn.GRU, gate order r, z, n
        const float rg = sigmoidf_(f.gi[i][0] + hr);
        const float zg = sigmoidf_(f.gi[i][1] + hz);
        const float ng = tanhf(f.gi[i][2] + rg * hn);
        hy[i] = (1.f - zg) * ng + zg * h0;
      }
      hy[i] += f.te[i];
    }
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = rbase + ty * RPW + i;
      if (r < p.row1) {
        float* o = p.out + static_cast<size_t>(r) * D + j;
        *o = p.accumulate ? (__ldcg(o) + hy[i]) : hy[i];
      }
    }
  }
}

template <int RPW>
__global__ void __launch_bounds__(kThreads, 2) gru_kernel(const TempGruArgs p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TM = RPW * 8;
  const int D = p.d;
  const int lda = D + 4;
  float* As = smem;             // [TM][lda]   decayed previous state rows
  float* Ws = As + TM * lda;    // [D][3*kGJ]  this CTA's slice of W_hh^T
  const int rbase = p.row0 + blockIdx.x * TM;
  const int jb = blockIdx.y * kGJ;
  bool w_pending = false;
  if (p.prev_row != nullptr) {
    gru_stage_weights(Ws, p.whh_t, D, jb);
    cp_async_commit();
    w_pending = true;
  }
  GruPrefetch<RPW> f;
  gru_phase_a<RPW>(p, rbase, jb, f);
  gru_phase_b<RPW, false>(p, As, lda, Ws, rbase, jb, f, w_pending);
  if (w_pending) cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Persistent scan: every GRU step of a window in ONE cooperative launch.  CTA c owns hidden-column
// block (c % JB) for the whole scan (its W_hh^T slice stays in shared memory across steps) and walks
// row tiles c / JB, c / JB + Q, ...; steps are separated by a grid-wide barrier.
// ------------------------------------------------------------------------------------------------
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

template <int RPW>
__global__ void __launch_bounds__(kThreads, 2) gru_scan_kernel(const TempGruScanArgs P) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TM = RPW * 8;
  const int D = P.steps[0].d;
  const int lda = D + 4;
  float* As = smem;
  float* Ws = As + TM * lda;
  const int JB = (D + kGJ - 1) / kGJ;
  const int jb = (blockIdx.x % JB) * kGJ;
  const int q = blockIdx.x / JB;
  const int Q = gridDim.x / JB;
  const float* cur_w = nullptr;
  unsigned n_bar = 0;

  for (int s = 0; s < P.n_steps; ++s) {
    const TempGruArgs p = P.steps[s];
    bool w_pending = false;
    if (p.whh_t != cur_w && p.prev_row != nullptr) {  // weights are static: stage before the dependency point
      gru_stage_weights(Ws, p.whh_t, D, jb);
      cp_async_commit();
      cur_w = p.whh_t;
      w_pending = true;
    }
    bool need_bar = s > 0;
    const int ntiles = (p.row1 - p.row0 + TM - 1) / TM;
    for (int tile = q; tile < ntiles || need_bar; tile += Q) {
      const bool has = tile < ntiles;
      const int rbase = p.row0 + tile * TM;
      GruPrefetch<RPW> f;
      if (has) gru_phase_a<RPW>(p, rbase, jb, f);
      if (need_bar) {
        grid_barrier(P.barrier, (++n_bar) * gridDim.x);
        need_bar = false;
      }
      if (!has) break;
      gru_phase_b<RPW, true>(p, As, lda, Ws, rbase, jb, f, w_pending);
      __syncthreads();  // As (and possibly Ws) are rewritten next
    }
    if (w_pending) cp_async_wait<0>();
    __syncthreads();
  }
  // self-cleaning barrier words: the last CTA to leave resets them for the next launch
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(P.barrier + 1, 1u);
    if (done == gridDim.x - 1) {
      P.barrier[0] = 0u;
      P.barrier[1] = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention over time slots: one warp per packed row, 32/heads lanes per head
// ------------------------------------------------------------------------------------------------
constexpr int kAttnMaxT = 16;  // channels per lane: dk / lanes_per_head

__global__ void __launch_bounds__(kThreads) attn_kernel(const TempAttnArgs p) {
  const int lane = threadIdx.x & 31;
  const int r = p.row0 + blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (r >= p.row1) return;
  const int D = p.d, H = p.heads;
  const int dk = D / H;
  const int lph = 32 / H;  // lanes per head
  const int head = lane / lph, sub = lane - head * lph;
  const int nt = (dk - sub + lph - 1) / lph;  // channels j = sub + lph*t < dk
  const float inv_sqrt = 1.f / sqrtf(static_cast<float>(dk));

  const float* qrow = p.qkv + static_cast<size_t>(r) * 3 * D;
  float q[kAttnMaxT], o[kAttnMaxT];
#pragma unroll
  for (int t = 0; t < kAttnMaxT; ++t) {
    q[t] = t < nt ? __ldg(qrow + head * dk + sub + lph * t) : 0.f;
    o[t] = 0.f;
  }
  float mx = -INFINITY, den = 0.f;
  for (int s = 0; s <= p.n_slots; ++s) {
    const float* kp;
    const float* vp;
    if (s < p.n_slots) {
      const int hr = __ldg(p.slot_row + static_cast<size_t>(r) * p.n_slots + s);
      if (hr < 0) continue;  // inactive slot: mask -1e10 -> softmax weight exactly 0 (SARGCN.py:50)
      kp = p.kv_hist + static_cast<size_t>(hr) * 2 * D;
      vp = kp + D;
    } else {
      kp = qrow + D;
      vp = qrow + 2 * D;
    }
    float part = 0.f;
#pragma unroll
    for (int t = 0; t < kAttnMaxT; ++t)
      if (t < nt) part = fmaf(q[t], __ldg(kp + head * dk + sub + lph * t), part);
    for (int off = lph >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    float sc = part * inv_sqrt;
    if (p.decay_wb != nullptr) sc -= fmaxf(fmaf(__ldg(p.decay_wb), __ldg(p.tau + s), __ldg(p.decay_wb + 1)), 0.f);
    const float mnew = fmaxf(mx, sc);
    const float scale = expf(mx - mnew);  // exp(-inf) = 0 on the first slot
    const float w = expf(sc - mnew);
    den = den * scale + w;
#pragma unroll
    for (int t = 0; t < kAttnMaxT; ++t)
      if (t < nt) o[t] = fmaf(w, __ldg(vp + head * dk + sub + lph * t), o[t] * scale);
    mx = mnew;
  }
  const float inv = 1.f / den;
  float* orow = p.out + static_cast<size_t>(r) * D;
#pragma unroll
  for (int t = 0; t < kAttnMaxT; ++t) {
    if (t < nt) {
      const int j = sub + lph * t;
      float v = o[t] * inv;
      float* dst = orow + j * H + head;  // [d_k major, head minor]
      if (p.combine_max) v = fmaxf(v, *dst);
      *dst = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fused training-time scorer: candidate gather + ComplEx / DistMult / TransE + cross-entropy (label 0)
// ------------------------------------------------------------------------------------------------
// One warp per positive triple; 8 lanes per candidate (a lane owns d/8 consecutive channels of the query and of the
// gathered candidate row: 4 candidates per warp iteration, each row read once as 8 x (d/8 x 4) contiguous bytes),
// 3 shuffles reduce a candidate's score, an online logsumexp runs per 8-lane group and the 4 groups are merged at the
// end.  Every score is linear in the candidate (or an L1 distance to a query), so the triple's query vector is formed
// once:  ComplEx tail  q = [re_s re_r - im_s im_r ; re_s im_r + im_s re_r],  head  q = [re_r re_o + im_r im_o ;
// re_r im_o - im_r re_o];  DistMult  q = s r | r o;  TransE  score = -|cand - q|_1 with q = o - r (head) and
// score = -|q - cand|_1 with q = s + r (tail).
constexpr int kScoreMaxPerLane = 32;  // d / 8 <= 32, i.e. d <= 256

// query vector of one triple for this lane's channels [c0, c0 + per)
__device__ __forceinline__ void score_query(float (&q)[kScoreMaxPerLane], const float* fixed, const float* rel, int c0,
                                            int per, int half, int score_fn, int corrupt_tail) {
#pragma unroll
  for (int t = 0; t < kScoreMaxPerLane; ++t) {
    q[t] = 0.f;
    if (t < per) {
      const int c = c0 + t;
      if (score_fn == TEMP_SCORE_COMPLEX) {
        const bool im = c >= half;               // the second half of the channels is the imaginary part
        const int k = im ? c - half : c;
        const float re_e = __ldg(fixed + k), im_e = __ldg(fixed + half + k);
        const float re_r = __ldg(rel + k), im_r = __ldg(rel + half + k);
        if (corrupt_tail) q[t] = im ? re_e * im_r + im_e * re_r : re_e * re_r - im_e * im_r;
        else q[t] = im ? re_r * im_e - im_r * re_e : re_r * re_e + im_r * im_e;
      } else if (score_fn == TEMP_SCORE_DISTMULT) {
        q[t] = __ldg(fixed + c) * __ldg(rel + c);
      } else {
        q[t] = corrupt_tail ? __ldg(fixed + c) + __ldg(rel + c) : __ldg(fixed + c) - __ldg(rel + c);
      }
    }
  }
}

// this lane's share of one candidate's score (row = the candidate's channels [c0, c0 + per)); the 8 lanes of a
// candidate add their shares with score_reduce8
__device__ __forceinline__ float score_partial(const float4* row, const float (&q)[kScoreMaxPerLane], int per, bool l1) {
  float part = 0.f;
#pragma unroll
  for (int t4 = 0; t4 < kScoreMaxPerLane / 4; ++t4) {
    if (4 * t4 < per) {
      const float4 v = __ldg(row + t4);
      if (l1) {  // head: |cand + r - o| = |cand - q| ; tail: |s + r - cand| = |q - cand|
        part += fabsf(v.x - q[4 * t4]) + fabsf(v.y - q[4 * t4 + 1]) + fabsf(v.z - q[4 * t4 + 2]) + fabsf(v.w - q[4 * t4 + 3]);
      } else {
        part = fmaf(v.x, q[4 * t4], part);
        part = fmaf(v.y, q[4 * t4 + 1], part);
        part = fmaf(v.z, q[4 * t4 + 2], part);
        part = fmaf(v.w, q[4 * t4 + 3], part);
      }
    }
  }
  return part;
}

// compile-time width (PER4 float4 loads per lane): branch-free, so that independent candidates' loads can be batched
template <int PER4>
__device__ __forceinline__ void score_load(float4 (&v)[PER4], const float4* row) {
#pragma unroll
  for (int t4 = 0; t4 < PER4; ++t4) v[t4] = __ldg(row + t4);
}

template <int PER4>
__device__ __forceinline__ float score_dot(const float4 (&v)[PER4], const float (&q)[4 * PER4], bool l1) {
  float a = 0.f, b = 0.f, c = 0.f, d = 0.f;   // four short chains instead of one of 4 * PER4 dependent operations
#pragma unroll
  for (int t4 = 0; t4 < PER4; ++t4) {
    if (l1) {
      a += fabsf(v[t4].x - q[4 * t4]);
      b += fabsf(v[t4].y - q[4 * t4 + 1]);
      c += fabsf(v[t4].z - q[4 * t4 + 2]);
      d += fabsf(v[t4].w - q[4 * t4 + 3]);
    } else {
      a = fmaf(v[t4].x, q[4 * t4], a);
      b = fmaf(v[t4].y, q[4 * t4 + 1], b);
      c = fmaf(v[t4].z, q[4 * t4 + 2], c);
      d = fmaf(v[t4].w, q[4 * t4 + 3], d);
    }
  }
  return (a + b) + (c + d);
}

__device__ __forceinline__ float score_reduce8(float part) {
  part += __shfl_xor_sync(0xffffffffu, part, 4);
  part += __shfl_xor_sync(0xffffffffu, part, 2);
  part += __shfl_xor_sync(0xffffffffu, part, 1);
  return part;
}

__global__ void __launch_bounds__(kThreads) score_loss_kernel(const TempScoreLossArgs p) {
  const int lane = threadIdx.x & 31;
  const int pos = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (pos >= p.n_pos) return;
  const int D = p.d, per = D >> 3, half = D >> 1;
  const int grp = lane >> 3, sub = lane & 7;
  const int c0 = sub * per;                      // this lane's first channel
  const long long s_id = p.triples[3 * pos], r_id = p.triples[3 * pos + 1], o_id = p.triples[3 * pos + 2];
  const float* fixed = p.ent_embed + static_cast<size_t>(p.corrupt_tail ? s_id : o_id) * D;
  const float* rel = p.rel_embeds + static_cast<size_t>(r_id) * D;
  float q[kScoreMaxPerLane];
  score_query(q, fixed, rel, c0, per, half, p.score_fn, p.corrupt_tail);
  const bool l1 = p.score_fn == TEMP_SCORE_TRANSE;
  float mx = -INFINITY, den = 0.f, first = 0.f;
  const int64_t* cand = p.cand + static_cast<size_t>(pos) * p.n_cand;
  for (int base = 0; base < p.n_cand; base += 4) {
    const int c = base + grp;
    float part = 0.f;
    if (c < p.n_cand)
      part = score_partial(reinterpret_cast<const float4*>(p.table + static_cast<size_t>(cand[c]) * D + c0), q, per, l1);
    part = score_reduce8(part);
    if (c < p.n_cand) {
      const float sc = l1 ? -part : part;
      if (c == 0) first = sc;
      const float mnew = fmaxf(mx, sc);
      den = den * expf(mx - mnew) + expf(sc - mnew);
      mx = mnew;
    }
  }
  // merge the 4 groups' (max, sum) pairs; candidate 0 lives in group 0
#pragma unroll
  for (int off = 16; off >= 8; off >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, off), oden = __shfl_xor_sync(0xffffffffu, den, off);
    const float mnew = fmaxf(mx, omx);
    den = (mx == -INFINITY ? 0.f : den * expf(mx - mnew)) + (omx == -INFINITY ? 0.f : oden * expf(omx - mnew));
    mx = mnew;
  }
  first = __shfl_sync(0xffffffffu, first, 0);
  if (lane == 0) p.loss[pos] = (mx + logf(den)) - first;
}

// ------------------------------------------------------------------------------------------------
// backward of the fused scorer
// ------------------------------------------------------------------------------------------------
// Same work split as the forward (one warp per positive, 8 lanes per candidate, 4 candidates per warp iteration).  Pass 1
// recomputes the logsumexp of the positive's candidate scores, pass 2 walks the candidates again: the weight
// w = softmax - [c == 0] times the upstream gradient scales (i) the candidate row's gradient -- q for the bilinear
// scores, -sign(row - q) for TransE -- added to grad_table with one 16-byte vector atomic per 4 channels, and (ii) the
// query vector's gradient, kept in registers, reduced over the warp's 4 candidate groups at the end and pushed through
// the query's definition into grad_ent_embed / grad_rel_embeds.
__global__ void __launch_bounds__(kThreads) score_loss_bwd_kernel(const TempScoreLossBwdArgs p) {
  const int lane = threadIdx.x & 31;
  const int pos = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (pos >= p.n_pos) return;
  const int D = p.d, per = D >> 3, half = D >> 1;
  const int grp = lane >> 3, sub = lane & 7;
  const int c0 = sub * per;
  const long long s_id = p.triples[3 * pos], r_id = p.triples[3 * pos + 1], o_id = p.triples[3 * pos + 2];
  const size_t fixed_row = static_cast<size_t>(p.corrupt_tail ? s_id : o_id) * D;
  const float* fixed = p.ent_embed + fixed_row;
  const float* rel = p.rel_embeds + static_cast<size_t>(r_id) * D;
  float q[kScoreMaxPerLane], gq[kScoreMaxPerLane];
  score_query(q, fixed, rel, c0, per, half, p.score_fn, p.corrupt_tail);
#pragma unroll
  for (int t = 0; t < kScoreMaxPerLane; ++t) gq[t] = 0.f;
  const bool l1 = p.score_fn == TEMP_SCORE_TRANSE;
  const int64_t* cand = p.cand + static_cast<size_t>(pos) * p.n_cand;
  // ---- pass 1: logsumexp (the forward's arithmetic) -------------------------------------------------------------------
  float mx = -INFINITY, den = 0.f;
  for (int base = 0; base < p.n_cand; base += 4) {
    const int c = base + grp;
    float part = 0.f;
    if (c < p.n_cand)
      part = score_partial(reinterpret_cast<const float4*>(p.table + static_cast<size_t>(cand[c]) * D + c0), q, per, l1);
    part = score_reduce8(part);
    if (c < p.n_cand) {
      const float sc = l1 ? -part : part;
      const float mnew = fmaxf(mx, sc);
      den = den * expf(mx - mnew) + expf(sc - mnew);
      mx = mnew;
    }
  }
#pragma unroll
  for (int off = 16; off >= 8; off >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, off), oden = __shfl_xor_sync(0xffffffffu, den, off);
    const float mnew = fmaxf(mx, omx);
    den = (mx == -INFINITY ? 0.f : den * expf(mx - mnew)) + (omx == -INFINITY ? 0.f : oden * expf(omx - mnew));
    mx = mnew;
  }
  const float lse = mx + logf(den);
  const float g = __ldg(p.grad_loss + pos);
  // ---- pass 2: gradients -------------------------------------------------------------------------------------------------
  for (int base = 0; base < p.n_cand; base += 4) {
    const int c = base + grp;
    const bool on = c < p.n_cand;
    const size_t row_off = on ? static_cast<size_t>(cand[c]) * D + c0 : 0;
    const float4* row = reinterpret_cast<const float4*>(p.table + row_off);
    float4 v[kScoreMaxPerLane / 4];
    float part = 0.f;
#pragma unroll
    for (int t4 = 0; t4 < kScoreMaxPerLane / 4; ++t4) {
      if (on && 4 * t4 < per) {
        v[t4] = __ldg(row + t4);
        if (l1) {
          part += fabsf(v[t4].x - q[4 * t4]) + fabsf(v[t4].y - q[4 * t4 + 1]) + fabsf(v[t4].z - q[4 * t4 + 2]) + fabsf(v[t4].w - q[4 * t4 + 3]);
        } else {
          part = fmaf(v[t4].x, q[4 * t4], part);
          part = fmaf(v[t4].y, q[4 * t4 + 1], part);
          part = fmaf(v[t4].z, q[4 * t4 + 2], part);
          part = fmaf(v[t4].w, q[4 * t4 + 3], part);
        }
      }
    }
    part = score_reduce8(part);
    if (!on) continue;                                   // (after the shuffles: they need the whole warp)
    const float sc = l1 ? -part : part;
    const float coef = g * (expf(sc - lse) - (c == 0 ? 1.f : 0.f));
    float* grow = p.grad_table + row_off;
#pragma unroll
    for (int t4 = 0; t4 < kScoreMaxPerLane / 4; ++t4) {
      if (4 * t4 < per) {
        float4 d;
        if (l1) {  // score = -|row - q|_1:  d/d row = -sign(row - q),  d/d q = +sign(row - q)   (sign(0) = 0 like torch)
          const float sx = (v[t4].x > q[4 * t4]) - (v[t4].x < q[4 * t4]);
          const float sy = (v[t4].y > q[4 * t4 + 1]) - (v[t4].y < q[4 * t4 + 1]);
          const float sz = (v[t4].z > q[4 * t4 + 2]) - (v[t4].z < q[4 * t4 + 2]);
          const float sw = (v[t4].w > q[4 * t4 + 3]) - (v[t4].w < q[4 * t4 + 3]);
          d = make_float4(-coef * sx, -coef * sy, -coef * sz, -coef * sw);
          gq[4 * t4] = fmaf(coef, sx, gq[4 * t4]);
          gq[4 * t4 + 1] = fmaf(coef, sy, gq[4 * t4 + 1]);
          gq[4 * t4 + 2] = fmaf(coef, sz, gq[4 * t4 + 2]);
          gq[4 * t4 + 3] = fmaf(coef, sw, gq[4 * t4 + 3]);
        } else {   // score = <q, row>
          d = make_float4(coef * q[4 * t4], coef * q[4 * t4 + 1], coef * q[4 * t4 + 2], coef * q[4 * t4 + 3]);
          gq[4 * t4] = fmaf(coef, v[t4].x, gq[4 * t4]);
          gq[4 * t4 + 1] = fmaf(coef, v[t4].y, gq[4 * t4 + 1]);
          gq[4 * t4 + 2] = fmaf(coef, v[t4].z, gq[4 * t4 + 2]);
          gq[4 * t4 + 3] = fmaf(coef, v[t4].w, gq[4 * t4 + 3]);
        }
        atomicAdd(reinterpret_cast<float4*>(grow + 4 * t4), d);
      }
    }
  }
  // ---- the query's gradient: sum over the 4 candidate groups, then through q's definition ---------------------------------
#pragma unroll
  for (int t = 0; t < kScoreMaxPerLane; ++t) {
    if (t < per) {
      gq[t] += __shfl_xor_sync(0xffffffffu, gq[t], 8);
      gq[t] += __shfl_xor_sync(0xffffffffu, gq[t], 16);
    }
  }
  if (grp != 0) return;
  float* g_ent = p.grad_ent_embed + fixed_row;
  float* g_rel = p.grad_rel_embeds + static_cast<size_t>(r_id) * D;
#pragma unroll
  for (int t = 0; t < kScoreMaxPerLane; ++t) {
    if (t < per) {
      const int c = c0 + t;
      const float gv = gq[t];
      if (p.score_fn == TEMP_SCORE_COMPLEX) {
        const bool im = c >= half;
        const int k = im ? c - half : c;
        const float re_e = __ldg(fixed + k), im_e = __ldg(fixed + half + k);
        const float re_r = __ldg(rel + k), im_r = __ldg(rel + half + k);
        if (p.corrupt_tail) {   // q_re = re_e re_r - im_e im_r ; q_im = re_e im_r + im_e re_r
          if (!im) {
            atomicAdd(g_ent + k, gv * re_r);
            atomicAdd(g_ent + half + k, -gv * im_r);
            atomicAdd(g_rel + k, gv * re_e);
            atomicAdd(g_rel + half + k, -gv * im_e);
          } else {
            atomicAdd(g_ent + k, gv * im_r);
            atomicAdd(g_ent + half + k, gv * re_r);
            atomicAdd(g_rel + k, gv * im_e);
            atomicAdd(g_rel + half + k, gv * re_e);
          }
        } else {                // q_re = re_r re_e + im_r im_e ; q_im = re_r im_e - im_r re_e
          if (!im) {
            atomicAdd(g_ent + k, gv * re_r);
            atomicAdd(g_ent + half + k, gv * im_r);
            atomicAdd(g_rel + k, gv * re_e);
            atomicAdd(g_rel + half + k, gv * im_e);
          } else {
            atomicAdd(g_ent + k, -gv * im_r);
            atomicAdd(g_ent + half + k, gv * re_r);
            atomicAdd(g_rel + k, gv * im_e);
            atomicAdd(g_rel + half + k, -gv * re_e);
          }
        }
      } else if (p.score_fn == TEMP_SCORE_DISTMULT) {
        atomicAdd(g_ent + c, gv * __ldg(rel + c));
        atomicAdd(g_rel + c, gv * __ldg(fixed + c));
      } else {                  // q = s + r (tail) | o - r (head)
        atomicAdd(g_ent + c, gv);
        atomicAdd(g_rel + c, p.corrupt_tail ? gv : -gv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// filtered ranking: position of the target in the stable descending sort of sigmoid(masked scores)
// ------------------------------------------------------------------------------------------------
// grid (group of kRankQ queries, entity chunk): the CTA's warps walk the chunk's rows of the all-entity table once for
// the group's queries, 4 candidates per warp
// iteration exactly like the scorer above (8 lanes per candidate, the query vector in registers), and count the
// candidates that sort ahead of the target; every candidate is first counted with its real score, then the CTA of
// chunk 0 walks the query's filter list and replaces each filtered entity's contribution by that of sigmoid(-10e6)
// (= 0.f: ahead of the target only when the target's own sigmoid underflowed to 0 and the entity id is smaller).
// Counts are integers added with atomics, so the result does not depend on the schedule.
__device__ __forceinline__ float rank_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

constexpr int kRankQ = 4;   // queries that share one pass over the table rows (the pass is L2-bandwidth bound per query)

template <int PER4>
__global__ void __launch_bounds__(kThreads, PER4 <= 4 ? 2 : 1) rank_filtered_kernel(const TempRankArgs p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.d, per = D >> 3, half = D >> 1;
  const int grp = lane >> 3, sub = lane & 7;
  const int c0 = sub * per;
  const bool l1 = p.score_fn == TEMP_SCORE_TRANSE;
  // warp n < kRankQ forms query n's vector and the target's sigmoid once for the CTA; everybody picks them up from smem
  __shared__ __align__(16) float sq[kRankQ][8 * kScoreMaxPerLane];
  __shared__ float svt[kRankQ];
  __shared__ int stgt[kRankQ];
  if (warp < kRankQ) {
    const int qi = min(blockIdx.x * kRankQ + warp, p.n_query - 1);   // a short last group repeats its last query (not written)
    const long long s_id = p.triples[3 * qi], r_id = p.triples[3 * qi + 1], o_id = p.triples[3 * qi + 2];
    const float* fixed = p.ent_embed + static_cast<size_t>(p.corrupt_tail ? s_id : o_id) * D;
    const float* rel = p.rel_embeds + static_cast<size_t>(r_id) * D;
    float full[kScoreMaxPerLane];
    score_query(full, fixed, rel, c0, per, half, p.score_fn, p.corrupt_tail);
    float qn[4 * PER4];
#pragma unroll
    for (int t = 0; t < 4 * PER4; ++t) qn[t] = full[t];
    const int target = static_cast<int>(p.target[qi]);
    float4 v[PER4];
    score_load<PER4>(v, reinterpret_cast<const float4*>(p.table + static_cast<size_t>(target) * D + c0));
    const float tot = score_reduce8(score_dot<PER4>(v, qn, l1));
    if (grp == 0) {
#pragma unroll
      for (int t = 0; t < 4 * PER4; ++t) sq[warp][c0 + t] = qn[t];
    }
    if (lane == 0) {
      svt[warp] = rank_sigmoid(l1 ? -tot : tot);
      stgt[warp] = target;
    }
  }
  __syncthreads();
  float q[kRankQ][4 * PER4];
#pragma unroll
  for (int n = 0; n < kRankQ; ++n) {
#pragma unroll
    for (int t4 = 0; t4 < PER4; ++t4) {
      const float4 w = *reinterpret_cast<const float4*>(&sq[n][c0 + 4 * t4]);
      q[n][4 * t4] = w.x, q[n][4 * t4 + 1] = w.y, q[n][4 * t4 + 2] = w.z, q[n][4 * t4 + 3] = w.w;
    }
  }
  // after the 8-lane reduction every lane of a candidate's group holds all kRankQ totals: lane sub == n finishes query n
  // (sigmoid, comparison, count), so the kRankQ sigmoids of a candidate cost one pass instead of kRankQ
  static_assert(kRankQ <= 8, "one finishing lane per query");
  const int mine = sub < kRankQ ? sub : 0;
  const bool finisher = sub < kRankQ;
  const float vt = svt[mine];
  const int tgt = stgt[mine];
  int count = 0;
  const int chunk = (p.num_ents + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * chunk, hi = min(p.num_ents, lo + chunk);
  // two independent candidate groups per warp iteration: both groups' row loads are in flight before the first reduction
  constexpr int kRankUnroll = 2;
  constexpr int kStride = (kThreads / 32) * 4;
  for (int base = lo + warp * 4; base < hi; base += kStride * kRankUnroll) {
    float4 v[kRankUnroll][PER4];
#pragma unroll
    for (int u = 0; u < kRankUnroll; ++u) {
      const int j = min(base + u * kStride + grp, hi - 1);          // clamped: the load is always in range, the count is not taken
      score_load<PER4>(v[u], reinterpret_cast<const float4*>(p.table + static_cast<size_t>(j) * D + c0));
    }
#pragma unroll
    for (int u = 0; u < kRankUnroll; ++u) {
      const int j = base + u * kStride + grp;
      float tot = 0.f;
#pragma unroll
      for (int n = 0; n < kRankQ; ++n) {
        const float tn = score_reduce8(score_dot<PER4>(v[u], q[n], l1));
        tot = sub == n ? tn : tot;
      }
      if (finisher && j < hi && j != tgt) {
        const float val = rank_sigmoid(l1 ? -tot : tot);
        count += (val > vt || (val == vt && j < tgt)) ? 1 : 0;
      }
    }
  }
  if (blockIdx.y == 0 && p.filter_ptr != nullptr) {
#pragma unroll
    for (int n = 0; n < kRankQ; ++n) {
      const int qi = blockIdx.x * kRankQ + n;
      if (qi >= p.n_query) break;
      const float vtn = svt[n];
      const int tgn = stgt[n];
      const int f0 = p.filter_ptr[qi], f1 = p.filter_ptr[qi + 1];
      for (int base = f0 + warp * 4; base < f1; base += (kThreads / 32) * 4) {
        const int f = base + grp;
        const int j = f < f1 ? p.filter_ids[f] : -1;
        float4 v[PER4];
        score_load<PER4>(v, reinterpret_cast<const float4*>(p.table + static_cast<size_t>(max(j, 0)) * D + c0));
        const float tot = score_reduce8(score_dot<PER4>(v, q[n], l1));
        if (j >= 0 && j != tgn && sub == n) {                   // counted on query n's finishing lane
          const float val = rank_sigmoid(l1 ? -tot : tot);
          const int real = (val > vtn || (val == vtn && j < tgn)) ? 1 : 0;
          const int masked = (0.f == vtn && j < tgn) ? 1 : 0;    // sigmoid(-10e6) == 0 exactly; 0 > vt never holds
          count += masked - real;
        }
      }
    }
  }
  count += __shfl_xor_sync(0xffffffffu, count, 8);              // the four candidate groups of the warp
  count += __shfl_xor_sync(0xffffffffu, count, 16);
  const int qi = blockIdx.x * kRankQ + lane;
  if (lane < kRankQ && qi < p.n_query) {
    if (blockIdx.y == 0 && warp == 0) count += 1;               // 1-indexed
    if (count != 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.rank + qi), static_cast<unsigned long long>(static_cast<long long>(count)));
  }
}

// ------------------------------------------------------------------------------------------------
// cross-GPU barrier of the fused all-gather (signal every peer, wait for every peer)
// ------------------------------------------------------------------------------------------------
__global__ void peer_barrier_kernel(uint32_t* flags_local, uint32_t* const* flag_peers, int world, int rank, uint32_t seq) {
  const int k = threadIdx.x;
  if (k >= world) return;
  uint32_t* remote = flag_peers[k] + rank;   // (kernel parameters and this table are not written by the predecessor)
  // launched with programmatic stream serialization: resident while the step's last kernel still runs, so the signal
  // leaves as soon as that grid has completed and flushed instead of one launch latency later
  // (and the next step's first kernel may become resident behind it: every kernel of this library reads predecessor data
  // only after its own griddepcontrol.wait, which waits for THIS grid's completion, i.e. for all peers' signals)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(seq) : "memory");
  const uint32_t* mine = flags_local + k;
  uint32_t v;
  unsigned spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (++spins > (1u << 28)) __trap();  // a peer that never arrives is a bug in the caller: do not hang the device
  } while (static_cast<int32_t>(v - seq) < 0);
}

// ------------------------------------------------------------------------------------------------
// small data-movement kernels
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const TempGatherArgs p) {
  const int nvec = p.d >> 2;
  const size_t total = static_cast<size_t>(p.n) * nvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx / nvec), c4 = static_cast<int>(idx - static_cast<size_t>(i) * nvec);
    const int s = __ldg(p.index + i);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= 0) v = ldg4(p.table + static_cast<size_t>(s) * p.d + c4 * 4);
    *reinterpret_cast<float4*>(p.out + static_cast<size_t>(i) * p.d + c4 * 4) = v;
  }
}

__global__ void scatter_rows_kernel(const TempScatterArgs p) {
  const int nvec = p.d >> 2;
  const size_t total = static_cast<size_t>(p.n) * nvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx / nvec), c4 = static_cast<int>(idx - static_cast<size_t>(i) * nvec);
    const int s = p.src_index != nullptr ? __ldg(p.src_index + i) : i;
    const int t = p.dst_index != nullptr ? __ldg(p.dst_index + i) : i;
    if (s < 0 || t < 0) continue;
    float4 v = ldg4(p.src + static_cast<size_t>(s) * p.d + c4 * 4);
    if (p.add_row != nullptr) {
      const float4 a = ldg4(p.add_row + c4 * 4);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(p.dst + static_cast<size_t>(t) * p.d + c4 * 4) = v;
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out,
                                 int out_ld) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = by + i, c = bx + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[static_cast<size_t>(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = bx + i, r = by + threadIdx.x;
    if (c < cols && r < rows) out[static_cast<size_t>(c) * out_ld + r] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_d(int d) {
  if (d <= 0 || d > TEMP_MAX_D || (d & 3)) return fail(TEMP_EINVAL, "d must be a multiple of 4 in (0, %s%ld]", "", TEMP_MAX_D);
  return TEMP_OK;
}

template <int Tag, typename K>
int ensure_smem(K kernel, size_t bytes, const char* name) {
  static size_t configured = 0;  // one instance per Tag (= per kernel); one device per process
  if (bytes <= configured) return TEMP_OK;
  int dev = 0, lim = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (bytes > static_cast<size_t>(lim)) return fail(TEMP_EUNSUPPORTED, "%s needs %ld B shared memory", name, static_cast<long>(bytes));
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return cuda_fail(e, name);
  configured = bytes;
  return TEMP_OK;
}

int launch_score_loss_bwd(const TempScoreLossBwdArgs* a, cudaStream_t st) {
  if (a == nullptr || a->n_pos < 0 || a->n_cand <= 0) return fail(TEMP_EINVAL, "bad score args%s", "");
  if (a->n_pos == 0) return TEMP_OK;
  if (a->d <= 0 || (a->d & 31) || a->d / 8 > kScoreMaxPerLane) return fail(TEMP_EUNSUPPORTED, "score: d must be a multiple of 32 up to %s%ld", "", 8L * kScoreMaxPerLane);
  if (a->score_fn < TEMP_SCORE_DISTMULT || a->score_fn > TEMP_SCORE_TRANSE) return fail(TEMP_EINVAL, "bad score_fn%s", "");
  if (!a->ent_embed || !a->rel_embeds || !a->table || !a->triples || !a->cand || !a->grad_loss || !a->grad_ent_embed ||
      !a->grad_rel_embeds || !a->grad_table)
    return fail(TEMP_EINVAL, "score backward has null pointers%s", "");
  if (!aligned16(a->table) || !aligned16(a->grad_table)) return fail(TEMP_EINVAL, "score table / grad_table misaligned%s", "");
  const int per_cta = kThreads / 32;
  score_loss_bwd_kernel<<<(a->n_pos + per_cta - 1) / per_cta, kThreads, 0, st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "score_loss_bwd_kernel launch");
  return TEMP_OK;
}

// SM count of the current device (148 on a B200), looked up once per process (one CUDA device per process)
int sm_count_cached() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      return 148;
  }
  return sms;
}

int launch_layer(const TempRgcnLayerArgs* a, cudaStream_t st) {
  if (a == nullptr) return fail(TEMP_EINVAL, "null args%s", "");
  if (int rc = check_d(a->d)) return rc;
  const int rows = a->row1 - a->row0;
  if (rows < 0) return fail(TEMP_EINVAL, "row1 < row0%s", "");
  if (rows == 0) return TEMP_OK;
  if (a->n_terms < 0 || a->n_terms > TEMP_MAX_TERMS) return fail(TEMP_EINVAL, "n_terms out of range%s", "");
  if (a->row_ptr != nullptr) {
    if (!a->e_src || !a->e_rel || !a->norm || !a->x || !a->weight) return fail(TEMP_EINVAL, "graph part has null pointers%s", "");
    if (a->n_bases <= 0 || a->si <= 0 || a->so <= 0 || a->n_bases * a->so != a->d || a->n_bases * a->si != a->d)
      return fail(TEMP_EINVAL, "n_bases*si and n_bases*so must equal d%s", "");
    if (!aligned16(a->x) || !aligned16(a->weight)) return fail(TEMP_EINVAL, "x / weight must be 16-byte aligned%s", "");
  }
  for (int t = 0; t < a->n_terms; ++t) {
    if (!a->terms[t].a || !a->terms[t].w) return fail(TEMP_EINVAL, "dense term %s%ld has null pointers", "", t);
    if (!aligned16(a->terms[t].a) || !aligned16(a->terms[t].w)) return fail(TEMP_EINVAL, "dense term %s%ld misaligned", "", t);
  }
  if (a->residual && a->n_terms == 0) return fail(TEMP_EINVAL, "residual needs term 0%s", "");
  if ((a->te_out || a->te_chain) && a->time_embed == nullptr) return fail(TEMP_EINVAL, "time_embed is null%s", "");
  if (a->chain_w != nullptr) {
    if (a->chain_out == nullptr || a->chain_n <= 0 || (a->chain_n & 3) || a->chain_ld < a->chain_n || (a->chain_ld & 3))
      return fail(TEMP_EINVAL, "bad chain output%s", "");
    if (!aligned16(a->chain_w) || !aligned16(a->chain_out)) return fail(TEMP_EINVAL, "chain buffers misaligned%s", "");
  } else if (a->h_out == nullptr) {
    return fail(TEMP_EINVAL, "layer has no output%s", "");
  }
  if (a->h_out != nullptr && !aligned16(a->h_out)) return fail(TEMP_EINVAL, "h_out misaligned%s", "");
  if (temp_internal::tc_layer_supported(a)) return temp_internal::tc_launch_layer(a, st);
  if (a->chain_peers != nullptr) return fail(TEMP_EUNSUPPORTED, "peer stores of the chained output need the tcgen05 layer (d == 128)%s", "");
  const int Kp = (a->d + kKC - 1) / kKC * kKC;
  const size_t smem = (static_cast<size_t>(kTM) * (Kp + 4) * (a->chain_w ? 2 : 1) + 2 * kKC * kNC) * sizeof(float);
  if (int rc = ensure_smem<0>(rgcn_layer_kernel, smem, "rgcn_layer_kernel")) return rc;
  const int grid = (rows + kTM - 1) / kTM;
  int gy = 1;
  if (a->chain_w != nullptr) {   // few row tiles (the Bi centre step, the last steps of a window): fill the SMs with column blocks
    const int chunks2 = (a->chain_n + kNC - 1) / kNC;
    gy = sm_count_cached() / grid;
    if (gy > chunks2) gy = chunks2;
    if (gy < 1) gy = 1;
  }
  rgcn_layer_kernel<<<dim3(grid, gy), kThreads, smem, st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "rgcn_layer_kernel launch");
  return TEMP_OK;
}

int launch_gru(const TempGruArgs* a, cudaStream_t st) {
  if (a == nullptr) return fail(TEMP_EINVAL, "null args%s", "");
  if (int rc = check_d(a->d)) return rc;
  const int rows = a->row1 - a->row0;
  if (rows < 0) return fail(TEMP_EINVAL, "row1 < row0%s", "");
  if (rows == 0) return TEMP_OK;
  if (!a->gi || !a->whh_t || !a->b_hh || !a->out) return fail(TEMP_EINVAL, "gru has null pointers%s", "");
  if (a->prev_row != nullptr && a->state == nullptr) return fail(TEMP_EINVAL, "prev_row without state%s", "");
  if (!aligned16(a->whh_t) || (a->state && !aligned16(a->state))) return fail(TEMP_EINVAL, "gru buffers misaligned%s", "");
  if (a->cell_type != TEMP_CELL_TORCH_GRU && a->cell_type != TEMP_CELL_TYPE1) return fail(TEMP_EINVAL, "bad cell_type%s", "");
  if (a->d == 128 && (a->prev_row == nullptr || a->whh_packed != nullptr)) {
    TempGruScanArgs one;
    memset(&one, 0, sizeof(one));
    one.n_steps = 1;
    one.steps[0] = *a;
    if (temp_internal::tc_scan_supported(&one)) return temp_internal::tc_launch_scan(&one, st);
  }
  if (a->d != 128 && temp_internal::tcw_gru_supported(a)) return temp_internal::tcw_launch_gru(a, st);
  const int jblocks = (a->d + kGJ - 1) / kGJ;
  cudaError_t e;
  // small steps: 32-row tiles so that the step still spreads over all SMs
  if (rows <= sm_count_cached() * 32) {
    const size_t smem = (static_cast<size_t>(32) * (a->d + 4) + static_cast<size_t>(a->d) * 3 * kGJ) * sizeof(float);
    if (int rc = ensure_smem<1>(gru_kernel<4>, smem, "gru_kernel<4>")) return rc;
    dim3 grid((rows + 31) / 32, jblocks);
    gru_kernel<4><<<grid, kThreads, smem, st>>>(*a);
  } else {
    const size_t smem = (static_cast<size_t>(64) * (a->d + 4) + static_cast<size_t>(a->d) * 3 * kGJ) * sizeof(float);
    if (int rc = ensure_smem<2>(gru_kernel<8>, smem, "gru_kernel<8>")) return rc;
    dim3 grid((rows + 63) / 64, jblocks);
    gru_kernel<8><<<grid, kThreads, smem, st>>>(*a);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gru_kernel launch");
  return TEMP_OK;
}

int check_gru(const TempGruArgs* a) {
  if (int rc = check_d(a->d)) return rc;
  if (a->row1 < a->row0) return fail(TEMP_EINVAL, "row1 < row0%s", "");
  if (!a->gi || !a->whh_t || !a->b_hh || !a->out) return fail(TEMP_EINVAL, "gru has null pointers%s", "");
  if (a->prev_row != nullptr && a->state == nullptr) return fail(TEMP_EINVAL, "prev_row without state%s", "");
  if (!aligned16(a->whh_t) || (a->state && !aligned16(a->state))) return fail(TEMP_EINVAL, "gru buffers misaligned%s", "");
  if (a->cell_type != TEMP_CELL_TORCH_GRU && a->cell_type != TEMP_CELL_TYPE1) return fail(TEMP_EINVAL, "bad cell_type%s", "");
  return TEMP_OK;
}

template <int RPW, int Tag>
int launch_scan_t(const TempGruScanArgs* a, int max_rows, cudaStream_t st) {
  const int D = a->steps[0].d;
  const int TM = RPW * 8;
  const size_t smem = (static_cast<size_t>(TM) * (D + 4) + static_cast<size_t>(D) * 3 * kGJ) * sizeof(float);
  if (int rc = ensure_smem<Tag>(gru_scan_kernel<RPW>, smem, "gru_scan_kernel")) return rc;
  static int sm_count = 0, per_sm = 0;
  static size_t occ_smem = 0;
  if (sm_count == 0 || occ_smem != smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gru_scan_kernel<RPW>, kThreads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "occupancy query");
    occ_smem = smem;
  }
  if (per_sm < 1) return fail(TEMP_EUNSUPPORTED, "gru_scan_kernel does not fit on an SM%s", "");
  const int JB = (D + kGJ - 1) / kGJ;
  int want = ((max_rows + TM - 1) / TM) * JB;
  int cap = (sm_count * per_sm) / JB * JB;
  int grid = want < cap ? want : cap;
  if (grid < JB) grid = JB;
  void* params[] = {const_cast<TempGruScanArgs*>(a)};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gru_scan_kernel<RPW>), dim3(grid), dim3(kThreads),
                                              params, smem, st);
  if (e != cudaSuccess) return cuda_fail(e, "gru_scan_kernel launch");
  return TEMP_OK;
}

int launch_scan(const TempGruScanArgs* a, cudaStream_t st) {
  if (a == nullptr) return fail(TEMP_EINVAL, "null args%s", "");
  if (a->n_steps < 0 || a->n_steps > TEMP_MAX_SCAN_STEPS) return fail(TEMP_EINVAL, "n_steps out of range%s", "");
  if (a->n_steps == 0) return TEMP_OK;
  if (a->parts != nullptr && (a->n_parts < 0 || a->part_stride <= 0)) return fail(TEMP_EINVAL, "bad chain-partition table%s", "");
  int max_rows = 0;
  for (int s = 0; s < a->n_steps; ++s) {
    if (int rc = check_gru(&a->steps[s])) return rc;
    if (a->steps[s].d != a->steps[0].d) return fail(TEMP_EINVAL, "all scan steps must share d%s", "");
    const int rows = a->steps[s].row1 - a->steps[s].row0;
    if (rows > max_rows) max_rows = rows;
  }
  if (max_rows == 0) return TEMP_OK;
  if (temp_internal::tc_scan_supported(a)) return temp_internal::tc_launch_scan(a, st);
  if (a->steps[0].d != 128 && temp_internal::tcw_scan_supported(a)) return temp_internal::tcw_launch_scan(a, st);
  if (a->push_bufs != nullptr) return fail(TEMP_EUNSUPPORTED, "the fused peer all-gather needs the tcgen05 scan (d == 128, partition table)%s", "");
  if (a->barrier == nullptr) return fail(TEMP_EINVAL, "scan needs a zero-initialised 8-byte barrier word%s", "");
  if (max_rows <= sm_count_cached() * 64) return launch_scan_t<4, 3>(a, max_rows, st);
  return launch_scan_t<8, 4>(a, max_rows, st);
}

int launch_attn(const TempAttnArgs* a, cudaStream_t st) {
  if (a == nullptr) return fail(TEMP_EINVAL, "null args%s", "");
  if (int rc = check_d(a->d)) return rc;
  const int rows = a->row1 - a->row0;
  if (rows < 0) return fail(TEMP_EINVAL, "row1 < row0%s", "");
  if (rows == 0) return TEMP_OK;
  if (a->heads <= 0 || 32 % a->heads != 0 || a->d % a->heads != 0) return fail(TEMP_EINVAL, "heads must divide 32 and d%s", "");
  const int dk = a->d / a->heads, lph = 32 / a->heads;
  if ((dk + lph - 1) / lph > kAttnMaxT) return fail(TEMP_EUNSUPPORTED, "head dim too large%s", "");
  if (!a->qkv || !a->out || a->n_slots < 0) return fail(TEMP_EINVAL, "attention has null pointers%s", "");
  if (a->n_slots > 0 && (!a->kv_hist || !a->slot_row)) return fail(TEMP_EINVAL, "attention history is null%s", "");
  if (a->decay_wb != nullptr && a->tau == nullptr) return fail(TEMP_EINVAL, "decay without tau%s", "");
  const int rows_per_cta = kThreads / 32;
  attn_kernel<<<(rows + rows_per_cta - 1) / rows_per_cta, kThreads, 0, st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "attn_kernel launch");
  return TEMP_OK;
}

int launch_score_loss(const TempScoreLossArgs* a, cudaStream_t st) {
  if (a == nullptr || a->n_pos < 0 || a->n_cand <= 0) return fail(TEMP_EINVAL, "bad score args%s", "");
  if (a->n_pos == 0) return TEMP_OK;
  if (a->d <= 0 || (a->d & 31) || a->d / 8 > kScoreMaxPerLane) return fail(TEMP_EUNSUPPORTED, "score: d must be a multiple of 32 up to %s%ld", "", 8L * kScoreMaxPerLane);
  if (a->score_fn < TEMP_SCORE_DISTMULT || a->score_fn > TEMP_SCORE_TRANSE) return fail(TEMP_EINVAL, "bad score_fn%s", "");
  if (!a->ent_embed || !a->rel_embeds || !a->table || !a->triples || !a->cand || !a->loss) return fail(TEMP_EINVAL, "score has null pointers%s", "");
  if (!aligned16(a->table)) return fail(TEMP_EINVAL, "score table misaligned%s", "");
  const int per_cta = kThreads / 32;
  score_loss_kernel<<<(a->n_pos + per_cta - 1) / per_cta, kThreads, 0, st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "score_loss_kernel launch");
  return TEMP_OK;
}

int launch_rank_filtered(const TempRankArgs* a, cudaStream_t st) {
  if (a == nullptr || a->n_query < 0 || a->num_ents <= 0) return fail(TEMP_EINVAL, "bad rank args%s", "");
  if (a->n_query == 0) return TEMP_OK;
  if (a->d <= 0 || (a->d & 31) || a->d / 8 > kScoreMaxPerLane) return fail(TEMP_EUNSUPPORTED, "rank: d must be a multiple of 32 up to %s%ld", "", 8L * kScoreMaxPerLane);
  if (a->score_fn < TEMP_SCORE_DISTMULT || a->score_fn > TEMP_SCORE_TRANSE) return fail(TEMP_EINVAL, "bad score_fn%s", "");
  if (!a->ent_embed || !a->rel_embeds || !a->table || !a->triples || !a->target || !a->rank) return fail(TEMP_EINVAL, "rank has null pointers%s", "");
  if ((a->filter_ptr == nullptr) != (a->filter_ids == nullptr)) return fail(TEMP_EINVAL, "rank: filter_ptr and filter_ids go together%s", "");
  if (!aligned16(a->table)) return fail(TEMP_EINVAL, "rank table misaligned%s", "");
  cudaError_t e = cudaMemsetAsync(a->rank, 0, sizeof(int64_t) * static_cast<size_t>(a->n_query), st);
  if (e != cudaSuccess) return cuda_fail(e, "rank memset");
  // entity chunks: about two CTAs per SM over the whole grid, at least 32 table rows per warp
  const int groups = (a->n_query + kRankQ - 1) / kRankQ;
  int chunks = (sm_count_cached() * 2) / groups;           // one wave of two CTAs per SM: never a few CTAs left for a second
  const int max_chunks = (a->num_ents + 32 * (kThreads / 32) - 1) / (32 * (kThreads / 32));
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const dim3 grid(groups, chunks);
  switch (a->d / 32) {
    case 1: rank_filtered_kernel<1><<<grid, kThreads, 0, st>>>(*a); break;
    case 2: rank_filtered_kernel<2><<<grid, kThreads, 0, st>>>(*a); break;
    case 3: rank_filtered_kernel<3><<<grid, kThreads, 0, st>>>(*a); break;
    case 4: rank_filtered_kernel<4><<<grid, kThreads, 0, st>>>(*a); break;
    case 5: rank_filtered_kernel<5><<<grid, kThreads, 0, st>>>(*a); break;
    case 6: rank_filtered_kernel<6><<<grid, kThreads, 0, st>>>(*a); break;
    case 7: rank_filtered_kernel<7><<<grid, kThreads, 0, st>>>(*a); break;
    default: rank_filtered_kernel<8><<<grid, kThreads, 0, st>>>(*a); break;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "rank_filtered_kernel launch");
  return TEMP_OK;
}

int grid_for(size_t work_items) {
  size_t g = (work_items + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count_cached()) * 8;
  if (g > cap) g = cap;
  return static_cast<int>(g == 0 ? 1 : g);
}

int launch_gather(const TempGatherArgs* a, cudaStream_t st) {
  if (a == nullptr || a->n < 0) return fail(TEMP_EINVAL, "bad gather args%s", "");
  if (a->n == 0) return TEMP_OK;
  if (a->d <= 0 || (a->d & 3)) return fail(TEMP_EINVAL, "gather width must be a positive multiple of 4%s", "");
  if (!a->table || !a->index || !a->out) return fail(TEMP_EINVAL, "gather has null pointers%s", "");
  gather_rows_kernel<<<grid_for(static_cast<size_t>(a->n) * (a->d / 4)), 256, 0, st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gather_rows_kernel launch");
  return TEMP_OK;
}

int launch_scatter(const TempScatterArgs* a, cudaStream_t st) {
  if (a == nullptr || a->n < 0) return fail(TEMP_EINVAL, "bad scatter args%s", "");
  if (a->n == 0) return TEMP_OK;
  if (a->d <= 0 || (a->d & 3)) return fail(TEMP_EINVAL, "scatter width must be a positive multiple of 4%s", "");
  if (!a->src || !a->dst) return fail(TEMP_EINVAL, "scatter has null pointers%s", "");
  scatter_rows_kernel<<<grid_for(static_cast<size_t>(a->n) * (a->d / 4)), 256, 0, st>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "scatter_rows_kernel launch");
  return TEMP_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int temp_abi_version(void) { return TEMP_ABI_VERSION; }

const char* temp_last_error_string(void) { return g_err; }

int temp_device_info(int32_t* sm_count, int32_t* max_smem, int32_t* cc) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  int sm = 0, smem = 0, major = 0, minor = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (sm_count) *sm_count = sm;
  if (max_smem) *max_smem = smem;
  if (cc) *cc = major * 10 + minor;
  return TEMP_OK;
}

int temp_rgcn_layer_fwd(const TempRgcnLayerArgs* args, void* stream) {
  return launch_layer(args, static_cast<cudaStream_t>(stream));
}

int temp_gru_fwd(const TempGruArgs* args, void* stream) { return launch_gru(args, static_cast<cudaStream_t>(stream)); }

int temp_rgcn_gather_fwd(const TempRgcnLayerArgs* args, void* stream) {
  if (args == nullptr) return fail(TEMP_EINVAL, "null args%s", "");
  if (args->d <= 0 || (args->d & 3) != 0 || args->d > 256 || args->row_ptr == nullptr || args->si != args->so ||
      (args->si != 1 && args->si != 2 && args->si != 4) || args->n_bases * args->si != args->d || args->agg_scratch == nullptr ||
      args->x == nullptr || args->weight == nullptr || args->norm == nullptr || args->e_src == nullptr || args->e_rel == nullptr)
    return fail(TEMP_EUNSUPPORTED, "temp_rgcn_gather_fwd: d %% 4 == 0, d <= 256, 1x1 / 2x2 / 4x4 relation blocks and an aggregate buffer are required%s", "");
  return temp_internal::tc_launch_gather(args, static_cast<cudaStream_t>(stream));
}

int temp_gru_scan_fwd(const TempGruScanArgs* args, void* stream) {
  return launch_scan(args, static_cast<cudaStream_t>(stream));
}

int temp_attention_fwd(const TempAttnArgs* args, void* stream) {
  return launch_attn(args, static_cast<cudaStream_t>(stream));
}

int temp_gather_rows(const TempGatherArgs* args, void* stream) {
  return launch_gather(args, static_cast<cudaStream_t>(stream));
}

int temp_scatter_rows(const TempScatterArgs* args, void* stream) {
  return launch_scatter(args, static_cast<cudaStream_t>(stream));
}

int temp_score_loss_bwd(const TempScoreLossBwdArgs* args, void* stream) {
  return launch_score_loss_bwd(args, static_cast<cudaStream_t>(stream));
}

int temp_rank_filtered_fwd(const TempRankArgs* args, void* stream) {
  return launch_rank_filtered(args, static_cast<cudaStream_t>(stream));
}

int temp_score_loss_fwd(const TempScoreLossArgs* args, void* stream) {
  return launch_score_loss(args, static_cast<cudaStream_t>(stream));
}

int temp_transpose(const float* in, int32_t rows, int32_t cols, float* out, int32_t out_ld, void* stream) {
  if (!in || !out || rows <= 0 || cols <= 0 || out_ld < rows) return fail(TEMP_EINVAL, "bad transpose args%s", "");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(in, rows, cols, out, out_ld);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "transpose_kernel launch");
  return TEMP_OK;
}

int64_t temp_packed_weights_bytes(int32_t k, int32_t n) {
  if (k != 128 || n <= 0 || (n & 127)) return temp_internal::tcw_packed_bytes(k, n);
  return static_cast<int64_t>(n / 128) * 4 * 2 * 128 * 128;
}

int temp_pack_weights(const float* w_kn, int32_t k, int32_t n, void* packed, void* stream) {
  return temp_internal::tc_pack_weights(w_kn, k, n, packed, static_cast<cudaStream_t>(stream));
}

int64_t temp_packed_gru_bytes(int32_t d) {
  return d == 128 ? static_cast<int64_t>(4) * 4 * 2 * 128 * 128 : temp_internal::tcw_packed_gru_bytes(d);
}

int temp_pack_gru_weights(const float* whh_t, int32_t d, void* packed, void* stream) {
  return temp_internal::tc_pack_gru_weights(whh_t, d, packed, static_cast<cudaStream_t>(stream));
}

int temp_program_kernel_count(const TempOp* ops, int32_t n) {
  if (ops == nullptr || n < 0) return TEMP_EINVAL;
  int k = 0;
  for (int i = 0; i < n; ++i) {
    switch (ops[i].kind) {
      case TEMP_OP_LAYER: {
        const TempRgcnLayerArgs& a = ops[i].u.layer;
        if (a.row1 > a.row0)
          k += 1 + (temp_internal::tc_layer_supported(&a) ? temp_internal::tc_gather_launches(&a) : 0);
        break;
      }
      case TEMP_OP_GRU: k += ops[i].u.gru.row1 > ops[i].u.gru.row0 ? 1 : 0; break;
      case TEMP_OP_GRU_SCAN: {
        const TempGruScanArgs& sc = ops[i].u.scan;
        if (sc.n_steps > 0 && sc.n_steps <= TEMP_MAX_SCAN_STEPS && sc.steps[0].d != 128 && !temp_internal::tc_scan_supported(&sc) &&
            temp_internal::tcw_scan_supported(&sc)) {   // one cooperative launch, or one gru_step_tcw_kernel launch per non-empty step
          k += temp_internal::tcw_scan_launches(&sc);
        } else {
          k += sc.n_steps > 0 ? 1 : 0;
        }
        break;
      }
      case TEMP_OP_ATTN: k += ops[i].u.attn.row1 > ops[i].u.attn.row0 ? 1 : 0; break;
      case TEMP_OP_GATHER: k += ops[i].u.gather.n > 0 ? 1 : 0; break;
      case TEMP_OP_SCATTER: k += ops[i].u.scatter.n > 0 ? 1 : 0; break;
      case TEMP_OP_MEMCPY_H2D:
      case TEMP_OP_MEMCPY_D2H: break;
      default: return TEMP_EINVAL;
    }
  }
  return k;
}

int temp_run_program(const TempOp* ops, int32_t n, void* stream) {
  if (ops == nullptr || n < 0) return fail(TEMP_EINVAL, "bad program%s", "");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n; ++i) {
    int rc = TEMP_OK;
    switch (ops[i].kind) {
      case TEMP_OP_LAYER: rc = launch_layer(&ops[i].u.layer, st); break;
      case TEMP_OP_GRU: rc = launch_gru(&ops[i].u.gru, st); break;
      case TEMP_OP_GRU_SCAN: rc = launch_scan(&ops[i].u.scan, st); break;
      case TEMP_OP_ATTN: rc = launch_attn(&ops[i].u.attn, st); break;
      case TEMP_OP_GATHER: rc = launch_gather(&ops[i].u.gather, st); break;
      case TEMP_OP_SCATTER: rc = launch_scatter(&ops[i].u.scatter, st); break;
      case TEMP_OP_MEMCPY_H2D:
      case TEMP_OP_MEMCPY_D2H: {
        if (ops[i].u.copy.bytes == 0) break;
        cudaError_t e = cudaMemcpyAsync(ops[i].u.copy.dst, ops[i].u.copy.src, ops[i].u.copy.bytes,
                                        ops[i].kind == TEMP_OP_MEMCPY_H2D ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync");
        break;
      }
      default: rc = fail(TEMP_EINVAL, "unknown op kind at index %s%ld", "", i);
    }
    if (rc != TEMP_OK) return rc;
  }
  return TEMP_OK;
}

int temp_peer_barrier(uint32_t* flags_local, uint32_t* const* flag_peers, int32_t world, int32_t rank, uint32_t seq,
                      void* stream) {
  if (flags_local == nullptr || flag_peers == nullptr || world <= 0 || world > 32 || rank < 0 || rank >= world)
    return fail(TEMP_EINVAL, "bad peer barrier args%s", "");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(32);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, peer_barrier_kernel, flags_local, flag_peers, world, rank, seq);
  if (e != cudaSuccess) return cuda_fail(e, "peer_barrier_kernel launch");
  return TEMP_OK;
}

int temp_graph_create(const TempOp* ops, int32_t n, void** graph_exec_out) {
  if (ops == nullptr || n <= 0 || graph_exec_out == nullptr) return fail(TEMP_EINVAL, "bad graph args%s", "");
  *graph_exec_out = nullptr;
  cudaStream_t cap = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    cudaStreamDestroy(cap);
    return cuda_fail(e, "cudaStreamBeginCapture");
  }
  const int rc = temp_run_program(ops, n, cap);
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(cap, &graph);
  cudaStreamDestroy(cap);
  if (rc != TEMP_OK) {
    if (graph != nullptr) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc;
  }
  if (e != cudaSuccess || graph == nullptr) return cuda_fail(e, "cudaStreamEndCapture");
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
  *graph_exec_out = exec;
  return TEMP_OK;
}

int temp_graph_launch(void* graph_exec, void* stream) {
  if (graph_exec == nullptr) return fail(TEMP_EINVAL, "null graph%s", "");
  cudaError_t e = cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
  return TEMP_OK;
}

int temp_graph_destroy(void* graph_exec) {
  if (graph_exec == nullptr) return TEMP_OK;
  cudaError_t e = cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec));
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphExecDestroy");
  return TEMP_OK;
}

}  // extern "C"
