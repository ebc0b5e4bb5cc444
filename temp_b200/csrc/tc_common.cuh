// temp_b200 -- sm_100a primitives shared by the tensor-core kernels: mbarrier, 1-D bulk copies (TMA engine),
// tcgen05 (TMEM allocation, UMMA descriptors, MMA issue, commit, TMEM loads) and the 3xTF32 operand split.
//
// Operand convention of every tcgen05 GEMM in this library ("features on TMEM lanes"):
//     D^T[m, n] += sum_k  A[m, k] * B[n, k]
//   A : weights, 128 output features x K, K-major, pre-packed on the device once per parameter version
//       (pack_weights_kernel) in the exact shared-memory image, fetched with cp.async.bulk;
//   B : activations, n = packed row of the tile, K-major, written by the CTA's threads with st.shared;
//   D : fp32 accumulators in TMEM, lane = output feature, column = packed row, so that for a fixed row the 32
//       lanes of a warp hold 32 consecutive features: every global load/store of activations is a coalesced
//       128-byte line and no transposition is ever staged through shared memory.
// Both operands use the canonical K-major SWIZZLE_128B layout: a "k-atom" is 32 fp32 (128 B) per row, rows in
// groups of 8 (1024 B, the swizzle period), k-atoms of one operand `rows * 128` bytes apart.
//
// fp32 parity (1e-4 relative, north star): every product is computed as  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  with
// hi = the top 11 significand bits (what kind::tf32 keeps), lo = a - hi (exact); the dropped a_lo*b_lo term and
// the tf32 rounding of the lo parts are O(2^-22) relative.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// high word of every K-major SWIZZLE_128B UMMA descriptor of this library (SBO 1024 B, version 1, layout type 2)
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
  return ok != 0;
}
// Bounded spin: a barrier that never completes is a bug in this library; trap instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// ---- 1-D bulk copy global -> shared (TMA engine, completes on an mbarrier) -----------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- programmatic dependent launch -------------------------------------------------------------------------
// The kernels of a forward are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel lets its
// successor start launching at once (pdl_launch_dependents at entry: the successor's CTAs become eligible once every
// CTA of this grid has started, i.e. they never starve it) and does everything that does not depend on the
// predecessor's output -- barrier init, TMEM allocation, weight prefetch, plan indices -- before pdl_wait(), which
// returns when the predecessor grid has completed and its memory is visible.  No global write precedes pdl_wait().
// Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Loads of data the PREDECESSOR kernel produced (activations, aggregates, input gates).  With programmatic dependent launch
// this kernel is already running while the predecessor still writes them, so they are NOT read-only for the kernel's
// lifetime: __ldg (ld.global.nc) tells the compiler exactly that, and it may then hoist such a load above
// griddepcontrol.wait -- seen in round 2 as the first self-loop row of a warp read before the previous layer had stored
// it, depending on nothing but instruction scheduling.  These are volatile asm statements: never reordered with the
// (volatile) wait.  Parameters and plan arrays, which no kernel of the program writes, keep __ldg.
__device__ __forceinline__ float4 ld_dep_f32x4(const void* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_dep_f32(const void* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers of the calling thread (lane = warp's quadrant lane).
// The tcgen05.wait::ld is part of the SAME asm statement as the load: as a separate statement it has no data dependency on
// the destination registers, and the compiler is free to schedule their first use ahead of it (seen in round 2: the first
// row of every 16-row epilogue chunk read stale registers once an unrelated change altered the schedule).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one fp32 per lane -> column `taddr` of the warp's 32 TMEM lanes (warp-collective)
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
// 8 consecutive columns (v[0..7]) of the warp's 32 TMEM lanes (warp-collective)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// explicit shared-space accesses (pointers derived from the aligned dynamic-smem base are generic to the compiler)
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_f32x4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}

// ---- UMMA descriptors ----------------------------------------------------------------------------------
// K-major SWIZZLE_128B operand tile whose first k-atom starts at `saddr` (1024-byte aligned): 8-row groups are
// 1024 B apart (SBO); LBO is unused for swizzled K-major layouts.  Bits: [0,14) addr>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version = 1 (sm_100), [61,64) layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major.  Bits: [4,6) D format 1 = f32, [7,10) A format 2 = tf32,
// [10,13) B format 2 = tf32, [15] A major, [16] B major (0 = K), [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// one elected thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread complete -> one arrival on `bar` (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand layout ----------------------------------------------------------------------------------------
// byte offset of element (row r, k in [0, 32)) inside one k-atom block of a K-major SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t k) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((k >> 2) ^ (r & 7u)) << 4) | ((k & 3u) << 2));
}

__host__ __device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
  hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
#else
  union { float f; uint32_t u; } c;
  c.f = a;
  c.u &= 0xffffe000u;
  hi = c.f;
#endif
  lo = a - hi;
}

// ---- TMEM-resident A operand ("TS" form): A[m, k] lives on TMEM lane m, one fp32 column per k -----------------------
// 32 consecutive columns (v[0..31]) of the warp's 32 TMEM lanes (warp-collective)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]),
        "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]),
        "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]),
        "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::tf32: a_tmem = TMEM address of the first of the 8 k-columns of this k-step
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}

// ---- cluster-scope mbarrier signalling (a CTA arrives on a barrier in a PEER CTA's shared memory) -------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// named barrier over `count` threads (a multiple of 32) with an OR reduction of a predicate
__device__ __forceinline__ bool bar_red_or(int name, int count, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 q, %3, 0;\n\t"
      "barrier.cta.red.or.pred p, %1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(name), "r"(count), "r"(static_cast<uint32_t>(pred))
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void bar_named(int name, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(name), "r"(count) : "memory");
}

constexpr int kAtomK = 32;                          // fp32 elements of K per 128-byte swizzle atom
constexpr int kUmmaK = 8;                           // K per tcgen05.mma kind::tf32
constexpr int kWChunkBytes = 2 * 128 * 128;         // packed weight chunk: 128 features x 32 k, hi image then lo image

// The descriptors of one kernel differ only in the 14-bit address field of the low word, so the issuing thread
// keeps 32-bit low words and pairs them with the constant high word (the MMA issue loop of a single thread is on the
// critical path of the small-N GEMMs: 64-bit descriptor arithmetic per MMA made it issue-bound).
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_tf32_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}

// D^T[128, N] (+)= A_chunk . B_katom^T for one k-atom: 4 k-steps x 3 split passes = 12 MMAs.
// a_chunk: smem address of a packed weight chunk (hi image, lo image 16 KB later);
// b_hi / b_lo: smem addresses of this k-atom's block of the activation tile (hi, lo).
__device__ __forceinline__ void umma_katom_3x(uint32_t tmem_d, uint32_t a_chunk, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                              bool first) {
  const uint32_t ah = umma_desc_lo(a_chunk), al = ah + ((128u * 128u) >> 4);
  const uint32_t bh = umma_desc_lo(b_hi), bl = umma_desc_lo(b_lo);
#pragma unroll
  for (int ks = 0; ks < kAtomK / kUmmaK; ++ks) {
    const uint32_t adv = static_cast<uint32_t>(ks * kUmmaK * 4) >> 4;  // +32 B per k-step inside the atom
    umma_tf32_lo(tmem_d, al + adv, bh + adv, idesc, (first && ks == 0) ? 0u : 1u);   // small terms first
    umma_tf32_lo(tmem_d, ah + adv, bl + adv, idesc, 1u);
    umma_tf32_lo(tmem_d, ah + adv, bh + adv, idesc, 1u);
  }
}

}  // namespace tc
