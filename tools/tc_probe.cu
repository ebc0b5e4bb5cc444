// Stand-alone self-checking probe of the tcgen05 building block used by the kernels (tc_common.cuh):
//   D^T[128, N] = A[128, 128] . B[N, 128]^T  with A delivered as a pre-packed SWIZZLE_128B image through
//   cp.async.bulk, B written by threads with st.shared, accumulators read back from TMEM.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I temp_b200/csrc -o tools/tc_probe tools/tc_probe.cu
// Run on a B200: prints the max relative error of the 1-pass (plain tf32) and 3-pass (split) products
// against a double-precision host reference.  Development tool, not part of the product path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_common.cuh"

template <int N>
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ a_packed, const float* __restrict__ b,
                                                       float* __restrict__ out, int passes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_img = smem;                               // 4 chunks x 32 KB
  uint8_t* b_hi = a_img + 4 * tc::kWChunkBytes;         // 4 k-atom blocks x N*128 B
  uint8_t* b_lo = b_hi + 4 * N * 128;
  __shared__ uint64_t bar_a, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(&bar_a, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, N < 32 ? 32 : N);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (tid == 0) {
    tc::mbar_expect_tx(&bar_a, 4 * tc::kWChunkBytes);
    for (int c = 0; c < 4; ++c) tc::bulk_g2s(a_img + c * tc::kWChunkBytes, a_packed + c * tc::kWChunkBytes, tc::kWChunkBytes, &bar_a);
  }
  for (int idx = tid; idx < N * 128; idx += 128) {
    const int n = idx >> 7, k = idx & 127;
    float hi, lo;
    tc::split_tf32(b[idx], hi, lo);
    const uint32_t off = (k >> 5) * (N * 128) + tc::sw128_off(n, k & 31);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  tc::fence_proxy_async();
  __syncthreads();

  if (tid == 0) {
    tc::mbar_wait(&bar_a, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::umma_idesc_tf32(128, N);
    for (int ka = 0; ka < 4; ++ka) {
      const uint32_t a_chunk = tc::smem_u32(a_img + ka * tc::kWChunkBytes);
      const uint32_t bh = tc::smem_u32(b_hi + ka * N * 128), bl = tc::smem_u32(b_lo + ka * N * 128);
      if (passes == 3) {
        tc::umma_katom_3x(tbase, a_chunk, bh, bl, idesc, ka == 0);
      } else {
        for (int ks = 0; ks < 4; ++ks)
          tc::umma_tf32(tbase, tc::umma_desc(a_chunk) + 2 * ks, tc::umma_desc(bh) + 2 * ks, idesc, (ka | ks) ? 1u : 0u);
      }
    }
    tc::umma_commit(&bar_mma);
  }
  __syncwarp();
  tc::mbar_wait(&bar_mma, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tbase + (static_cast<uint32_t>(32 * warp) << 16) + c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) out[(32 * warp + lane) * N + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, N < 32 ? 32 : N);
}

static void pack_host(const std::vector<float>& a /*[128][128] m,k*/, std::vector<uint8_t>& img) {
  img.assign(4 * tc::kWChunkBytes, 0);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 128; ++k) {
      float hi, lo;
      tc::split_tf32(a[m * 128 + k], hi, lo);
      const size_t off = static_cast<size_t>(k >> 5) * tc::kWChunkBytes + tc::sw128_off(m, k & 31);
      memcpy(&img[off], &hi, 4);
      memcpy(&img[off + 128 * 128], &lo, 4);
    }
}

template <int N>
static int run(int passes, double tol) {
  std::vector<float> a(128 * 128), b(N * 128);
  srand(1234 + N);
  for (auto& x : a) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : b) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  std::vector<uint8_t> img;
  pack_host(a, img);
  uint8_t* d_img;
  float *d_b, *d_out;
  cudaMalloc(&d_img, img.size());
  cudaMalloc(&d_b, b.size() * 4);
  cudaMalloc(&d_out, 128 * N * 4);
  cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0xff, 128 * N * 4);
  const size_t smem = 4 * tc::kWChunkBytes + 2 * 4 * N * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<N><<<1, 128, smem>>>(d_img, d_b, d_out, passes);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d passes=%d: CUDA error %s\n", N, passes, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> out(128 * N);
  cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  int bad_m = -1, bad_n = -1;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < 128; ++k) ref += (double)a[m * 128 + k] * (double)b[n * 128 + k];
      const double err = fabs(ref - out[m * N + n]);
      if (!(err <= maxerr)) { maxerr = err; bad_m = m; bad_n = n; }
      if (fabs(ref) > maxref) maxref = fabs(ref);
    }
  const double rel = maxerr / maxref;
  printf("N=%3d passes=%d: max abs err %.3e, max |ref| %.3e, rel %.3e (worst at m=%d n=%d: got %.6f)  %s\n", N, passes, maxerr,
         maxref, rel, bad_m, bad_n, out[bad_m * N + bad_n], rel < tol ? "OK" : "FAIL");
  if (!(rel < tol)) {
    for (int m = 0; m < 2; ++m)
      for (int n = 0; n < 4; ++n) {
        double ref = 0;
        for (int k = 0; k < 128; ++k) ref += (double)a[m * 128 + k] * (double)b[n * 128 + k];
        printf("   D[%d][%d] got %.6f want %.6f\n", m, n, out[m * N + n], ref);
      }
  }
  cudaFree(d_img); cudaFree(d_b); cudaFree(d_out);
  return rel < tol ? 0 : 1;
}

int main() {
  int rc = 0;
  rc |= run<64>(1, 5e-3);
  rc |= run<64>(3, 2e-6);
  rc |= run<32>(3, 2e-6);
  rc |= run<16>(3, 2e-6);
  printf(rc == 0 ? "tc_probe: ALL OK\n" : "tc_probe: FAILED\n");
  return rc;
}
