// Development probe: cycles of a batch of 48 tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of N, of the A operand's
// home (tensor memory "TS" / shared memory "SS") and of the number of accumulators the batch alternates between.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I temp_b200/csrc -I include -o tools/probes/mma_probe tools/probes/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace tc;

__global__ void __launch_bounds__(128, 1) probe(int N, int n_acc, int ts, int n_mma, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 1024) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f + (i % 7) * 0.125f;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tb, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tb;
  {  // A operand in TMEM: columns [0, 256)
    float v[32];
    for (int i = 0; i < 32; ++i) v[i] = 0.5f;
    for (int c = 0; c < 8; ++c) tmem_st32(tbase + (uint32_t(32 * warp) << 16) + 32 * c, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
    const uint32_t tbu = __shfl_sync(0xffffffffu, tbase, 0);
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint32_t a_s = umma_desc_lo(smem_u32(smem)), b_s = umma_desc_lo(smem_u32(smem) + 64 * 1024);
    const uint32_t dstride = n_acc > 1 ? ((N + 31) & ~31) : 0;
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    if (pred) {
      for (int rep = 0; rep < 3; ++rep) {
        const unsigned long long t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < n_mma; i += 4) {
          const uint32_t acc = i >= 4 ? 1u : 0u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t d = tbu + 256 + (n_acc == 4 ? q : (n_acc == 2 ? (q & 1) : 0)) * dstride;
            if (ts) umma_tf32_ts(d, tbu + ((i + q) & 15) * 8, b_s + 2 * q, idesc, acc);
            else umma_tf32_lo(d, a_s + 2 * q, b_s + 2 * q, idesc, acc);
          }
        }
        const unsigned long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, rep & 1);
        const unsigned long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%-4s %-5s %-5s %-6s %10s %10s %10s\n", "N", "n_acc", "A", "n_mma", "issue_cyc", "total_cyc", "cyc/mma");
  const int Ns[] = {16, 32, 48, 64, 96, 128, 256};
  for (int ts = 1; ts >= 0; --ts)
    for (int N : Ns)
      for (int n_acc : {1, 2, 4}) {
        if (n_acc * ((N + 31) & ~31) > 256) continue;
        for (int n_mma : {48, 96}) {
          probe<<<1, 128, 200 * 1024>>>(N, n_acc, ts, n_mma, d);
          unsigned long long h[2];
          if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          printf("%-4d %-5d %-5s %-6d %10llu %10llu %10.1f\n", N, n_acc, ts ? "tmem" : "smem", n_mma, h[0], h[1], double(h[1]) / n_mma);
        }
      }
  return 0;
}
