// Micro-benchmark: completion time of back-to-back tcgen05.mma kind::tf32 / kind::f16 from shared-memory
// operands (SS), M = 128, issued by one thread, as a function of N, of the accumulator pattern and of
// concurrent shared-memory traffic from the other warps.  Build: nvcc -arch=sm_100a -o tc_mma_rate tc_mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define PPSCI_EMUL_SKIP
#include "../../include/ppsci_b200.h"
#include "../../paddlescience_b200/csrc/kernels_tc.cuh"
using namespace ppsci::tc;

__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 tf32 same acc, 1 tf32 two accs alternating, 2 bf16 same acc ; traffic: other warps hammer smem
__global__ void __launch_bounds__(256, 1) k_rate(int N, int nmma, int mode, int traffic, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* bp = smem_dyn + (base - smem_u32(smem_dyn));
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bars = base + 160 * 1024;
  if (tid == 0) { mbar_init(bars, 1); fence_barrier_init(); fence_proxy_async(); }
  if (warp == 1) { tmem_alloc(base + 160 * 1024 + 64, 512); tmem_relinquish(); }
  for (int i = tid; i < 160 * 1024 / 16; i += 256) reinterpret_cast<float4*>(bp)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *reinterpret_cast<volatile uint32_t*>(bp + 160 * 1024 + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(bp + 160 * 1024 + 128);
  if (tid == 0) *stop = 0;
  __syncthreads();
  if (tid == 0) {
    const uint64_t da = make_smem_desc(base), db = make_smem_desc(base + 32 * 1024);
    const uint32_t idesc = (mode == 2) ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                                       : make_idesc_tf32(128, N);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t d = (mode == 1 && (i & 1)) ? acc + 256 : acc;
      const uint64_t inc = (uint64_t)(2 * (i & 3));
      if (mode == 2) mma_f16(d, da + inc, db + inc, idesc, i > 1);
      else mma_tf32(d, da + inc, db + inc, idesc, i > 1);
    }
    const long long t1 = clock64();
    mma_commit(bars);
    mbar_wait(bars, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    *stop = 1;
  } else if (traffic && warp >= 2) {
    // generic-proxy shared-memory traffic on a disjoint region while the MMAs run
    // traffic 1: light dependent LDS/STS; traffic 2: back-to-back independent STS.128 (producer-like);
    // traffic 3: back-to-back independent LDS.128
    float4* scratch = reinterpret_cast<float4*>(bp + 96 * 1024);  // 64 KB region, disjoint from the operands
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f), acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int it = 0;
    while (!*stop && it < 2000000) {
      if (traffic == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) scratch[(tid + j * 256 + it * 64) & 4095] = v;
      } else if (traffic == 3) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float4 t = scratch[(tid + j * 256 + it * 64) & 4095]; acc4.x += t.x; acc4.y += t.y; }
      } else {
        const int idx = (tid + it * 64) & 4095;
        acc4.x += scratch[idx].x;
        scratch[(idx + 2048) & 4095] = acc4;
      }
      ++it;
    }
    if (acc4.x == 123.f) out[7] = it;
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(acc, 512);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 64);
  const int smem = 160 * 1024 + 1024 + 256;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* mname[3] = {"tf32 same-acc", "tf32 two-acc ", "bf16 same-acc"};
  for (int grid : {1, 148})
    for (int traffic : {0, 2})
    for (int mode = 0; mode < 3; mode += 2)
      for (int N : {128, 256})
        for (int nm : {64, 4000}) {
          k_rate<<<grid, 256, smem>>>(N, nm, mode, traffic, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
          printf("grid=%3d %s traffic=%d N=%3d nmma=%4d: issue %6lld clk, complete %6lld clk -> %.1f clk/MMA  (%s)\n", grid, mname[mode], traffic, N, nm,
                 h[0], h[1], (double)h[1] / nm, cudaGetErrorString(e));
        }
  return 0;
}
