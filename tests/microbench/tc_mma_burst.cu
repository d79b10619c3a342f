// Micro-benchmark 2: bursts of tcgen05.mma separated by idle gaps (the pattern of a producer-limited kernel).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/ppsci_b200.h"
#include "../../paddlescience_b200/csrc/kernels_tc.cuh"
using namespace ppsci::tc;

__global__ void __launch_bounds__(256, 1) k_burst(int N, int burst, int gap, int nbursts, int wait_each, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* bp = smem_dyn + (base - smem_u32(smem_dyn));
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bars = base + 160 * 1024;
  if (tid == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); fence_barrier_init(); fence_proxy_async(); }
  if (warp == 1) { tmem_alloc(base + 160 * 1024 + 64, 512); tmem_relinquish(); }
  for (int i = tid; i < 160 * 1024 / 16; i += 256) reinterpret_cast<float4*>(bp)[i] = make_float4(1.f, 0.5f, 0.25f, 2.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *reinterpret_cast<volatile uint32_t*>(bp + 160 * 1024 + 64);
  if (tid == 0) {
    const uint64_t da = make_smem_desc(base), db = make_smem_desc(base + 32 * 1024);
    const uint32_t idesc = make_idesc_tf32(128, N);
    long long sum_issue = 0, sum_done = 0, max_done = 0;
    const long long tstart = clock64();
    for (int b = 0; b < nbursts; ++b) {
      const long long t0 = clock64();
      for (int i = 0; i < burst; ++i) mma_tf32(acc, da + (uint64_t)(2 * (i & 3)), db + (uint64_t)(2 * (i & 3)), idesc, (b | i) > 0);
      const long long t1 = clock64();
      mma_commit(bars + 8 * (b & 1));
      if (wait_each) {
        mbar_wait(bars + 8 * (b & 1), (b >> 1) & 1);
      } else if (b >= 1) {  // like the kernel: wait for the PREVIOUS burst only
        mbar_wait(bars + 8 * ((b - 1) & 1), ((b - 1) >> 1) & 1);
      }
      const long long t2 = clock64();
      sum_issue += t1 - t0;
      sum_done += t2 - t0;
      if (t2 - t0 > max_done) max_done = t2 - t0;
      while (clock64() - t2 < gap) {}
    }
    if (!wait_each) mbar_wait(bars + 8 * ((nbursts - 1) & 1), ((nbursts - 1) >> 1) & 1);
    if (blockIdx.x == 0) { out[0] = sum_issue / nbursts; out[1] = sum_done / nbursts; out[2] = max_done; out[3] = (clock64() - tstart) / nbursts; }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(acc, 512);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 64);
  const int smem = 160 * 1024 + 1024 + 256;
  cudaFuncSetAttribute(k_burst, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148})
    for (int wait_each : {1, 0})
      for (int N : {128, 256})
        for (int burst : {8, 12})
          for (int gap : {0, 500, 2000}) {
            k_burst<<<grid, 256, smem>>>(N, burst, gap, 200, wait_each, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[4]; cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost);
            printf("grid=%3d wait=%s N=%3d burst=%2d gap=%4d: issue %5lld  issue+wait %5lld (max %6lld)  period %5lld clk  (%s)\n", grid,
                   wait_each ? "this" : "prev", N, burst, gap, h[0], h[1], h[2], h[3], cudaGetErrorString(e));
          }
  return 0;
}
