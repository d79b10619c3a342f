// cuda_emul.h — TEST INFRASTRUCTURE ONLY.
//
// A minimal CPU shim that lets the SIMT kernel sources (paddlescience_b200/csrc/*.cuh, *.cu)
// be compiled with g++ (-DPPSCI_EMUL) and executed with one OS thread per CUDA thread, one
// block at a time.  It exists so that indexing / tiling / barrier logic can be checked against
// the oracle on a box without a GPU.  It is never linked into the product library and the
// package never loads it (tests/emul/build_emul.py builds tests/emul/_build/libppsci_b200_emul.so,
// loaded only by tests/).  The tcgen05 kernels are NOT emulated.
#pragma once
#ifndef PPSCI_EMUL
#error "cuda_emul.h is only for -DPPSCI_EMUL builds"
#endif

#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static  // one block runs at a time, so a function-static IS block-shared
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emul {
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char* g_smem = nullptr;
inline pthread_barrier_t g_barrier;
}  // namespace emul

#define threadIdx (emul::t_threadIdx)
#define blockIdx (emul::t_blockIdx)
#define blockDim (emul::g_blockDim)
#define gridDim (emul::g_gridDim)

inline void __syncthreads() { pthread_barrier_wait(&emul::g_barrier); }

// warp shuffles: lanes of a warp rendezvous on a per-warp barrier and exchange through scratch
namespace emul {
inline pthread_barrier_t g_warp_barrier[64];
inline double g_warp_scratch[64][32];
}  // namespace emul
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  const unsigned t = emul::t_threadIdx.x + emul::g_blockDim.x * (emul::t_threadIdx.y + emul::g_blockDim.y * emul::t_threadIdx.z);
  const unsigned w = t >> 5, l = t & 31;
  emul::g_warp_scratch[w][l] = (double)v;
  pthread_barrier_wait(&emul::g_warp_barrier[w]);
  const T r = (T)emul::g_warp_scratch[w][l ^ (unsigned)lane_mask];
  pthread_barrier_wait(&emul::g_warp_barrier[w]);
  return r;
}

inline float atomicAdd(float* addr, float v) {
  uint32_t* ia = reinterpret_cast<uint32_t*>(addr);
  uint32_t old = __atomic_load_n(ia, __ATOMIC_RELAXED);
  for (;;) {
    float f;
    memcpy(&f, &old, 4);
    float nf = f + v;
    uint32_t ni;
    memcpy(&ni, &nf, 4);
    if (__atomic_compare_exchange_n(ia, &old, ni, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
  }
}
inline double atomicAdd(double* addr, double v) {
  uint64_t* ia = reinterpret_cast<uint64_t*>(addr);
  uint64_t old = __atomic_load_n(ia, __ATOMIC_RELAXED);
  for (;;) {
    double f;
    memcpy(&f, &old, 8);
    double nf = f + v;
    uint64_t ni;
    memcpy(&ni, &nf, 8);
    if (__atomic_compare_exchange_n(ia, &old, ni, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
  }
}

// ---- runtime API stubs -------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; int major; int minor; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; p->major = 10; p->minor = 0; return cudaSuccess; }

typedef void* cudaEvent_t;
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* t, cudaEvent_t, cudaEvent_t) { *t = 0.f; return cudaSuccess; }

namespace emul {
// Run `body` once per (block, thread).  Blocks run sequentially; the threads of a block are
// real OS threads synchronised by a pthread barrier, so __syncthreads() semantics are honest.
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  g_blockDim = block;
  g_gridDim = grid;
  const unsigned nthr = block.x * block.y * block.z;
  std::vector<unsigned char> smem(smem_bytes + 64);
  g_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 63) & ~uintptr_t(63));
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        pthread_barrier_init(&g_barrier, nullptr, nthr);
        for (unsigned w = 0; w < (nthr + 31) / 32; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, std::min(32u, nthr - 32 * w));
        std::vector<std::thread> th;
        th.reserve(nthr);
        for (unsigned t = 0; t < nthr; ++t) {
          th.emplace_back([=, &body]() {
            t_blockIdx = dim3(bx, by, bz);
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            body();
          });
        }
        for (auto& x : th) x.join();
        pthread_barrier_destroy(&g_barrier);
        for (unsigned w = 0; w < (nthr + 31) / 32; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
      }
}
}  // namespace emul

#define PPSCI_DYN_SMEM(name) unsigned char* name = emul::g_smem
#define PPSCI_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emul::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })

struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
