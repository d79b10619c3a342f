// kernels_tc.cuh — tcgen05 / TMEM / TMA kernels (sm_100a only) for the wide fp32 layers.
//
// fp32-faithful tensor-core contraction: tcgen05 has no fp32 MMA, kind::tf32 keeps 10 mantissa
// bits.  Every operand is split x = hi + lo with hi = rn_tf32(x), lo = x - hi (exact in fp32),
// and each K-chunk issues three MMAs   D += A_hi B_hi + A_lo B_hi + A_hi B_lo   (3xTF32; the
// dropped A_lo B_lo term is O(2^-22) relative).  Issued MMA flops = 3 x algorithmic flops.
//
// Layout of one operand tile in shared memory: K-major, 32 fp32 (=128 B) per row, rows at
// 128 B pitch, 8-row groups at 1024 B pitch, SWIZZLE_128B (16-byte chunk index XOR row%8) —
// the canonical UMMA "Layout_K_SW128" atom.  The A tile is produced by the CTA's threads
// (activation jets of the previous layer's pre-activations, applied on the fly); the B tile is a
// pre-swizzled image of the layer's weights prepared once per call by k_tc_prep_w and staged
// with one cp.async.bulk (TMA, no tensor map) per K-chunk onto an mbarrier.
// Accumulator: 128 lanes x N fp32 columns in TMEM; read back with tcgen05.ld.32x32b.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "kernels_simt.cuh"

namespace ppsci {
namespace tc {

constexpr int KCH = 32;  // K elements per chunk = one 128-byte swizzle row of tf32
constexpr int THREADS = 256;
constexpr int A_TILE_BYTES = 128 * KCH * 4;  // 16 KB (one of hi / lo)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, kk) inside a [rows x 32 fp32] K-major SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t sw128(int row, int kk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ row) & 7) << 4) + ((kk & 3) << 2));
}

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA bulk copy global -> shared (1-D, no tensor map) --------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B,
// 8-row group pitch (SBO) 1024 B, LBO unused for swizzled K-major (=1), version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset >> 4, bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset >> 4, bits [32,46)
  d |= (uint64_t)1 << 46;                    // version, bits [46,48)
  d |= (uint64_t)2 << 61;                    // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, K-major A and B.
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- weight image: Wimg[j][part][sw128(n, kk)] = split(W_as_B[n][32 j + kk]) ---------------------------
// transposed = 0: B[n][k] = W[k * N + n]   (forward: W is [K=in][N=out] row-major)
// transposed = 1: B[n][k] = W[n * K + k]   (dx: contraction over the layer's outputs)
__global__ void k_tc_prep_w(const float* __restrict__ W, float* __restrict__ img, int K, int N, int transposed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K * N) return;
  const int n = (int)(i / K), k = (int)(i % K);
  const float w = transposed ? W[(long long)n * K + k] : W[(long long)k * N + n];
  const float hi = tf32_rn(w);
  const float lo = w - hi;
  const int j = k / KCH, kk = k % KCH;
  float* blk = img + (long long)j * 2 * N * KCH;
  const uint32_t off = sw128(n, kk) >> 2;
  blk[off] = hi;
  blk[(long long)N * KCH + off] = lo;
}

struct TcFwdArgs {
  AOperand<float> A;
  JetLayout J;
  const float* Wimg;  // [K/32][2][N*32] swizzled hi / lo images
  int Kdim;
  int Nout;
  const float* bias;
  float* Out;
  int ldo;
  long long oplane;
  long long Np;
  int TP;
  int num_tiles;
};

__host__ __device__ inline int tc_stage_bytes(int N) { return 2 * A_TILE_BYTES + 2 * N * KCH * 4; }
__host__ __device__ inline uint32_t tc_tmem_cols(int N) {
  uint32_t c = 32;
  while ((int)c < N) c <<= 1;
  return c;
}

// Forward layer  Z_l = act_jets(Z_{l-1}) W_l + b_l  on the tensor cores.
// Persistent: CTA t handles tiles t, t+grid, ...; each tile = 128 rows = TP points x C channels.
template <int KMAX>
__global__ void __launch_bounds__(THREADS, 1) k_tc_fwd(TcFwdArgs g) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout;
  const int stage_bytes = tc_stage_bytes(N);
  const uint32_t bars = base + 2 * stage_bytes;  // full[2], mma_done[2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 2 * stage_bytes + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_tmem_cols(N);

  if (tid == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 8, 1);
    mbar_init(bars + 16, 1);
    mbar_init(bars + 24, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(base + 2 * stage_bytes + 64, ncols);
    tmem_relinquish();
  }
  // rows that never receive data (>= C*TP) must read as zero: clear both A tiles once
  for (int s = 0; s < 2; ++s) {
    float4* az = reinterpret_cast<float4*>(base_ptr + s * stage_bytes);
    for (int i = tid; i < 2 * A_TILE_BYTES / 16; i += THREADS) az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_tf32(128, N);
  const int nchunks = g.Kdim / KCH;
  const uint32_t b_bytes = (uint32_t)(2 * N * KCH * 4);
  const int TP = g.TP;
  const int rows_used = g.J.C * TP;

  uint32_t it = 0;  // running chunk counter (stage = it & 1, use index = it >> 1)
  for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
    const long long p0 = (long long)tile * TP;
    for (int j = 0; j < nchunks; ++j, ++it) {
      const uint32_t s = it & 1u, u = it >> 1;
      unsigned char* stage_ptr = base_ptr + s * stage_bytes;
      const uint32_t stage_addr = base + s * stage_bytes;
      if (u >= 1) mbar_wait(bars + 16 + 8 * s, (u - 1) & 1u);  // MMAs that read this stage have retired
      if (tid == 0) {
        mbar_expect_tx(bars + 8 * s, b_bytes);
        bulk_g2s(stage_addr + 2 * A_TILE_BYTES, g.Wimg + (long long)j * 2 * N * KCH, b_bytes, bars + 8 * s);
      }
      // produce the A chunk (hi, lo) for k in [32 j, 32 j + 32)
      for (int item = tid; item < TP * KCH; item += THREADS) {
        const int kk = item & (KCH - 1), pl = item / KCH;
        const long long p = p0 + pl;
        const int k = j * KCH + kk;
        produce_a<float, KMAX>(g.A, g.J, p, k, p < g.Np, [&](int c, float v) {
          const int r = c * TP + pl;
          const float hi = tf32_rn(v);
          const uint32_t off = sw128(r, kk);
          *reinterpret_cast<float*>(stage_ptr + off) = hi;
          *reinterpret_cast<float*>(stage_ptr + A_TILE_BYTES + off) = v - hi;
        });
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncthreads();
      if (tid == 0) {
        mbar_wait(bars + 8 * s, u & 1u);  // weight images of this chunk have landed
        tc_fence_after();
        const uint32_t a_hi = stage_addr, a_lo = stage_addr + A_TILE_BYTES;
        const uint32_t b_hi = stage_addr + 2 * A_TILE_BYTES, b_lo = b_hi + (uint32_t)(N * KCH * 4);
#pragma unroll
        for (int ks = 0; ks < KCH / 8; ++ks) {  // one MMA consumes K = 8 tf32 = 32 bytes of every row
          const uint64_t dah = make_smem_desc(a_hi + ks * 32), dal = make_smem_desc(a_lo + ks * 32);
          const uint64_t dbh = make_smem_desc(b_hi + ks * 32), dbl = make_smem_desc(b_lo + ks * 32);
          mma_tf32(tmem_base, dah, dbh, idesc, (j > 0 || ks > 0) ? 1u : 0u);
          mma_tf32(tmem_base, dal, dbh, idesc, 1u);
          mma_tf32(tmem_base, dah, dbl, idesc, 1u);
        }
        mma_commit(bars + 16 + 8 * s);  // arrives when every MMA issued so far has completed
      }
    }
    // ---- epilogue: TMEM -> registers -> (+bias) -> Z_l in HBM ----
    {
      const uint32_t last = it - 1;
      mbar_wait(bars + 16 + 8 * (last & 1u), (last >> 1) & 1u);
      tc_fence_after();
      const int q = warp & 3, half = warp >> 2;
      const int r = q * 32 + lane;
      const bool row_ok = r < rows_used;
      const int c = row_ok ? r / TP : 0, pl = row_ok ? r % TP : 0;
      const long long p = p0 + pl;
      const bool st_ok = row_ok && p < g.Np;
      float* out_row = g.Out + (long long)c * g.oplane + p * g.ldo;
      const int ncb = N / 32;  // 32-column blocks; warps with half=0 take even blocks, half=1 odd blocks
      for (int cb = half; cb < ncb; cb += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
        tmem_ld_wait();
        if (st_ok) {
#pragma unroll
          for (int t = 0; t < 32; t += 4) {
            const int n = cb * 32 + t;
            float4 o;
            o.x = __uint_as_float(v[t]);
            o.y = __uint_as_float(v[t + 1]);
            o.z = __uint_as_float(v[t + 2]);
            o.w = __uint_as_float(v[t + 3]);
            if (c == 0 && g.bias) {
              o.x += g.bias[n];
              o.y += g.bias[n + 1];
              o.z += g.bias[n + 2];
              o.w += g.bias[n + 3];
            }
            *reinterpret_cast<float4*>(out_row + n) = o;
          }
        }
      }
      if (N % 32) {  // N is a multiple of 16: one trailing 16-column block, handled by half 0 via a x32 read is not safe
        // (not reached: eligibility requires N % 32 == 0)
      }
      tc_fence_before();
      __syncthreads();  // accumulator drained before the next tile's first MMA overwrites it
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}


// =====================================================================================================
// Backward dx on the tensor cores:  Abar = Zbar_l W_l^T  (contraction over the layer's outputs), then the
// activation adjoint  Zbar_{l-1} = adj(Abar, Z_{l-1})  through a shared-memory exchange tile (the C
// channels of one point live in different TMEM lanes, the adjoint needs them together).
// =====================================================================================================
struct TcDxArgs {
  AOperand<float> A;   // A_PLAIN over Zbar_l
  JetLayout J;
  const float* Wimg;   // transposed image: gemm N = fan-in of the layer, gemm K = fan-out
  int Kdim;            // gemm K  (= N_l)
  int Nout;            // gemm N  (= K_l = width of layer l-1)
  const float* Zprev;  // Z_{l-1} [C][Np][ldz]
  int ldz;
  long long zplane;
  int act;
  float* Out;          // Zbar_{l-1} [C][Np][ldo]
  int ldo;
  long long oplane;
  long long Np;
  int TP;
  int num_tiles;
};

constexpr int XLD = 33;                          // exchange-tile row pitch (floats): conflict-free both ways
constexpr int X_TILE_BYTES = 128 * XLD * 4;      // one 128 x 32 block

template <int KMAX>
__global__ void __launch_bounds__(THREADS, 1) k_tc_dx(TcDxArgs g) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout;
  const int stage_bytes = tc_stage_bytes(N);
  float* xch = reinterpret_cast<float*>(base_ptr + 2 * stage_bytes);  // 2 exchange tiles (one per warp half)
  const uint32_t bars_off = 2 * stage_bytes + 2 * X_TILE_BYTES;
  const uint32_t bars = base + bars_off;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + bars_off + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_tmem_cols(N);

  if (tid == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 8, 1);
    mbar_init(bars + 16, 1);
    mbar_init(bars + 24, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(base + bars_off + 64, ncols);
    tmem_relinquish();
  }
  for (int s = 0; s < 2; ++s) {
    float4* az = reinterpret_cast<float4*>(base_ptr + s * stage_bytes);
    for (int i = tid; i < 2 * A_TILE_BYTES / 16; i += THREADS) az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_tf32(128, N);
  const int nchunks = g.Kdim / KCH;
  const uint32_t b_bytes = (uint32_t)(2 * N * KCH * 4);
  const int TP = g.TP;

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
    const long long p0 = (long long)tile * TP;
    for (int j = 0; j < nchunks; ++j, ++it) {
      const uint32_t s = it & 1u, u = it >> 1;
      unsigned char* stage_ptr = base_ptr + s * stage_bytes;
      const uint32_t stage_addr = base + s * stage_bytes;
      if (u >= 1) mbar_wait(bars + 16 + 8 * s, (u - 1) & 1u);
      if (tid == 0) {
        mbar_expect_tx(bars + 8 * s, b_bytes);
        bulk_g2s(stage_addr + 2 * A_TILE_BYTES, g.Wimg + (long long)j * 2 * N * KCH, b_bytes, bars + 8 * s);
      }
      for (int item = tid; item < TP * KCH; item += THREADS) {
        const int kk = item & (KCH - 1), pl = item / KCH;
        const long long p = p0 + pl;
        const int k = j * KCH + kk;
        produce_a<float, KMAX>(g.A, g.J, p, k, p < g.Np, [&](int c, float v) {
          const int r = c * TP + pl;
          const float hi = tf32_rn(v);
          const uint32_t off = sw128(r, kk);
          *reinterpret_cast<float*>(stage_ptr + off) = hi;
          *reinterpret_cast<float*>(stage_ptr + A_TILE_BYTES + off) = v - hi;
        });
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(bars + 8 * s, u & 1u);
        tc_fence_after();
        const uint32_t a_hi = stage_addr, a_lo = stage_addr + A_TILE_BYTES;
        const uint32_t b_hi = stage_addr + 2 * A_TILE_BYTES, b_lo = b_hi + (uint32_t)(N * KCH * 4);
#pragma unroll
        for (int ks = 0; ks < KCH / 8; ++ks) {
          const uint64_t dah = make_smem_desc(a_hi + ks * 32), dal = make_smem_desc(a_lo + ks * 32);
          const uint64_t dbh = make_smem_desc(b_hi + ks * 32), dbl = make_smem_desc(b_lo + ks * 32);
          mma_tf32(tmem_base, dah, dbh, idesc, (j > 0 || ks > 0) ? 1u : 0u);
          mma_tf32(tmem_base, dal, dbh, idesc, 1u);
          mma_tf32(tmem_base, dah, dbl, idesc, 1u);
        }
        mma_commit(bars + 16 + 8 * s);
      }
    }
    // ---- epilogue: Abar (TMEM) -> exchange tile -> activation adjoint -> Zbar_{l-1} ----
    {
      const uint32_t last = it - 1;
      mbar_wait(bars + 16 + 8 * (last & 1u), (last >> 1) & 1u);
      tc_fence_after();
      const int q = warp & 3, half = warp >> 2;
      float* X = xch + half * (128 * XLD);
      const int ncb = N / 32;
      const int tih = tid & 127;  // thread index inside its half
      for (int cb0 = 0; cb0 < ncb; cb0 += 2) {
        const int cb = cb0 + half;
        if (cb < ncb) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
          tmem_ld_wait();
          float* xr = X + (q * 32 + lane) * XLD;
#pragma unroll
          for (int t = 0; t < 32; ++t) xr[t] = __uint_as_float(v[t]);
        }
        __syncthreads();
        if (cb < ncb) {
          for (int item = tih; item < TP * 32; item += 128) {
            const int nn = item & 31, pl = item >> 5;
            const long long p = p0 + pl;
            if (p >= g.Np) continue;
            const int n = cb * 32 + nn;
            const float* z = g.Zprev + p * g.ldz + n;
            float* zb_out = g.Out + p * g.ldo + n;
            float sc[6];
            float y0;
            act_coef<float, KMAX + 1>(g.act, z[0], y0, sc);
            const float y0b = X[pl * XLD + nn];
            float sb[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            for (int d = 0; d < g.J.n_dir; ++d) {
              const int K = g.J.dir_order[d];
              const int cbase = g.J.dir_base[d];
              float zz[4], yb[4], zb[4];
#pragma unroll
              for (int o = 0; o < 4; ++o) {
                const bool on = (o < KMAX && o < K);
                zz[o] = on ? z[(long long)(cbase + o) * g.zplane] : 0.f;
                yb[o] = on ? X[((cbase + o) * TP + pl) * XLD + nn] : 0.f;
                zb[o] = 0.f;
              }
              jet_adj_dir<float, KMAX>(sc, zz, yb, zb, sb);
#pragma unroll
              for (int o = 0; o < KMAX; ++o)
                if (o < K) zb_out[(long long)(cbase + o) * g.oplane] = zb[o];
            }
            zb_out[0] = jet_adj_z0<float, KMAX>(sc, y0b, sb);
          }
        }
        __syncthreads();
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// =====================================================================================================
// Backward dW on the tensor cores:  dW_l[k][n] += sum_rows A_{l-1}[row][k] * Zbar_l[row][n].
// Reduction dimension = jet rows (points x channels).  CTA (kt, split) owns dW rows [128 kt, 128 kt + 128)
// and a contiguous range of 32-row reduction chunks (PT points each); it accumulates in TMEM across its
// whole range and flushes once with red.global.add.  Both operands are written transposed into K-major
// SW128 tiles: warps 0-3 recompute A_{l-1} = act_jets(Z_{l-1}) (thread = dW row k), warps 4-7 split Zbar_l.
// =====================================================================================================
struct TcDwArgs {
  AOperand<float> A;   // A_ACT over Z_{l-1}
  JetLayout J;
  const float* Zbar;   // [C][Np][ldzb]
  int ldzb;
  long long zbplane;
  int Kdim;            // fan-in (rows of dW), multiple of 128
  int Nout;            // fan-out (cols of dW), multiple of 32, <= 256
  float* dW;           // [Kdim][Nout]
  long long Np;
  int PT;              // points per 32-row reduction chunk
  int chunks_per_split;
};

template <int KMAX>
__global__ void __launch_bounds__(THREADS, 1) k_tc_dw(TcDwArgs g) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ int row_c[KCH], row_pl[KCH];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout;
  const int stage_bytes = tc_stage_bytes(N);
  const uint32_t bars_off = 2 * stage_bytes;
  const uint32_t bars = base + bars_off;  // mma_done[2] at +16, +24 (no TMA here: +0, +8 unused)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + bars_off + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_tmem_cols(N);
  const int PT = g.PT;
  const int rows_used = g.J.C * PT;

  if (tid == 0) {
    mbar_init(bars + 16, 1);
    mbar_init(bars + 24, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (tid < KCH) {
    row_c[tid] = tid < rows_used ? tid / PT : -1;
    row_pl[tid] = tid < rows_used ? tid % PT : 0;
  }
  if (warp == 1) {
    tmem_alloc(base + bars_off + 64, ncols);
    tmem_relinquish();
  }
  // clear both stages completely (reduction columns >= rows_used stay zero forever)
  for (int s = 0; s < 2; ++s) {
    float4* az = reinterpret_cast<float4*>(base_ptr + s * stage_bytes);
    for (int i = tid; i < stage_bytes / 16; i += THREADS) az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_tf32(128, N);
  const int k0 = blockIdx.x * 128;
  const long long total_chunks = (g.Np + PT - 1) / PT;
  const long long ch_begin = (long long)blockIdx.y * g.chunks_per_split;
  long long ch_end = ch_begin + g.chunks_per_split;
  if (ch_end > total_chunks) ch_end = total_chunks;

  uint32_t it = 0;
  for (long long ch = ch_begin; ch < ch_end; ++ch, ++it) {
    const uint32_t s = it & 1u, u = it >> 1;
    unsigned char* stage_ptr = base_ptr + s * stage_bytes;
    const uint32_t stage_addr = base + s * stage_bytes;
    if (u >= 1) mbar_wait(bars + 16 + 8 * s, (u - 1) & 1u);
    const long long pbase = ch * PT;
    if (warp < 4) {
      // A' tile: row = dW row k (this thread), column rr = c*PT + pl
      const int k = tid;  // 0..127
      for (int pl = 0; pl < PT; ++pl) {
        const long long p = pbase + pl;
        produce_a<float, KMAX>(g.A, g.J, p, k0 + k, p < g.Np, [&](int c, float v) {
          const int rr = c * PT + pl;
          const float hi = tf32_rn(v);
          const uint32_t off = sw128(k, rr);
          *reinterpret_cast<float*>(stage_ptr + off) = hi;
          *reinterpret_cast<float*>(stage_ptr + A_TILE_BYTES + off) = v - hi;
        });
      }
    } else {
      // B' tile: row = n, 16-byte chunk q holds reduction columns 4q..4q+3
      unsigned char* b_hi = stage_ptr + 2 * A_TILE_BYTES;
      unsigned char* b_lo = b_hi + N * KCH * 4;
      const int nq = (rows_used + 3) / 4;
      for (int item = tid - 128; item < N * nq; item += 128) {
        const int n = item % N, q = item / N;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int rr = 4 * q + e;
          const int c = row_c[rr];
          const long long p = pbase + row_pl[rr];
          const float v = (c >= 0 && p < g.Np) ? g.Zbar[(long long)c * g.zbplane + p * g.ldzb + n] : 0.f;
          hi[e] = tf32_rn(v);
          lo[e] = v - hi[e];
        }
        const uint32_t off = sw128(n, 4 * q);
        *reinterpret_cast<float4*>(b_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(b_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi = stage_addr, a_lo = stage_addr + A_TILE_BYTES;
      const uint32_t b_hi = stage_addr + 2 * A_TILE_BYTES, b_lo = b_hi + (uint32_t)(N * KCH * 4);
#pragma unroll
      for (int ks = 0; ks < KCH / 8; ++ks) {
        const uint64_t dah = make_smem_desc(a_hi + ks * 32), dal = make_smem_desc(a_lo + ks * 32);
        const uint64_t dbh = make_smem_desc(b_hi + ks * 32), dbl = make_smem_desc(b_lo + ks * 32);
        mma_tf32(tmem_base, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        mma_tf32(tmem_base, dal, dbh, idesc, 1u);
        mma_tf32(tmem_base, dah, dbl, idesc, 1u);
      }
      mma_commit(bars + 16 + 8 * s);
    }
  }
  if (it > 0) {
    const uint32_t last = it - 1;
    mbar_wait(bars + 16 + 8 * (last & 1u), (last >> 1) & 1u);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int k = k0 + q * 32 + lane;
    float* dw_row = g.dW + (long long)k * N;
    const int ncb = N / 32;
    for (int cb = half; cb < ncb; cb += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
      tmem_ld_wait();
      if (k < g.Kdim) {
#pragma unroll
        for (int t = 0; t < 32; ++t) atomicAdd(dw_row + cb * 32 + t, __uint_as_float(v[t]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// db_l[n] += sum over points of Zbar_l[channel 0][p][n]   (tiny, HBM-bound)
__global__ void k_bias_grad(const float* __restrict__ Zbar0, int ld, long long Np, int N, float* __restrict__ db,
                            int pts_per_block) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long p_begin = (long long)blockIdx.y * pts_per_block;
  long long p_end = p_begin + pts_per_block;
  if (p_end > Np) p_end = Np;
  float acc = 0.f;
  for (long long p = p_begin; p < p_end; ++p) acc += Zbar0[p * ld + n];
  atomicAdd(db + n, acc);
}

}  // namespace tc

// ---- host side ------------------------------------------------------------------------------------
inline bool tc_layer_ok(const ppsci_plan_spec& s, int l) {
  // hidden -> hidden layers only: the input layer has K = n_feat (2..3), the last layer N = n_out
  if (s.dtype != PPSCI_F32) return false;
  if (l < 2 || l >= s.n_layers) return false;
  const int K = s.widths[l - 1], N = s.widths[l];
  return (K % tc::KCH == 0) && K >= 32 && K <= 1024 && (N % 32 == 0) && N >= 32 && N <= 256;
}

// dx through layer l (produces Zbar_{l-1}): gemm K = N_l, gemm N = K_l; the layer below must be hidden
inline bool tc_dx_ok(const ppsci_plan_spec& s, int l) {
  if (s.dtype != PPSCI_F32) return false;
  if (l < 2 || l > s.n_layers) return false;
  const int K = s.widths[l], N = s.widths[l - 1];
  return (K % tc::KCH == 0) && K >= 32 && K <= 1024 && (N % 32 == 0) && N >= 32 && N <= 256;
}
// dW of layer l: rows = fan-in (multiple of 128), cols = fan-out (multiple of 32, <= 256)
inline bool tc_dw_ok(const ppsci_plan_spec& s, int l) {
  if (s.dtype != PPSCI_F32) return false;
  if (l < 2 || l > s.n_layers) return false;
  const int K = s.widths[l - 1], N = s.widths[l];
  return (K % 128 == 0) && K >= 128 && K <= 1024 && (N % 32 == 0) && N >= 32 && N <= 256;
}

inline bool tc_plan_supported(const ppsci_plan_spec& s, int /*C*/, int /*kmax*/) {
  for (int l = 2; l < s.n_layers; ++l)
    if (tc_layer_ok(s, l)) return true;
  return false;
}

// scratch = weight images (hi+lo) of every eligible layer, forward orientation
inline size_t tc_img_offset(const ppsci_plan_spec& s, int layer) {
  size_t off = 0;
  for (int l = 2; l < layer; ++l)
    if (tc_layer_ok(s, l)) off += (size_t)s.widths[l - 1] * s.widths[l] * 8;
  return off;
}
inline size_t tc_imgT_offset(const ppsci_plan_spec& s, int layer) {  // transposed (dx) images follow the forward ones
  size_t off = tc_img_offset(s, s.n_layers);
  for (int l = 2; l < layer; ++l)
    if (tc_dx_ok(s, l)) off += (size_t)s.widths[l - 1] * s.widths[l] * 8;
  return off;
}
inline size_t tc_scratch_bytes_impl(const ppsci_plan_spec& s, int /*C*/, int64_t /*nc*/) {
  return tc_imgT_offset(s, s.n_layers + 1) + 1024;
}

}  // namespace ppsci
