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

#ifndef PPSCI_EMUL
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <string>

#include "kernels_simt.cuh"

namespace ppsci {
namespace tc {

constexpr int KCH = 32;  // K elements per chunk = one 128-byte swizzle row of tf32
constexpr int THREADS = 512;  // 16 warps (single-CTA kernels: 15 operand-producer warps + 1 MMA / weight-copy warp)
constexpr int NPROD = THREADS - 32;  // producer threads (warps 0..14)
constexpr int NPW = NPROD / 32;      // producer warps
constexpr int MMA_WARP = THREADS / 32 - 1;
constexpr int A_TILE_BYTES = 128 * KCH * 4;  // 16 KB (one of hi / lo)

// byte offset of element (row, kk) inside a [rows x 32 fp32] K-major SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t sw128(int row, int kk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ row) & 7) << 4) + ((kk & 3) << 2));
}

// ---- UMMA descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B,
// 8-row group pitch (SBO) 1024 B, LBO unused for swizzled K-major (=1), version 1 (Blackwell).
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
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
// ---- PTX wrappers (sm_100a).  The -DPPSCI_EMUL test build (tests/emul/) swaps in CPU emulations with the same names. ----
#ifdef PPSCI_EMUL
}  // namespace tc
}  // namespace ppsci
#include "tc_emul_prims.h"
namespace ppsci {
namespace tc {
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {  // non-suspending probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
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
// One lane polls the mbarrier, the rest of the warp waits at __syncwarp: 32x fewer pollers on the barrier word.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
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

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // PPSCI_EMUL

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

// ---- jet layouts ---------------------------------------------------------------------------------
// The operand producers are instruction-bound, so the channel structure is a compile-time parameter for
// the common PDE layouts (all index math folds into constants); DLay is the runtime fallback.
template <int O0, int O1, int O2, int O3>
struct SLay {
  static constexpr bool kStatic = true;
  static constexpr int ND = (O0 > 0) + (O1 > 0) + (O2 > 0) + (O3 > 0);
  static constexpr int C = 1 + O0 + O1 + O2 + O3;
  static constexpr int CS = C;  // compile-time channel count
  static constexpr int KMAX = (O0 > O1 ? O0 : O1) > (O2 > O3 ? O2 : O3) ? (O0 > O1 ? O0 : O1) : (O2 > O3 ? O2 : O3);
  static constexpr int KM = KMAX < 1 ? 1 : KMAX;
  __device__ static __forceinline__ int nd(const JetLayout&) { return ND; }
  __device__ static __forceinline__ int order(const JetLayout&, int d) { return d == 0 ? O0 : d == 1 ? O1 : d == 2 ? O2 : O3; }
  __device__ static __forceinline__ int cbase(const JetLayout&, int d) {
    return 1 + (d > 0 ? O0 : 0) + (d > 1 ? O1 : 0) + (d > 2 ? O2 : 0);
  }
  __device__ static __forceinline__ int nchan(const JetLayout&) { return C; }
  __device__ static __forceinline__ int tp(int) { return 128 / C; }
  __device__ static __forceinline__ int pt(int) { return KCH / C; }
};
template <int KMAX_>
struct DLay {
  static constexpr bool kStatic = false;
  static constexpr int ND = PPSCI_MAX_DIR;
  static constexpr int CS = 1;  // (channel count only known at run time)
  static constexpr int KM = KMAX_;
  __device__ static __forceinline__ int nd(const JetLayout& J) { return J.n_dir; }
  __device__ static __forceinline__ int order(const JetLayout& J, int d) { return J.dir_order[d]; }
  __device__ static __forceinline__ int cbase(const JetLayout& J, int d) { return J.dir_base[d]; }
  __device__ static __forceinline__ int nchan(const JetLayout& J) { return J.C; }
  __device__ static __forceinline__ int tp(int rt) { return rt; }
  __device__ static __forceinline__ int pt(int rt) { return rt; }
};
// ACT < 0: activation id taken from the arguments at run time
template <int ACT>
__device__ __forceinline__ int act_id(int rt) { return ACT >= 0 ? ACT : rt; }

// ---- block staging with cp.async (LDGSTS), used by the dx epilogue for the Z_{l-1} blocks -----------------
// (Measured: one cp.async.bulk per 128-byte row costs ~100 issue cycles each and serialises; 16-byte
// cp.async pieces spread over all threads are 2x faster here.)  The piece -> (row, 16-byte column) mapping
// does not change from block to block, so each thread precomputes its pieces once.  The main loops do NOT
// stage operand rows this way any more: they go global -> registers -> swizzled tiles (see k_tc_fwd).
#ifndef PPSCI_EMUL
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#endif

constexpr int ROW_PIECES = (128 * 8 + NPROD - 1) / NPROD;  // 16-byte pieces of a [128 x 32] tile per producer thread


// pieces of a [rows_used x 32] tile whose row r = c*TP + pl comes from plane c, point p0 + pl
struct RowPieces {
  long long src[ROW_PIECES];  // c*plane + pl*ld + 4q  (element offset; add p0*ld + col0)
  uint32_t dst[ROW_PIECES];   // r*128 + 16 q
  int pl[ROW_PIECES];         // -1: no piece
  __device__ __forceinline__ void init(int TP, int rows_used, long long plane, int ld) {
#pragma unroll
    for (int j = 0; j < ROW_PIECES; ++j) {
      const int i = threadIdx.x + j * NPROD;
      const int r = i >> 3, q = i & 7;
      const int c = r / TP, pl_ = r - c * TP;
      pl[j] = (threadIdx.x < NPROD && r < rows_used) ? pl_ : -1;
      src[j] = (long long)c * plane + (long long)pl_ * ld + q * 4;
      dst[j] = (uint32_t)(r * 128 + q * 16);
    }
  }
  __device__ __forceinline__ void issue(uint32_t dst_base, const float* Z, long long p0_ld_col0, int valid_pts) const {
#pragma unroll
    for (int j = 0; j < ROW_PIECES; ++j) {
      if (pl[j] >= 0) {
        const bool ok = pl[j] < valid_pts;
        cp_async16(dst_base + dst[j], ok ? Z + src[j] + p0_ld_col0 : Z, ok);
      }
    }
  }
};

__host__ __device__ inline int tc_stage_bytes(int N) { return 2 * A_TILE_BYTES + 2 * N * KCH * 4; }
__host__ __device__ inline uint32_t tc_pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

// 3xTF32 MMAs of one 32-wide K chunk.  Two accumulators: acc0 takes ONLY the exact-product "big" term A_hi B_hi
// (4 accumulations per chunk), acc1 the two small cross terms.  The accumulator rounds toward zero at its own
// magnitude (measured, tests/microbench/tc_numerics.cu), so the cross terms — 2^-11 of the result — must not be added
// into the big accumulator: acc0 sees 4 instead of 12 truncations per chunk (error of a 256-deep contraction
// 6.3e-7 instead of 1.9e-6; the round-1 alternating split gave 1.1e-6), acc1's truncations are 2^-11 smaller.
// The epilogue adds acc0 * (1 + rz compensation) + acc1 with round-to-nearest.
// The issuing lane is the kernel's critical serial path: descriptors are a precomputed base (stage 0,
// k-step 0) plus increments of the 14-bit start-address field (units of 16 bytes).
struct ChunkDescs {
  uint64_t a_hi, a_lo, b_hi, b_lo;  // stage 0, k-step 0
  uint64_t stage_inc;               // stage_bytes >> 4
  __device__ __forceinline__ void init(uint32_t base_addr, int stage_bytes, int N) {
    a_hi = make_smem_desc(base_addr);
    a_lo = make_smem_desc(base_addr + A_TILE_BYTES);
    b_hi = make_smem_desc(base_addr + 2 * A_TILE_BYTES);
    b_lo = make_smem_desc(base_addr + 2 * A_TILE_BYTES + (uint32_t)(N * KCH * 4));
    stage_inc = (uint64_t)(stage_bytes >> 4);
  }
};
__device__ __forceinline__ void issue_chunk_mmas(uint32_t acc0, uint32_t acc1, const ChunkDescs& D, uint32_t stage,
                                                 uint32_t idesc, bool first_chunk) {
  const uint64_t so = stage ? D.stage_inc : 0;
#pragma unroll
  for (int ks = 0; ks < KCH / 8; ++ks) {  // one MMA consumes K = 8 tf32 = 32 bytes (2 x 16 B) of every row
    const uint64_t inc = so + (uint64_t)(2 * ks);
    const uint64_t dah = D.a_hi + inc, dal = D.a_lo + inc, dbh = D.b_hi + inc, dbl = D.b_lo + inc;
    const bool first = first_chunk && ks == 0;
    mma_tf32(acc0, dah, dbh, idesc, first ? 0u : 1u);
    mma_tf32(acc1, dal, dbh, idesc, first ? 0u : 1u);
    mma_tf32(acc1, dah, dbl, idesc, 1u);
  }
}
// dW variant (N <= 128): B_hi and B_lo are adjacent 128-byte-row tiles, i.e. one K-major tile with 2N rows, so
//   accX[128 x 2N] += A_hi [B_hi ; B_lo]^T     (one MMA gives A_hi B_hi | A_hi B_lo)
//   accY[128 x  N] += A_lo  B_hi^T
// 8 MMAs per chunk instead of 12; the epilogue adds accX[:, :N] + accX[:, N:] + accY.
__device__ __forceinline__ void issue_chunk_mmas_cat(uint32_t accX, uint32_t accY, const ChunkDescs& D, uint32_t stage,
                                                     uint32_t idesc_2n, uint32_t idesc_n, bool first_chunk) {
  const uint64_t so = stage ? D.stage_inc : 0;
#pragma unroll
  for (int ks = 0; ks < KCH / 8; ++ks) {
    const uint64_t inc = so + (uint64_t)(2 * ks);
    const bool first = first_chunk && ks == 0;
    mma_tf32(accX, D.a_hi + inc, D.b_hi + inc, idesc_2n, first ? 0u : 1u);
    mma_tf32(accY, D.a_lo + inc, D.b_hi + inc, idesc_n, first ? 0u : 1u);
  }
}

// ---- compensation of the accumulator's round-toward-zero --------------------------------------------------
// Measured (tests/microbench/tc_numerics.cu, profiles/r02_tc_numerics.txt): tcgen05 adds the exactly summed K = 8
// products of one instruction to the fp32 accumulator with ROUND TOWARD ZERO.  Every accumulate therefore shrinks the
// partial sum by 0.72 * 2^-24 of its magnitude on average; for exchangeable terms E[partial_k | total] = (k / n)
// total, so n accumulations of significant magnitude shrink the result by a factor 1 - c0 * n / 2 — a pure, data
// independent bias (the three trials of the micro-benchmark, row magnitudes 1e-9 .. 1e3, agree to 3 digits):
//   c0 = 3.35e-8 per accumulation (model 0.5 * 0.7213 * 2^-23 = 4.3e-8, measured 78 % of it)
// The epilogues multiply the accumulator sum by (1 + c0 * sum_acc events_acc / 2 * share_acc); what remains is the
// zero-mean part of the rounding (fp32-chain level).
constexpr float TC_RZ_C0 = 3.35e-8f;
// issue_chunk_mmas(_2) and issue_chunk_mmas_cat: 4 accumulations of the exact-product term per 32-wide chunk at the
// full magnitude of the result (the cross terms live in their own accumulator / columns, 2^-11 of it)
__host__ __device__ __forceinline__ float tc_rz_comp_split(long long nchunks) { return 1.f + TC_RZ_C0 * 2.f * (float)nchunks; }
__host__ __device__ __forceinline__ float tc_rz_comp_cat(long long nchunks) { return 1.f + TC_RZ_C0 * 2.f * (float)nchunks; }

// (acc0 + acc1) * comp for 32 lanes x 32 columns
__device__ __forceinline__ void load_acc_sum(uint32_t acc0, uint32_t acc1, int q, int col, float (&out)[32], float comp) {
  uint32_t v0[32], v1[32];
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  tmem_ld32(acc0 + lane_off + (uint32_t)col, v0);
  tmem_ld32(acc1 + lane_off + (uint32_t)col, v1);
  tmem_ld_wait();
#pragma unroll
  for (int t = 0; t < 32; ++t) out[t] = fmaf(__uint_as_float(v0[t]), comp, __uint_as_float(v1[t]));
}

// store v = hi + lo into the (hi, lo) pair of K-major SW128 tiles; `off` = sw128(row, col)
__device__ __forceinline__ void store_split_at(unsigned char* a_hi, uint32_t off, float v) {
  const float hi = tf32_rn(v);
  *reinterpret_cast<float*>(a_hi + off) = hi;
  *reinterpret_cast<float*>(a_hi + A_TILE_BYTES + off) = v - hi;
}
// 128-bit variant: four consecutive k (one 16-byte chunk of the row); `off` = sw128_q(row, kq)
__device__ __forceinline__ void store_split4_at(unsigned char* a_hi, uint32_t off, const float (&v)[4]) {
  float4 h, l;
  h.x = tf32_rn(v[0]); h.y = tf32_rn(v[1]); h.z = tf32_rn(v[2]); h.w = tf32_rn(v[3]);
  l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
  *reinterpret_cast<float4*>(a_hi + off) = h;
  *reinterpret_cast<float4*>(a_hi + A_TILE_BYTES + off) = l;
}
__device__ __forceinline__ uint32_t sw128_q(int row, int kq) { return (uint32_t)(row * 128 + (((kq ^ row) & 7) << 4)); }
// column kk = lane: the swizzled in-row offset only depends on (row & 7)
__device__ __forceinline__ uint32_t sw128_lane(int row, int lane) {
  return (uint32_t)(row * 128 + ((((lane >> 2) ^ row) & 7) << 4) + ((lane & 3) << 2));
}

struct TcFwdArgs {
  AOperand<float> A;  // A_ACT over Z_{l-1}
  JetLayout J;
  const float* Wimg;  // [K/32][2][N*32] swizzled hi / lo images
  int Kdim;           // contraction length of the weight image (a multiple of the chunk width)
  int Kvalid;         // operand columns that exist (0 = Kdim): a dense first-layer operand [N][100] is read as [N][128]
  int Nout;
  const float* bias;
  float* Out;
  int ldo;
  long long oplane;
  float* Astash;      // optional: post-activation jets a_{l-1} [C][Np][lda] for the dW kernel
  int lda;
  long long aplane;
  long long Np;
  int TP;
  int num_tiles;
  long long* dbg;     // optional timeline buffer (bring-up instrumentation; null in production)
};

// Shared-memory map of k_tc_fwd / k_tc_dx (offsets from the 1024-aligned base):
//   [0, 2*stage)   two operand stages: A_hi | A_lo | B_hi | B_lo
//   then mbarriers b_full[2] (+0,+8), mma_done[2] (+16,+24), a_ready[2] (+32,+40) and the TMEM base slot (+64)
__host__ __device__ inline int tc_fwd_smem_bytes(int N) { return 2 * tc_stage_bytes(N) + 1024 + 256; }

// Shared prologue of the three kernels: barriers, TMEM, zeroed operand stages.
__device__ __forceinline__ void tc_setup(uint32_t base, unsigned char* base_ptr, uint32_t bars_off, int stage_bytes,
                                         int clear_bytes_per_stage, uint32_t ncols, int npw = NPW) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bars = base + bars_off;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);
    mbar_init(bars + 32, npw);  // a_ready[2]: one arrival per producer warp
    mbar_init(bars + 40, npw);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(base + bars_off + 64, ncols);
    tmem_relinquish();
  }
  for (int s = 0; s < 2; ++s) {
    float4* az = reinterpret_cast<float4*>(base_ptr + s * stage_bytes);
    for (int i = tid; i < clear_bytes_per_stage / 16; i += (int)blockDim.x) az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// Forward layer  Z_l = act_jets(Z_{l-1}) W_l + b_l  on the tensor cores.
// Persistent: CTA t handles tiles t, t+grid, ...; each tile = 128 rows = TP points x C channels.
// Pipeline per K chunk `it`: the operand rows of chunk it+1 are already in registers (prefetched while chunk `it`
// was produced), the producer warps write the A operand of chunk `it` into its stage and arrive on a_ready, the
// MMA warp's lane 0 waits for a_ready + the weight image (one TMA bulk copy per chunk) and issues the 12 MMAs.
// This single-CTA kernel is the fallback / cross-check of the CTA-pair kernel k_tc2_fwd (kernels_tc2.cuh).
template <class L, int ACT>
__global__ void __launch_bounds__(THREADS, 1) k_tc_fwd(TcFwdArgs g) {
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout;
  const int stage_bytes = tc_stage_bytes(N);
  const uint32_t bars_off = 2 * stage_bytes;
  const uint32_t bars = base + bars_off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(2 * N);
  tc_setup(base, base_ptr, bars_off, stage_bytes, 2 * A_TILE_BYTES, ncols);
  const uint32_t acc0 = *reinterpret_cast<volatile uint32_t*>(base_ptr + bars_off + 64), acc1 = acc0 + (uint32_t)N;
  const uint32_t idesc = make_idesc_tf32(128, N);
  ChunkDescs descs;
  descs.init(base, stage_bytes, N);
  const int nchunks = g.Kdim / KCH;
  const uint32_t b_bytes = (uint32_t)(2 * N * KCH * 4);
  const int TP = L::tp(g.TP);
  const int C = L::nchan(g.J);
  const int rows_used = C * TP;
  const int act = act_id<ACT>(g.A.act);
  const int my_tiles = ((int)blockIdx.x < g.num_tiles) ? (g.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t total_it = (uint32_t)my_tiles * (uint32_t)nchunks;

  auto issue_b = [&](uint32_t itb) {  // weight images of running chunk itb -> stage itb & 1
    const uint32_t sb = itb & 1u;
    mbar_expect_tx(bars + 8 * sb, b_bytes);
    bulk_g2s(base + sb * stage_bytes + 2 * A_TILE_BYTES, g.Wimg + (long long)(itb % nchunks) * 2 * N * KCH, b_bytes,
             bars + 8 * sb);
  };
  // Operand rows are fetched straight from HBM/L2 into REGISTERS one chunk ahead, and every LSU instruction of the
  // producers is 128 bits wide.  Measured on the timeline of these kernels: the LSU instruction rate (global loads,
  // shared stores), not bytes, bounds the producers -- cp.async staging cost ~900 cycles of issue per chunk plus a
  // second LDS pass, and 32-bit loads/stores cost 4x the instructions.  An item = (point, 4 consecutive k): per
  // channel one LDG.128, one STG.128 (a-stash) and two STS.128 (hi / lo; 4 consecutive k = one 16-byte chunk of a
  // K-major SW128 row).  Lanes: kq = lane & 7 (k quad of the 32-wide chunk), point = 4 warp + (lane >> 3): a
  // quarter warp writes the 8 chunks of one row (bank-conflict free) and reads 128 contiguous bytes.
  constexpr int CS = L::CS;
  constexpr int PPR = NPW * 4;  // points per pass of the producer warps
  constexpr int MAXI = L::kStatic ? (128 / CS + PPR - 1) / PPR : (128 + NPW - 1) / NPW;
  float4 zreg[L::kStatic ? MAXI : 1][CS];
  const bool is_mma = (warp == MMA_WARP);
  const int kq = lane & 7, psub = lane >> 3;
  auto prefetch = [&](int tile_r, int jr) {
    if constexpr (L::kStatic) {
      const long long p0r = (long long)tile_r * TP;
      const int col = jr * KCH + 4 * kq;
      const bool col_ok = col < (g.Kvalid ? g.Kvalid : g.Kdim);  // zero columns beyond a padded contraction length
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int pl = warp * 4 + psub + i * PPR;
        const long long p = p0r + pl;
        const bool ok = pl < TP && p < g.Np && col_ok;
        const float* src = g.A.Z + p * g.A.ld + col;
#pragma unroll
        for (int c = 0; c < CS; ++c)
          zreg[i][c] = ok ? __ldg(reinterpret_cast<const float4*>(src + (long long)c * g.A.plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  if (total_it > 0) {
    if (!is_mma) prefetch(blockIdx.x, 0);
    if (is_mma && lane == 0) issue_b(0);
  }
  const bool dbgp = g.dbg && blockIdx.x == 0 && tid == 0;
  const bool dbgm = g.dbg && blockIdx.x == 0 && is_mma && lane == 0;
#ifdef PPSCI_B200_TIMELINE
#define DBG_STAMP(cond, slot) do { if ((cond) && it < 48) g.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define DBG_STAMP(cond, slot) do { } while (0)
#endif
  uint32_t it = 0;  // running chunk counter (stage = it & 1, use index = it >> 1)
  for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
    const long long p0 = (long long)tile * TP;
    for (int j = 0; j < nchunks; ++j, ++it) {
      const uint32_t s = it & 1u, u = it >> 1;
      if (!is_mma) {
        DBG_STAMP(dbgp, 0);
        unsigned char* stage_ptr = base_ptr + s * stage_bytes;
        if constexpr (L::kStatic) {
          float4 zcur[MAXI][CS];
#pragma unroll
          for (int i = 0; i < MAXI; ++i)
#pragma unroll
            for (int c = 0; c < CS; ++c) zcur[i][c] = zreg[i][c];
          if (it + 1 < total_it) {  // next chunk's rows start flying now and land while this chunk is produced
            int ntile = tile, nj = j + 1;
            if (nj == nchunks) { nj = 0; ntile = tile + gridDim.x; }
            prefetch(ntile, nj);
          }
          DBG_STAMP(dbgp, 1);
          if (u >= 1) mbar_wait_warp(bars + 16 + 8 * s, (u - 1) & 1u);  // MMAs that read this stage have retired
          DBG_STAMP(dbgp, 3);
#pragma unroll
          for (int i = 0; i < MAXI; ++i) {
            const int pl = warp * 4 + psub + i * PPR;
            if (pl >= TP) continue;
            const long long p = p0 + pl;
            const bool valid = p < g.Np;
            float yout[CS][4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              auto comp = [&](const float4& v) { return t == 0 ? v.x : t == 1 ? v.y : t == 2 ? v.z : v.w; };
              float sc[6];
              float y0;
              act_coef<float, L::KM>(act, comp(zcur[i][0]), y0, sc);
              yout[0][t] = valid ? y0 : 0.f;
#pragma unroll
              for (int d = 0; d < L::ND; ++d) {
                const int K = L::order(g.J, d), cb = L::cbase(g.J, d);
                float zz[4], yy[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) zz[o] = (o < L::KM && o < K) ? comp(zcur[i][(cb + o) < CS ? (cb + o) : 0]) : 0.f;
                jet_fwd_dir<float, L::KM>(sc, zz, yy);
#pragma unroll
                for (int o = 0; o < L::KM; ++o)
                  if (o < K && cb + o < CS) yout[cb + o][t] = valid ? yy[o] : 0.f;
              }
            }
            float* ast = (g.Astash && valid) ? g.Astash + p * g.lda + j * KCH + 4 * kq : nullptr;
#pragma unroll
            for (int c = 0; c < CS; ++c) {
              store_split4_at(stage_ptr, sw128_q(c * TP + pl, kq), yout[c]);
              if (ast)
                *reinterpret_cast<float4*>(ast + (long long)c * g.aplane) = make_float4(yout[c][0], yout[c][1], yout[c][2], yout[c][3]);
            }
          }
        } else {
          // runtime-layout fallback: item = (point, lane column), scalar and unprefetched
          if (u >= 1) mbar_wait_warp(bars + 16 + 8 * s, (u - 1) & 1u);
          const int k = j * KCH + lane;
          for (int pl = warp; pl < TP; pl += NPW) {
            const long long p = p0 + pl;
            if (p < g.Np) {
              const float* zsrc = g.A.Z + p * g.A.ld + k;
              float sc[6];
              float y0;
              act_coef<float, L::KM>(act, zsrc[0], y0, sc);
              store_split_at(stage_ptr, sw128_lane(pl, lane), y0);
              float* ast = g.Astash ? g.Astash + p * g.lda + k : nullptr;
              if (ast) ast[0] = y0;
#pragma unroll
              for (int d = 0; d < L::ND; ++d) {
                if (d < L::nd(g.J)) {
                  const int K = L::order(g.J, d), cb = L::cbase(g.J, d);
                  float zz[4], yy[4];
#pragma unroll
                  for (int o = 0; o < 4; ++o) zz[o] = (o < L::KM && o < K) ? zsrc[(long long)(cb + o) * g.A.plane] : 0.f;
                  jet_fwd_dir<float, L::KM>(sc, zz, yy);
#pragma unroll
                  for (int o = 0; o < L::KM; ++o)
                    if (o < K) {
                      store_split_at(stage_ptr, sw128_lane((cb + o) * TP + pl, lane), yy[o]);
                      if (ast) ast[(long long)(cb + o) * g.aplane] = yy[o];
                    }
                }
              }
            } else {
              for (int c = 0; c < C; ++c) store_split_at(stage_ptr, sw128_lane(c * TP + pl, lane), 0.f);
            }
          }
        }
        DBG_STAMP(dbgp, 4);
        fence_proxy_async();  // generic-proxy smem accesses ordered before the async-proxy ones that follow
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 32 + 8 * s);  // this warp's share of the A operand is in place
        DBG_STAMP(dbgp, 5);
      }
      if (is_mma) {
        if (lane == 0) {
          DBG_STAMP(dbgm, 8);
          mbar_wait(bars + 32 + 8 * s, u & 1u);  // every producer warp has delivered its rows of this chunk
          DBG_STAMP(dbgm, 9);
          mbar_wait(bars + 8 * s, u & 1u);  // weight images of this chunk have landed
          DBG_STAMP(dbgm, 12);
          tc_fence_after();
          issue_chunk_mmas(acc0, acc1, descs, s, idesc, j == 0);
          mma_commit(bars + 16 + 8 * s);  // arrives when every MMA issued so far has completed
          DBG_STAMP(dbgm, 10);
          if (it + 1 < total_it) {  // weights of chunk it+1 -> other stage, once chunk it-1's MMAs have left it
            if (it >= 1) mbar_wait(bars + 16 + 8 * ((it + 1) & 1u), ((it - 1) >> 1) & 1u);
            issue_b(it + 1);
          }
        }
        __syncwarp();
      }
    }
    // ---- epilogue: TMEM -> shared exchange tiles -> (+bias) -> Z_l in HBM, 128-bit and coalesced ----
    // A thread owns one accumulator ROW after tcgen05.ld.  Four 128x32 blocks at a time are exchanged through the
    // (now idle) A regions of both stages -- X[b][row][chunk ^ (row & 7)], 16-byte chunks, conflict-free both
    // ways -- and written back as (row, k quad) items: 8 lanes cover 128 contiguous bytes of one output row.
    {
      const uint32_t last = it - 1;
      mbar_wait_warp(bars + 16 + 8 * (last & 1u), (last >> 1) & 1u);
      tc_fence_after();
      if (dbgp && last < 48) g.dbg[last * 16 + 13] = clock64();
      const int q = warp & 3, part = warp >> 2;
      const int ncb = N / 32;
      for (int cb0 = 0; cb0 < ncb; cb0 += 4) {
        const int cb = cb0 + part;
        unsigned char* Xb = base_ptr + (part >> 1) * stage_bytes + (part & 1) * A_TILE_BYTES;
        if (cb < ncb) {
          float v[32];
          load_acc_sum(acc0, acc1, q, cb * 32, v, tc_rz_comp_split(nchunks));
          const int row = q * 32 + lane;
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4)
            *reinterpret_cast<float4*>(Xb + sw128_q(row, t4)) = make_float4(v[4 * t4], v[4 * t4 + 1], v[4 * t4 + 2], v[4 * t4 + 3]);
        }
        __syncthreads();
        for (int r = warp * 4 + psub; r < rows_used; r += (THREADS / 32) * 4) {
          const int c = r / TP, pl = r - c * TP;
          const long long p = p0 + pl;
          if (p < g.Np) {
            float* out_row = g.Out + (long long)c * g.oplane + p * g.ldo + cb0 * 32 + 4 * kq;
#pragma unroll
            for (int b4 = 0; b4 < 4; ++b4) {
              if (cb0 + b4 < ncb) {
                const unsigned char* Xr = base_ptr + (b4 >> 1) * stage_bytes + (b4 & 1) * A_TILE_BYTES;
                float4 val = *reinterpret_cast<const float4*>(Xr + sw128_q(r, kq));
                if (c == 0 && g.bias) {  // (the flat parameter vector does not guarantee 16-byte aligned biases)
                  const float* bp = g.bias + (cb0 + b4) * 32 + 4 * kq;
                  val.x += __ldg(bp); val.y += __ldg(bp + 1); val.z += __ldg(bp + 2); val.w += __ldg(bp + 3);
                }
                *reinterpret_cast<float4*>(out_row + b4 * 32) = val;
              }
            }
          }
        }
        __syncthreads();
      }
      // (rows >= rows_used of the exchange tiles hold accumulators of all-zero operand rows, i.e. zeros: the
      //  A regions' pad rows stay valid zero operands)
      if (dbgp && last < 48) g.dbg[last * 16 + 14] = clock64();
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();  // accumulators drained / scratch released before the next tile starts
    }
  }
#undef DBG_STAMP
  __syncthreads();
  if (warp == 1) tmem_dealloc(acc0, ncols);
}

// =====================================================================================================
// Backward dx on the tensor cores:  Abar = Zbar_l W_l^T  (contraction over the layer's outputs), then the
// activation adjoint  Zbar_{l-1} = adj(Abar, Z_{l-1}).  The C channels of one point live in different TMEM
// lanes and the adjoint needs them together, so each 32-column block goes through a shared-memory
// exchange tile; the matching Z_{l-1} block is prefetched by cp.async one block ahead.  During the
// epilogue all MMAs have retired, so the exchange tile lives in stage 0's A_hi region and the two Z
// buffers in stage 1's A_hi / A_lo regions (same 128-byte row pitch; pad rows stay zero).
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
  long long* dbg;      // optional timeline buffer (bring-up instrumentation; null in production)
};

template <class L, int ACT>
__global__ void __launch_bounds__(THREADS, 1) k_tc_dx(TcDxArgs g) {
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout;
  const int stage_bytes = tc_stage_bytes(N);
  const uint32_t bars_off = 2 * stage_bytes;
  const uint32_t bars = base + bars_off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(2 * N);
  tc_setup(base, base_ptr, bars_off, stage_bytes, 2 * A_TILE_BYTES, ncols);
  const uint32_t acc0 = *reinterpret_cast<volatile uint32_t*>(base_ptr + bars_off + 64), acc1 = acc0 + (uint32_t)N;
  const uint32_t idesc = make_idesc_tf32(128, N);
  ChunkDescs descs;
  descs.init(base, stage_bytes, N);
  const int nchunks = g.Kdim / KCH;
  const uint32_t b_bytes = (uint32_t)(2 * N * KCH * 4);
  const int TP = L::tp(g.TP);
  const int C = L::nchan(g.J);
  const int rows_used = C * TP;
  const int act = act_id<ACT>(g.act);
  const uint32_t zbuf_addr = base + stage_bytes;   // two Z blocks: stage 1, A_hi and A_lo regions
  const float* zbuf_ptr = reinterpret_cast<const float*>(base_ptr + stage_bytes);
  const int my_tiles = ((int)blockIdx.x < g.num_tiles) ? (g.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t total_it = (uint32_t)my_tiles * (uint32_t)nchunks;

  auto issue_b = [&](uint32_t itb) {
    const uint32_t sb = itb & 1u;
    mbar_expect_tx(bars + 8 * sb, b_bytes);
    bulk_g2s(base + sb * stage_bytes + 2 * A_TILE_BYTES, g.Wimg + (long long)(itb % nchunks) * 2 * N * KCH, b_bytes,
             bars + 8 * sb);
  };
  auto valid_pts = [&](long long p0r) {
    const long long vp = g.Np - p0r;
    return vp >= TP ? TP : (vp > 0 ? (int)vp : 0);
  };
  RowPieces zpcs;
  zpcs.init(TP, rows_used, g.zplane, g.ldz);
  // Zbar rows of the next chunk are prefetched into registers; items = (row, 4 consecutive k), 128-bit loads and
  // shared stores (see k_tc_fwd for why)
  constexpr int RPP = NPW * 4;  // rows per pass of the producer warps
  constexpr int MAXR = (128 + RPP - 1) / RPP;
  float4 zreg[MAXR];
  const bool is_mma = (warp == MMA_WARP);
  const int kq = lane & 7, psub = lane >> 3;
  auto prefetch = [&](int tile_r, int jr) {
    const long long p0r = (long long)tile_r * TP;
    const int col = jr * KCH + 4 * kq;
#pragma unroll
    for (int i = 0; i < MAXR; ++i) {
      const int r = warp * 4 + psub + i * RPP;
      const int c = r / TP, pl = r - c * TP;
      const bool ok = r < rows_used && p0r + pl < g.Np;
      zreg[i] = ok ? __ldg(reinterpret_cast<const float4*>(g.A.Z + (long long)c * g.A.plane + (p0r + pl) * g.A.ld + col))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (total_it > 0) {
    if (!is_mma) prefetch(blockIdx.x, 0);
    if (is_mma && lane == 0) issue_b(0);
  }
  const bool dbgp = g.dbg && blockIdx.x == 0 && tid == 0;
  const bool dbgm = g.dbg && blockIdx.x == 0 && is_mma && lane == 0;
#ifdef PPSCI_B200_TIMELINE
#define DBG_STAMP(cond, slot) do { if ((cond) && it < 48) g.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define DBG_STAMP(cond, slot) do { } while (0)
#endif
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
    const long long p0 = (long long)tile * TP;
    const int vpts = valid_pts(p0);
    for (int j = 0; j < nchunks; ++j, ++it) {
      const uint32_t s = it & 1u, u = it >> 1;
      if (!is_mma) {
        DBG_STAMP(dbgp, 0);
        unsigned char* stage_ptr = base_ptr + s * stage_bytes;
        float4 zcur[MAXR];
#pragma unroll
        for (int i = 0; i < MAXR; ++i) zcur[i] = zreg[i];
        if (it + 1 < total_it) {
          int ntile = tile, nj = j + 1;
          if (nj == nchunks) { nj = 0; ntile = tile + gridDim.x; }
          prefetch(ntile, nj);
        }
        DBG_STAMP(dbgp, 1);
        if (u >= 1) mbar_wait_warp(bars + 16 + 8 * s, (u - 1) & 1u);
        DBG_STAMP(dbgp, 3);
#pragma unroll
        for (int i = 0; i < MAXR; ++i) {  // plain split
          const int r = warp * 4 + psub + i * RPP;
          if (r < rows_used) {
            const float v[4] = {zcur[i].x, zcur[i].y, zcur[i].z, zcur[i].w};
            store_split4_at(stage_ptr, sw128_q(r, kq), v);
          }
        }
        DBG_STAMP(dbgp, 4);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 32 + 8 * s);
        DBG_STAMP(dbgp, 5);
      }
      if (is_mma) {
        if (lane == 0) {
          DBG_STAMP(dbgm, 8);
          mbar_wait(bars + 32 + 8 * s, u & 1u);
          DBG_STAMP(dbgm, 9);
          mbar_wait(bars + 8 * s, u & 1u);
          DBG_STAMP(dbgm, 12);
          tc_fence_after();
          issue_chunk_mmas(acc0, acc1, descs, s, idesc, j == 0);
          mma_commit(bars + 16 + 8 * s);
          DBG_STAMP(dbgm, 10);
          // weights of chunk it+1 go to the other stage; at a tile boundary that stage's A region is used as
          // epilogue scratch, but its B region is not, so the copy may be in flight across the epilogue
          if (it + 1 < total_it) {
            if (it >= 1) mbar_wait(bars + 16 + 8 * ((it + 1) & 1u), ((it - 1) >> 1) & 1u);
            issue_b(it + 1);
          }
        }
        __syncwarp();
      }
    }
    // ---- epilogue: Abar (TMEM) -> exchange tile -> activation adjoint -> Zbar_{l-1} ----
    {
      const uint32_t last = it - 1;
      mbar_wait_warp(bars + 16 + 8 * (last & 1u), (last >> 1) & 1u);
      tc_fence_after();
      if (dbgp && last < 48) g.dbg[last * 16 + 13] = clock64();
      const int ncb = N / 32;
      unsigned char* Xb = base_ptr;  // exchange tile: stage 0, A_hi region; X[row][chunk ^ (row & 7)]
      zpcs.issue(zbuf_addr, g.Zprev, p0 * g.ldz, vpts);
      cp_async_commit();
      for (int cb = 0; cb < ncb; ++cb) {
        if (cb + 1 < ncb) zpcs.issue(zbuf_addr + ((cb + 1) & 1) * A_TILE_BYTES, g.Zprev, p0 * g.ldz + (cb + 1) * 32, vpts);
        cp_async_commit();
        if (warp < 4) {  // 128 lanes x 32 columns of Abar
          float v[32];
          load_acc_sum(acc0, acc1, warp, cb * 32, v, tc_rz_comp_split(nchunks));
          const int row = warp * 32 + lane;
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4)
            *reinterpret_cast<float4*>(Xb + sw128_q(row, t4)) = make_float4(v[4 * t4], v[4 * t4 + 1], v[4 * t4 + 2], v[4 * t4 + 3]);
        }
        cp_async_wait<1>();
        __syncthreads();
        const unsigned char* zb = reinterpret_cast<const unsigned char*>(zbuf_ptr) + (cb & 1) * A_TILE_BYTES;  // [row][32] linear
        if constexpr (L::kStatic) {
          constexpr int CS = L::CS;
          for (int pl = warp * 4 + psub; pl < vpts; pl += (THREADS / 32) * 4) {  // item = (point, k quad)
            float4 zc[CS], xc[CS];
#pragma unroll
            for (int c = 0; c < CS; ++c) {
              const int rr = c * TP + pl;
              zc[c] = *reinterpret_cast<const float4*>(zb + rr * 128 + kq * 16);
              xc[c] = *reinterpret_cast<const float4*>(Xb + sw128_q(rr, kq));
            }
            float ob[CS][4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              auto comp = [&](const float4& v) { return t == 0 ? v.x : t == 1 ? v.y : t == 2 ? v.z : v.w; };
              float sc[6];
              float y0;
              act_coef<float, L::KM + 1>(act, comp(zc[0]), y0, sc);
              float sb[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int d = 0; d < L::ND; ++d) {
                const int K = L::order(g.J, d), cbs = L::cbase(g.J, d);
                float zz[4], yb[4], zbv[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  const bool on = (o < L::KM && o < K && cbs + o < CS);
                  zz[o] = on ? comp(zc[on ? cbs + o : 0]) : 0.f;
                  yb[o] = on ? comp(xc[on ? cbs + o : 0]) : 0.f;
                  zbv[o] = 0.f;
                }
                jet_adj_dir<float, L::KM>(sc, zz, yb, zbv, sb);
#pragma unroll
                for (int o = 0; o < L::KM; ++o)
                  if (o < K && cbs + o < CS) ob[cbs + o][t] = zbv[o];
              }
              ob[0][t] = jet_adj_z0<float, L::KM>(sc, comp(xc[0]), sb);
            }
            float* out = g.Out + (p0 + pl) * g.ldo + cb * 32 + 4 * kq;
#pragma unroll
            for (int c = 0; c < CS; ++c)
              *reinterpret_cast<float4*>(out + (long long)c * g.oplane) = make_float4(ob[c][0], ob[c][1], ob[c][2], ob[c][3]);
          }
        } else {
          const int nn = lane;  // runtime-layout fallback: item = (point, lane column), scalar
          const uint32_t nsw = (uint32_t)((nn & 3) << 2);
          for (int pl = warp; pl < vpts; pl += THREADS / 32) {
            float* zb_out = g.Out + (p0 + pl) * g.ldo + cb * 32 + nn;
            float sc[6];
            float y0;
            act_coef<float, L::KM + 1>(act, *reinterpret_cast<const float*>(zb + pl * 128 + nn * 4), y0, sc);
            const float y0b = *reinterpret_cast<const float*>(Xb + sw128_q(pl, nn >> 2) + nsw);
            float sb[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < L::ND; ++d) {
              if (d < L::nd(g.J)) {
                const int K = L::order(g.J, d), cbs = L::cbase(g.J, d);
                float zz[4], yb[4], zbv[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  const bool on = (o < L::KM && o < K);
                  const int rr = (cbs + o) * TP + pl;
                  zz[o] = on ? *reinterpret_cast<const float*>(zb + rr * 128 + nn * 4) : 0.f;
                  yb[o] = on ? *reinterpret_cast<const float*>(Xb + sw128_q(rr, nn >> 2) + nsw) : 0.f;
                  zbv[o] = 0.f;
                }
                jet_adj_dir<float, L::KM>(sc, zz, yb, zbv, sb);
#pragma unroll
                for (int o = 0; o < L::KM; ++o)
                  if (o < K) zb_out[(long long)(cbs + o) * g.oplane] = zbv[o];
              }
            }
            zb_out[0] = jet_adj_z0<float, L::KM>(sc, y0b, sb);
          }
        }
        fence_proxy_async();
        __syncthreads();
      }
      if (dbgp && last < 48) g.dbg[last * 16 + 14] = clock64();
      tc_fence_before();
      __syncthreads();
    }
  }
#undef DBG_STAMP
  cp_async_wait<0>();
  __syncthreads();
  if (warp == 1) tmem_dealloc(acc0, ncols);
}

// =====================================================================================================
// Backward dW on the tensor cores:  dW_l[k][n] += sum_rows A_{l-1}[row][k] * Zbar_l[row][n].
// Reduction dimension = jet rows (points x channels).  CTA (kt, split, nb) owns dW rows [128 kt, +128),
// columns [NC nb, +NC) and a contiguous range of 32-row reduction chunks (PT points each); it accumulates in
// TMEM across its whole range and flushes once with red.global.add.  A_{l-1} (post-activation jets) was
// stashed by the forward kernel, so both operands are plain copies: 4x4 blocks are gathered with 128-bit loads,
// transposed in registers and split into K-major SW128 tiles (see the producer geometry note below).
// =====================================================================================================
struct TcDwArgs {
  const float* Aact;   // a_{l-1} [C][Np][lda]
  int lda;
  long long aplane;
  JetLayout J;
  const float* Zbar;   // [C][Np][ldzb]
  int ldzb;
  long long zbplane;
  int Kdim;            // fan-in (rows of dW), multiple of 128
  int Nout;            // columns per CTA (NC), multiple of 32, <= 128
  int n0_stride;
  float* dW;           // [Kdim][ldw]
  int ldw;             // full fan-out of the layer (row pitch of dW)
  long long Np;
  int PT;              // points per 32-row reduction chunk
  int chunks_per_split;
  long long* dbg;      // optional timeline buffer (bring-up instrumentation; null in production)
  float* db;           // pair kernel only: bias gradient db_l[n] += sum_p Zbar_l[0][p][n] (null: not fused)
};

__host__ __device__ inline int tc_dw_smem_bytes(int NC) {
  return 2 * tc_stage_bytes(NC) + 1024 + 256;
}

constexpr int DW_NPW = 16;                      // producer warps of k_tc_dw (one 32-row x 16-reduction-row task each)
constexpr int DW_THREADS = (DW_NPW + 1) * 32;   // + the MMA warp
constexpr int DW_MMA_WARP = DW_NPW;

// Producer geometry.  Both operands of dW are TRANSPOSES of row-major global data: tile row = dW row k (A') or dW
// column n (B'), tile column kk = reduction row rr = (channel, point) of the chunk.  A thread owns a 4 x 4 block:
// it loads rr = 4q..4q+3 as four LDG.128 (4 consecutive k / n each), transposes in registers (pure renaming) and
// stores four 16-byte chunks (one per k / n row) into each of the hi and lo tiles.  A warp = 8 row quads x 4 q:
// lane = rq_lo | q_lo << 1 | rq_hi << 3, so that a quarter warp's STS.128 hits 8 distinct 16-byte chunks
// ((q ^ row) & 7 with row & 7 = 4 rq_lo + i) and every LDG.128 instruction reads four full 128-byte lines.
// Tasks: A' = 4 row groups x 2 q groups, B' = N/32 row groups x 2 q groups  (16 tasks at N = 128).
template <class L>
__global__ void __launch_bounds__(DW_THREADS, 1) k_tc_dw(TcDwArgs g) {
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout;  // columns of this CTA
  const int stage_bytes = tc_stage_bytes(N);
  const uint32_t bars_off = 2 * stage_bytes;
  const uint32_t bars = base + bars_off;  // mma_done[2] at +16,+24, a_ready[2] at +32,+40
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(3 * N);
  const int PT = L::pt(g.PT);
  const int C = L::nchan(g.J);
  const int rows_used = C * PT;
  tc_setup(base, base_ptr, bars_off, stage_bytes, stage_bytes, ncols, DW_NPW);  // whole stages cleared (pad columns stay 0)
  // accX: 2N columns (A_hi B_hi | A_hi B_lo), accY: N columns (A_lo B_hi)
  const uint32_t accX = *reinterpret_cast<volatile uint32_t*>(base_ptr + bars_off + 64), accY = accX + (uint32_t)(2 * N);
  const uint32_t idesc_2n = make_idesc_tf32(128, 2 * N), idesc_n = make_idesc_tf32(128, N);
  ChunkDescs descs;
  descs.init(base, stage_bytes, N);
  const int k0 = blockIdx.x * 128;
  const int n0 = blockIdx.z * g.n0_stride;
  const long long total_chunks = (g.Np + PT - 1) / PT;
  const long long ch_begin = (long long)blockIdx.y * g.chunks_per_split;
  long long ch_end = ch_begin + g.chunks_per_split;
  if (ch_end > total_chunks) ch_end = total_chunks;

  auto valid_pts = [&](long long ch) {
    const long long vp = g.Np - ch * PT;
    return vp >= PT ? PT : (vp > 0 ? (int)vp : 0);
  };
  const bool is_mma = (warp == DW_MMA_WARP);
  const int n_tasks = 8 + N / 16;
  const bool has_task = !is_mma && warp < n_tasks;  // (N <= 128: at most one task per producer warp)
  const bool t_isA = warp < 8;
  uint32_t g_off[4];            // element offsets of the four source rows relative to the chunk's first point
  uint32_t g_plb = 0xFFFFFFFFu; // their point-in-chunk, one byte each (0xFF: no such reduction row)
  uint32_t d_off[4];            // byte offsets of the four destination chunks inside the hi tile
  int d_lo = 0;                 // hi -> lo tile distance
  {
    const int tl = t_isA ? warp : warp - 8;
    const int rowgroup = tl >> 1, qg = tl & 1;
    const int rq = ((lane >> 3) << 1) | (lane & 1), q = qg * 4 + ((lane >> 1) & 3);
    const int row0 = rowgroup * 32 + rq * 4;
    uint32_t plb = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int rr = 4 * q + e;
      uint32_t b8 = 0xFFu;
      g_off[e] = 0;
      if (has_task && rr < rows_used && (!t_isA || k0 + row0 < g.Kdim)) {  // fan-in rows beyond Kdim (100 -> 128): zeros
        const int c = rr / PT, pl = rr - c * PT;
        b8 = (uint32_t)pl;
        g_off[e] = t_isA ? (uint32_t)((long long)c * g.aplane + (long long)pl * g.lda + k0 + row0)
                         : (uint32_t)((long long)c * g.zbplane + (long long)pl * g.ldzb + n0 + row0);
      }
      plb |= b8 << (8 * e);
    }
    g_plb = plb;
#pragma unroll
    for (int i = 0; i < 4; ++i) d_off[i] = sw128_q(row0 + i, q) + (t_isA ? 0u : (uint32_t)(2 * A_TILE_BYTES));
    d_lo = t_isA ? A_TILE_BYTES : N * KCH * 4;
  }
  auto prefetch = [&](float4 (&buf)[4], long long ch) {
    const uint32_t vp = (uint32_t)valid_pts(ch);
    const float* bp = t_isA ? g.Aact + ch * PT * (long long)g.lda : g.Zbar + ch * PT * (long long)g.ldzb;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = ((g_plb >> (8 * e)) & 255u) < vp;
      buf[e] = ok ? __ldg(reinterpret_cast<const float4*>(bp + g_off[e])) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  const bool dbgp = g.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
  const bool dbgm = g.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && is_mma && lane == 0;
#ifdef PPSCI_B200_TIMELINE
#define DBG_STAMP(cond, slot) do { if ((cond) && it < 48) g.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define DBG_STAMP(cond, slot) do { } while (0)
#endif
  // one reduction chunk: producers split the register-resident block of chunk `ch` into the operand stage and
  // refill the same registers with chunk ch+2; the MMA warp issues the chunk's 8 MMAs
  auto step = [&](float4 (&buf)[4], long long ch, uint32_t it) {
    const uint32_t s = it & 1u, u = it >> 1;
    if (!is_mma) {
      DBG_STAMP(dbgp, 0);
      unsigned char* stage_ptr = base_ptr + s * stage_bytes;
      if (u >= 1) mbar_wait_warp(bars + 16 + 8 * s, (u - 1) & 1u);  // MMAs that read this stage have retired
      DBG_STAMP(dbgp, 3);
      if (has_task) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          auto comp = [&](const float4& v) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; };
          const float v[4] = {comp(buf[0]), comp(buf[1]), comp(buf[2]), comp(buf[3])};
          float4 h, l;
          h.x = tf32_rn(v[0]); h.y = tf32_rn(v[1]); h.z = tf32_rn(v[2]); h.w = tf32_rn(v[3]);
          l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
          *reinterpret_cast<float4*>(stage_ptr + d_off[i]) = h;
          *reinterpret_cast<float4*>(stage_ptr + d_off[i] + d_lo) = l;
        }
      }
      DBG_STAMP(dbgp, 4);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 32 + 8 * s);
      DBG_STAMP(dbgp, 5);
      if (has_task && ch + 2 < ch_end) prefetch(buf, ch + 2);
      DBG_STAMP(dbgp, 6);
    }
    if (is_mma) {
      if (lane == 0) {
        DBG_STAMP(dbgm, 8);
        mbar_wait(bars + 32 + 8 * s, u & 1u);
        DBG_STAMP(dbgm, 9);
        tc_fence_after();
        issue_chunk_mmas_cat(accX, accY, descs, s, idesc_2n, idesc_n, it == 0);
        mma_commit(bars + 16 + 8 * s);
        DBG_STAMP(dbgm, 10);
        if (dbgm) {  // true completion time of THIS chunk's MMAs (non-suspending poll; instrumentation only)
          while (!mbar_test_wait(bars + 16 + 8 * s, u & 1u)) {}
          DBG_STAMP(dbgm, 11);
        }
      }
      __syncwarp();
    }
  };
  float4 bufA[4], bufB[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) bufA[e] = bufB[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has_task) {
    if (ch_begin < ch_end) prefetch(bufA, ch_begin);
    if (ch_begin + 1 < ch_end) prefetch(bufB, ch_begin + 1);
  }
  uint32_t it = 0;
  for (long long ch = ch_begin; ch < ch_end; ch += 2, it += 2) {
    step(bufA, ch, it);
    if (ch + 1 < ch_end) step(bufB, ch + 1, it + 1);
  }
  it = ch_begin < ch_end ? (uint32_t)(ch_end - ch_begin) : 0u;
#undef DBG_STAMP
  if (it > 0 && !is_mma) {
    const uint32_t last = it - 1;
    mbar_wait_warp(bars + 16 + 8 * (last & 1u), (last >> 1) & 1u);
    tc_fence_after();
    const int q = warp & 3, part = warp >> 2;
    const int k = k0 + q * 32 + lane;
    float* dw_row = g.dW + (long long)k * g.ldw + n0;
    const int ncb = N / 32;
    const float rz_comp = tc_rz_comp_cat((long long)it);
    for (int cb = part; cb < ncb; cb += DW_NPW / 4) {
      float v[32], w[32];
      load_acc_sum(accX, accX + (uint32_t)N, q, cb * 32, v, 1.f);
      {
        uint32_t y[32];
        tmem_ld32(accY + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), y);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) w[t] = __uint_as_float(y[t]);
      }
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = (v[t] + w[t]) * rz_comp;
      if (k < g.Kdim) {
#pragma unroll
        for (int t = 0; t < 32; ++t) atomicAdd(dw_row + cb * 32 + t, v[t]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(accX, ncols);
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
  // hidden -> hidden layers, and a WIDE output layer (DeepONet sub-networks end in num_features = 128 units,
  // deeponet.py:96-119); the input layer has K = n_feat (2..3) and a PINN's output layer N = n_out (1..4): thin kernels
  if (s.dtype != PPSCI_F32) return false;
  if (l < 2 || l > s.n_layers) return false;
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
inline int tc_dw_cols_per_cta(int N) { return N <= 128 ? N : 128; }
inline bool tc_dw_ok(const ppsci_plan_spec& s, int l) {
  if (s.dtype != PPSCI_F32) return false;
  if (l < 2 || l > s.n_layers) return false;
  const int K = s.widths[l - 1], N = s.widths[l];
  return (K % 128 == 0) && K >= 128 && K <= 1024 && (N % 32 == 0) && N >= 32 && N <= 1024 &&
         (N % tc_dw_cols_per_cta(N) == 0);
}

// ---- kernel selection: static jet layouts for the common PDE structures, runtime layout otherwise ------
enum { TC_LAY_DYN = 0, TC_LAY_22 = 1, TC_LAY_12 = 2, TC_LAY_222 = 3, TC_LAY_VALUE = 4 };
inline int tc_pick_layout(const JetLayout& J, int act) {
  if (act != PPSCI_ACT_TANH) return TC_LAY_DYN;
  auto is = [&](int n, int a, int b, int c) {
    return J.n_dir == n && (n < 1 || J.dir_order[0] == a) && (n < 2 || J.dir_order[1] == b) && (n < 3 || J.dir_order[2] == c);
  };
  if (is(2, 2, 2, 0)) return TC_LAY_22;
  if (is(2, 1, 2, 0)) return TC_LAY_12;
  if (is(3, 2, 2, 2)) return TC_LAY_222;
  if (is(0, 0, 0, 0)) return TC_LAY_VALUE;
  return TC_LAY_DYN;
}

#define PPSCI_TC_LAUNCH_L(KERNEL, lay, kmax, grid, smem, stream, args, err_expr)  /* k_tc_dw: DW_THREADS */                               \
  do {                                                                                                            \
    void (*kfn_)(decltype(args)) = nullptr;                                                                       \
    switch (lay) {                                                                                                \
      case TC_LAY_22: kfn_ = tc::KERNEL<tc::SLay<2, 2, 0, 0>>; break;                                             \
      case TC_LAY_12: kfn_ = tc::KERNEL<tc::SLay<1, 2, 0, 0>>; break;                                             \
      case TC_LAY_222: kfn_ = tc::KERNEL<tc::SLay<2, 2, 2, 0>>; break;                                            \
      case TC_LAY_VALUE: kfn_ = tc::KERNEL<tc::SLay<0, 0, 0, 0>>; break;                                          \
      default: kfn_ = tc::KERNEL<tc::DLay<4>>;                                                                    \
    }                                                                                                             \
    cudaError_t e_ = cudaFuncSetAttribute(kfn_, cudaFuncAttributeMaxDynamicSharedMemorySize, (smem));             \
    if (e_ != cudaSuccess) { err_expr; }                                                                          \
    PPSCI_KLAUNCH(kfn_, (grid), dim3(tc::DW_THREADS), (smem), (stream), 1, args);                                  \
  } while (0)

#define PPSCI_TC_LAUNCH(KERNEL, lay, kmax, grid, smem, stream, args, err_expr)                                   \
  do {                                                                                                            \
    void (*kfn_)(decltype(args)) = nullptr;                                                                       \
    switch (lay) {                                                                                                \
      case TC_LAY_22: kfn_ = tc::KERNEL<tc::SLay<2, 2, 0, 0>, PPSCI_ACT_TANH>; break;                             \
      case TC_LAY_12: kfn_ = tc::KERNEL<tc::SLay<1, 2, 0, 0>, PPSCI_ACT_TANH>; break;                             \
      case TC_LAY_222: kfn_ = tc::KERNEL<tc::SLay<2, 2, 2, 0>, PPSCI_ACT_TANH>; break;                            \
      case TC_LAY_VALUE: kfn_ = tc::KERNEL<tc::SLay<0, 0, 0, 0>, PPSCI_ACT_TANH>; break;                          \
      default:                                                                                                    \
        kfn_ = (kmax) <= 1 ? tc::KERNEL<tc::DLay<1>, -1> : (kmax) == 2 ? tc::KERNEL<tc::DLay<2>, -1>              \
                                                                       : tc::KERNEL<tc::DLay<4>, -1>;             \
    }                                                                                                             \
    cudaError_t e_ = cudaFuncSetAttribute(kfn_, cudaFuncAttributeMaxDynamicSharedMemorySize, (smem));             \
    if (e_ != cudaSuccess) { err_expr; }                                                                          \
    PPSCI_KLAUNCH(kfn_, (grid), dim3(tc::THREADS), (smem), (stream), 1, args);                                     \
  } while (0)


inline bool tc_plan_supported(const ppsci_plan_spec& s, int /*C*/, int /*kmax*/) {
  for (int l = 2; l <= s.n_layers; ++l)
    if (tc_layer_ok(s, l)) return true;
  return false;
}
inline bool tc_dense_first_ok(const ppsci_plan_spec& s);

// scratch = weight images (hi+lo) of every eligible layer, forward orientation
inline size_t tc_img_offset(const ppsci_plan_spec& s, int layer) {
  size_t off = 0;
  for (int l = 2; l < layer; ++l)
    if (tc_layer_ok(s, l)) off += (size_t)s.widths[l - 1] * s.widths[l] * 8;
  return off;
}
inline size_t tc_imgT_offset(const ppsci_plan_spec& s, int layer) {  // transposed (dx) images follow the forward ones
  size_t off = tc_img_offset(s, s.n_layers + 1);
  for (int l = 2; l < layer; ++l)
    if (tc_dx_ok(s, l)) off += (size_t)s.widths[l - 1] * s.widths[l] * 8;
  return off;
}
// dense first layer (DeepONet branch net, K = num_loc): tensor-core eligible when its width is, the operand columns come in
// whole float4s and there are no input derivatives; its weight image (contraction length rounded up to the chunk width,
// zero rows beyond K) sits behind the other images
inline int tc_dense_kpad(const ppsci_plan_spec& s) { return (s.widths[0] + tc::KCH - 1) / tc::KCH * tc::KCH; }
inline bool tc_dense_first_ok(const ppsci_plan_spec& s) {
  if (s.dtype != PPSCI_F32 || !s.dense_in || s.n_dir != 0 || s.n_layers < 2) return false;
  const int K = s.widths[0], N = s.widths[1];
  return K % 4 == 0 && K >= 32 && tc_dense_kpad(s) <= 1024 && N % 32 == 0 && N >= 32 && N <= 256 && N % tc_dw_cols_per_cta(N) == 0;
}
inline size_t tc_dense_img_offset(const ppsci_plan_spec& s) { return tc_imgT_offset(s, s.n_layers + 1); }
inline size_t tc_scratch_bytes_impl(const ppsci_plan_spec& s, int /*C*/, int64_t /*nc*/) {
  return tc_imgT_offset(s, s.n_layers + 1) + (tc_dense_first_ok(s) ? (size_t)tc_dense_kpad(s) * s.widths[1] * 8 : 0) + 1024;
}

}  // namespace ppsci
