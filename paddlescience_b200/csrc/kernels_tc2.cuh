// kernels_tc2.cuh — CTA-pair (tcgen05 cta_group::2) versions of the tensor-core kernels.
//
// Why pairs.  Timelines of the single-CTA kernels (kernels_tc.cuh) showed the main loops waiting for the weight
// images: every CTA re-streams all of W_hi / W_lo (64 KB per 32-wide K chunk at N = 256) from L2 for every
// 128-row tile, and with 148 CTAs doing so the L2 -> SM fabric (~6.3 KB/clk chip-wide, ~43 B/clk per SM) is the
// bound, not the tensor cores.  A CTA pair executes one M = 256 MMA over two point tiles (one per SM) while each
// SM stages only HALF of the B operand (N/2 weight rows); the hardware exchanges the halves.  Weight traffic per
// SM halves, the B stage shrinks to 32 KB so three operand stages fit, and one thread issues the MMAs of two SMs.
//
// Roles per CTA (512 threads): warps 0..13 produce the A operand (and run the epilogue), warp 14 lane 0 streams
// this CTA's half of the weight images (free-running, up to 3 chunks ahead), warp 15 lane 0 issues the pair's
// MMAs (leader CTA, cluster rank 0) or relays "my A tile and my B half are in place" to the leader (peer CTA).
//   full[s]       ONE barrier per stage collects everything the MMAs of a chunk need: one arrival per producer
//                 warp, the weight streamer's arrive.expect_tx + the bytes of its two bulk copies and, on the
//                 leader, one remote arrival from the peer's relay lane (which waits for the peer's own full[s]).
//                 The issuing lane is the critical serial path of the main loop; each extra barrier it polls
//                 costs ~200 cycles even when already complete (measured), hence a single one.
//   mma_done[s]   tcgen05.commit, multicast to both CTAs: stage s may be overwritten / accumulators are final
#pragma once
#include "kernels_tc.cuh"

namespace ppsci {
namespace tc {

constexpr int T2_NPW = 14;        // producer / epilogue warps
constexpr int T2_TMA_WARP = 14;
constexpr int T2_MMA_WARP = 15;
constexpr int T2_NSTAGE = 3;
constexpr int T2_PROD_THREADS = T2_NPW * 32;

__device__ __forceinline__ float f4comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
#ifndef PPSCI_EMUL
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
inline void prefetch_l2(const void*) {}
#endif
#ifndef PPSCI_EMUL
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.
// Default (.release.cta) semantics on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU and the matching
// `.acquire.cluster` wait to CCTL.IVALL (an L1 invalidate per chunk) — the ncu source page + the two-CTA timeline of
// round 2 showed the MMA lane starting 1.1 - 1.6 k cycles after the last operand arrival because of that membar.
// What the MMAs read is shared memory the workers already published to the async proxy (fence.proxy.async before
// their own arrival); the relay only observes that barrier and forwards the signal.
__device__ __forceinline__ void mbar_remote_arrive(uint32_t local_bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp32 <-> fp16 bit patterns (round to nearest even)
__device__ __forceinline__ uint32_t f32_to_f16_bits(float x) {
  uint16_t h;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return (uint32_t)h;
}
__device__ __forceinline__ float f16_bits_to_f32(uint32_t h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"((uint16_t)h));
  return f;
}
// two floats -> packed half2 bits (low half = a, high half = b), one instruction; and back
__device__ __forceinline__ uint32_t f32x2_to_f16x2_bits(float a, float b) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ void f16x2_bits_to_f32x2(uint32_t h, float& a, float& b) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(a), "=f"(b) : "r"(h));
}
// completion of all MMAs issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_2(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

#endif  // PPSCI_EMUL (tests/emul/tc_emul_prims.h provides the same names)

__host__ __device__ inline int tc2_stage_bytes(int N) { return 2 * A_TILE_BYTES + 2 * (N / 2) * KCH * 4; }
// barriers: mma_done[3] +24, full[3] +48, TMEM base slot +128
__host__ __device__ inline int tc2_smem_bytes(int N) { return T2_NSTAGE * tc2_stage_bytes(N) + 1024 + 256; }

// 12 MMAs (3xTF32, exact-product | cross-term accumulators: see issue_chunk_mmas) of one K chunk, M = 256 over the pair
__device__ __forceinline__ void issue_chunk_mmas_2(uint32_t acc0, uint32_t acc1, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                                   uint64_t b_lo, uint32_t idesc, bool first_chunk) {
#pragma unroll
  for (int ks = 0; ks < KCH / 8; ++ks) {
    const uint64_t inc = (uint64_t)(2 * ks);
    const uint64_t dah = a_hi + inc, dal = a_lo + inc, dbh = b_hi + inc, dbl = b_lo + inc;
    const bool first = first_chunk && ks == 0;
    mma_tf32_2(acc0, dah, dbh, idesc, first ? 0u : 1u);  // exact-product term alone in acc0
    mma_tf32_2(acc1, dal, dbh, idesc, first ? 0u : 1u);  // cross terms (2^-11 of the result) in acc1
    mma_tf32_2(acc1, dah, dbl, idesc, 1u);
  }
}

// barrier among the producer / epilogue warps only
__device__ __forceinline__ void t2_prod_sync() { named_bar_sync(2, T2_PROD_THREADS); }

// mbarrier parity waits are only meaningful while the target is the barrier's current or immediately preceding phase
// (an older target aliases: the wait falls through early, or blocks on a FUTURE phase and deadlocks).  Rule used by
// the producer warps: a warp only waits on mma_done for (a) the chunk three before a chunk it produces itself and
// (b) in the epilogue, a chunk its own group produced.  In both cases the previous phase of that barrier is known to
// be complete (the warp's previous own chunk waited past it) and the next phase cannot complete before the wait
// (it needs this warp's own production).  Warps that do not own the chunk are released through the producers'
// named barrier instead of polling a phase they may be more than one phase away from (a timing-dependent deadlock
// observed with two producer groups when one group lagged a full chunk cycle).
__device__ __forceinline__ void t2_wait_chunk_done(uint32_t bars, uint32_t c) {
  mbar_wait_warp(bars + 24 + 8 * (c % T2_NSTAGE), (c / T2_NSTAGE) & 1u);
}

struct T2Stage {  // running (stage, parity) pair of a 3-deep ring
  uint32_t s = 0, par = 0;
  __device__ __forceinline__ void next() {
    if (++s == T2_NSTAGE) {
      s = 0;
      par ^= 1u;
    }
  }
};

// Shared prologue: barriers, pair-wide TMEM allocation, zeroed A regions, cluster rendezvous.
__device__ __forceinline__ uint32_t tc2_setup(uint32_t base, unsigned char* base_ptr, uint32_t bars_off, int stage_bytes,
                                              uint32_t ncols, int full_count) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bars = base + bars_off;
  if (tid == 0) {
    for (int i = 0; i < T2_NSTAGE; ++i) {
      mbar_init(bars + 24 + 8 * i, 1);           // mma_done
      mbar_init(bars + 48 + 8 * i, full_count);  // full
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc2(bars + 128, ncols);
    tmem_relinquish2();
  }
  for (int s = 0; s < T2_NSTAGE; ++s) {
    float4* az = reinterpret_cast<float4*>(base_ptr + s * stage_bytes);
    for (int i = tid; i < 2 * A_TILE_BYTES / 16; i += (int)blockDim.x) az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(base_ptr + bars_off + 128);
}

// =====================================================================================================
// Forward layer on a CTA pair:  Z_l = act_jets(Z_{l-1}) W_l + b_l.  CTA rank r of pair q handles point tiles
// 2 t + r for t = q, q + pairs, ...  (a tile index beyond the last tile simply has no valid points).
// Static jet layouts only (the runtime-layout fallback stays on k_tc_fwd).
// =====================================================================================================
template <class L, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_tc2_fwd(TcFwdArgs g) {
  static_assert(L::kStatic, "pair kernels are specialised for the static jet layouts");
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout, NH = N / 2;
  const int stage_bytes = tc2_stage_bytes(N);
  const uint32_t bars_off = T2_NSTAGE * stage_bytes;
  const uint32_t bars = base + bars_off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(2 * N);
  constexpr int CS = L::CS;
  constexpr int TP = 128 / CS;
  constexpr int rows_used = CS * TP;
  // Producer groups.  One chunk's items (TP points x 8 k quads) need W1 warps; when two such sets fit in the 14
  // producer warps, group g owns chunks it = g (mod NG): every warp then has NG chunk periods per chunk, and its
  // two register buffers hold its next two own chunks (2 NG chunk periods of prefetch distance; a loaded HBM
  // round trip was measured at ~3,500 cycles, about two chunk periods).
  constexpr int W1 = (TP + 3) / 4;
  constexpr int NG = (2 * W1 <= T2_NPW) ? 2 : 1;
  constexpr int WG = NG > 1 ? W1 : T2_NPW;  // warps per group
  const uint32_t rank = cluster_ctarank();
  // full[s]: producer warps of the chunk + weight streamer (+ the peer's relay on the leader)
  const uint32_t acc0 = tc2_setup(base, base_ptr, bars_off, stage_bytes, ncols, WG + 1 + (rank == 0 ? 1 : 0)), acc1 = acc0 + (uint32_t)N;
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int nchunks = g.Kdim / KCH;
  const int act = act_id<ACT>(g.A.act);
  const int n_tile_pairs = (g.num_tiles + 1) / 2;
  const int my_tp = pair < n_tile_pairs ? (n_tile_pairs - 1 - pair) / npairs + 1 : 0;
  const uint32_t total_it = (uint32_t)my_tp * (uint32_t)nchunks;
  const bool dbg0 = g.dbg && blockIdx.x == 0;
#ifdef PPSCI_B200_TIMELINE
#define DBG_STAMP(cond, slot) do { if (dbg0 && (cond) && it < 48) g.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define DBG_STAMP(cond, slot) do { } while (0)
#endif

  if (warp == T2_TMA_WARP) {
    // ---- weight streamer: this CTA's N/2 rows of W_hi and W_lo for every chunk, up to 3 chunks ahead ----
    if (lane == 0) {
      const uint32_t half_bytes = (uint32_t)(NH * KCH * 4);
      T2Stage st;
      int j = 0;
      for (uint32_t it = 0; it < total_it; ++it) {
        if (it >= T2_NSTAGE) mbar_wait(bars + 24 + 8 * st.s, st.par ^ 1u);  // MMAs that read this stage have retired
        const float* img = g.Wimg + (long long)j * 2 * N * KCH + (long long)rank * NH * KCH;
        const uint32_t dst = base + st.s * stage_bytes + 2 * A_TILE_BYTES;
        mbar_expect_tx(bars + 48 + 8 * st.s, 2 * half_bytes);
        bulk_g2s(dst, img, half_bytes, bars + 48 + 8 * st.s);
        bulk_g2s(dst + half_bytes, img + (long long)N * KCH, half_bytes, bars + 48 + 8 * st.s);
        DBG_STAMP(true, 7);
        st.next();
        if (++j == nchunks) j = 0;
      }
    }
  } else if (warp == T2_MMA_WARP) {
    if (lane == 0) {
      T2Stage st;
      if (rank == 0) {
        // ---- MMA issuer of the pair ----
        const uint32_t idesc = make_idesc_tf32(256, N);
        const uint64_t d_a_hi = make_smem_desc(base), d_a_lo = make_smem_desc(base + A_TILE_BYTES);
        const uint64_t d_b_hi = make_smem_desc(base + 2 * A_TILE_BYTES);
        const uint64_t d_b_lo = make_smem_desc(base + 2 * A_TILE_BYTES + (uint32_t)(NH * KCH * 4));
        const uint64_t stage_inc = (uint64_t)(stage_bytes >> 4);
        int j = 0;
        for (uint32_t it = 0; it < total_it; ++it) {
          DBG_STAMP(true, 8);
          mbar_wait_cluster(bars + 48 + 8 * st.s, st.par);  // both A tiles and both weight halves are in place
          DBG_STAMP(true, 15);
          tc_fence_after();
          const uint64_t so = (uint64_t)st.s * stage_inc;
          issue_chunk_mmas_2(acc0, acc1, d_a_hi + so, d_a_lo + so, d_b_hi + so, d_b_lo + so, idesc, j == 0);
          mma_commit_2(bars + 24 + 8 * st.s);
          DBG_STAMP(true, 10);
          st.next();
          if (++j == nchunks) j = 0;
        }
      } else {
        // ---- relay: tell the leader when this CTA's operands of a chunk are in place ----
        for (uint32_t it = 0; it < total_it; ++it) {
          mbar_wait(bars + 48 + 8 * st.s, st.par);
          mbar_remote_arrive(bars + 48 + 8 * st.s, 0);
          st.next();
        }
      }
    }
  } else {
    // ---- producers: item = (point, 4 consecutive k); register prefetch of the next own chunk; see k_tc_fwd ----
    constexpr int PPR = WG * 4;
    constexpr int MAXI = (TP + PPR - 1) / PPR;
    const int group = NG > 1 ? warp / W1 : 0, wg = warp - group * W1;
    const bool active = group < NG;
    const int kq = lane & 7, psub = lane >> 3;
    // operand rows of running chunk itn (of this CTA) -> registers
    auto prefetch = [&](float4 (&buf)[MAXI][CS], uint32_t itn) {
      const uint32_t tn = itn / (uint32_t)nchunks;
      const int jr = (int)(itn - tn * (uint32_t)nchunks);
      const long long p0r = (2LL * (pair + (long long)tn * npairs) + rank) * TP;
      const int col = jr * KCH + 4 * kq;
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int pl = wg * 4 + psub + i * PPR;
        const long long p = p0r + pl;
        const bool ok = pl < TP && p < g.Np;
        const float* src = g.A.Z + p * g.A.ld + col;
#pragma unroll
        for (int c = 0; c < CS; ++c)
          buf[i][c] = ok ? __ldg(reinterpret_cast<const float4*>(src + (long long)c * g.A.plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    uint32_t it = 0;
    // one own chunk: produce from the register buffer that holds it, hand off, then refill the buffer with the own
    // chunk after next (2 NG chunks ahead).  Two statically named buffers alternate: a loaded HBM round trip was
    // measured at ~3,500 cycles, longer than one own-chunk period.  The refill comes AFTER the hand-off because
    // fence.proxy.async also waits for the thread's outstanding global loads.
    auto own_step = [&](float4 (&buf)[MAXI][CS], long long p0, int j) {
      DBG_STAMP(tid == 0, 0);
      const uint32_t s = it % T2_NSTAGE;
      unsigned char* stage_ptr = base_ptr + s * stage_bytes;
      if (it >= T2_NSTAGE) t2_wait_chunk_done(bars, it - T2_NSTAGE);  // stage free
      DBG_STAMP(tid == 0, 3);
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int pl = wg * 4 + psub + i * PPR;
        if (pl >= TP) continue;
        const long long p = p0 + pl;
        const bool valid = p < g.Np;
        float yout[CS][4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          auto comp = [&](const float4& v) { return t == 0 ? v.x : t == 1 ? v.y : t == 2 ? v.z : v.w; };
          float sc[6];
          float y0;
          act_coef<float, L::KM>(act, comp(buf[i][0]), y0, sc);
          yout[0][t] = valid ? y0 : 0.f;
#pragma unroll
          for (int d = 0; d < L::ND; ++d) {
            const int K = L::order(g.J, d), cb = L::cbase(g.J, d);
            float zz[4], yy[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) zz[o] = (o < L::KM && o < K) ? comp(buf[i][(cb + o) < CS ? (cb + o) : 0]) : 0.f;
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
      DBG_STAMP(tid == 0, 4);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 48 + 8 * s);
      DBG_STAMP(tid == 0, 5);
      if (it + 2 * NG < total_it) prefetch(buf, it + 2 * NG);
      DBG_STAMP(tid == 0, 1);
    };
    float4 zA[MAXI][CS], zB[MAXI][CS];
    if (active) {
      if ((uint32_t)group < total_it) prefetch(zA, (uint32_t)group);
      if ((uint32_t)(group + NG) < total_it) prefetch(zB, (uint32_t)(group + NG));
    }
    for (int tp = pair; tp < n_tile_pairs; tp += npairs) {
      const long long tile = 2LL * tp + rank;
      const long long p0 = tile * TP;
      for (int j = 0; j < nchunks; ++j, ++it) {
        if (!(active && (NG == 1 || (int)(it % NG) == group))) continue;  // the other group's chunk
        if ((it / NG) & 1u) own_step(zB, p0, j);
        else own_step(zA, p0, j);
      }
      // ---- epilogue (producer warps only): TMEM -> exchange tiles in the A regions of stages 0 / 1 -> Z_l ----
      {
        const uint32_t lastc = it - 1;  // the tile's last chunk
        if (active && (NG == 1 || (int)(lastc % NG) == group)) t2_wait_chunk_done(bars, lastc);  // its producers wait ...
        t2_prod_sync();                                                                           // ... and release the rest
        tc_fence_after();
        if (dbg0 && tid == 0 && it - 1 < 48) g.dbg[(it - 1) * 16 + 13] = clock64();
        const int ncb = N / 32;
        for (int cb0 = 0; cb0 < ncb; cb0 += 4) {
          // 16 (quadrant, block) fills by 14 warps: warps 2 and 3 also take block 3 of quadrants 2 and 3
#pragma unroll
          for (int rep = 0; rep < 2; ++rep) {
            const int q = warp & 3;
            const int part = rep == 0 ? (warp >> 2) : 3;
            if (rep == 1 && !(warp == 2 || warp == 3)) break;
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
          }
          if (dbg0 && tid == 0 && lastc < 48) g.dbg[lastc * 16 + (cb0 == 0 ? 2 : 12)] = clock64();  // X tiles written
          t2_prod_sync();
          if (dbg0 && tid == 0 && lastc < 48 && cb0 == 0) g.dbg[lastc * 16 + 6] = clock64();  // ... by every warp
          // write-out: one warp instruction = one output row x 4 blocks = 512 contiguous bytes (lane = block x k quad)
          for (int r = warp; r < rows_used; r += T2_NPW) {
            const int c = r / TP, pl = r - c * TP;
            const long long p = p0 + pl;
            const int b4 = lane >> 3;
            if (p < g.Np && cb0 + b4 < ncb) {
              const unsigned char* Xr = base_ptr + (b4 >> 1) * stage_bytes + (b4 & 1) * A_TILE_BYTES;
              float4 val = *reinterpret_cast<const float4*>(Xr + sw128_q(r, kq));
              if (c == 0 && g.bias) {
                const float* bp = g.bias + (cb0 + b4) * 32 + 4 * kq;
                val.x += __ldg(bp); val.y += __ldg(bp + 1); val.z += __ldg(bp + 2); val.w += __ldg(bp + 3);
              }
              *reinterpret_cast<float4*>(g.Out + (long long)c * g.oplane + p * g.ldo + (cb0 + b4) * 32 + 4 * kq) = val;
            }
          }
          if (dbg0 && tid == 0 && lastc < 48 && cb0 == 0) g.dbg[lastc * 16 + 9] = clock64();  // my rows written out
          t2_prod_sync();
          if (dbg0 && tid == 0 && lastc < 48 && cb0 == 0) g.dbg[lastc * 16 + 11] = clock64();
        }
        if (dbg0 && tid == 0 && it - 1 < 48) g.dbg[(it - 1) * 16 + 14] = clock64();
        tc_fence_before();  // accumulator reads ordered before the a_ready arrivals that release the next tile's MMAs
      }
    }
  }
#undef DBG_STAMP
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal / read this CTA
  if (warp == 1) tmem_dealloc2(acc0, ncols);
}


// cp.async pieces of a [rows_used x 32] fp32 block (row r = c*TP + pl <- plane c, point p0 + pl) for the 448
// producer threads: the piece -> (row, 16-byte column) mapping is precomputed once per thread
struct T2RowPieces {
  static constexpr int NP = (128 * 8 + T2_PROD_THREADS - 1) / T2_PROD_THREADS;
  long long src[NP];
  uint32_t dst[NP];
  int pl[NP];
  __device__ __forceinline__ void init(int TP, int rows_used, long long plane, int ld) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int i = threadIdx.x + j * T2_PROD_THREADS;
      const int r = i >> 3, q = i & 7;
      const int c = r / TP, pl_ = r - c * TP;
      pl[j] = (threadIdx.x < T2_PROD_THREADS && r < rows_used) ? pl_ : -1;
      src[j] = (long long)c * plane + (long long)pl_ * ld + q * 4;
      dst[j] = (uint32_t)(r * 128 + q * 16);
    }
  }
  __device__ __forceinline__ void issue(uint32_t dst_base, const float* Z, long long p0_ld_col0, int valid_pts) const {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (pl[j] >= 0) {
        const bool ok = pl[j] < valid_pts;
        cp_async16(dst_base + dst[j], ok ? Z + src[j] + p0_ld_col0 : Z, ok);
      }
    }
  }
};

// =====================================================================================================
// Backward dx on a CTA pair:  Abar = Zbar_l W_l^T  (M = 256 over the pair, each CTA stages half of the transposed
// weight image), then the activation adjoint  Zbar_{l-1} = adj(Abar, Z_{l-1})  in the epilogue.
// Main loop: two producer groups of 7 warps own alternate K chunks (plain hi/lo split of Zbar rows, 128-bit).
// Epilogue (all 14 producer warps), software-pipelined over the 32-column blocks of Abar: warps 8..11 move block
// cb+1 from TMEM into an exchange tile while warps 0..6 run the adjoint of block cb; the matching Z_{l-1} blocks
// arrive by cp.async three blocks ahead.  Scratch = the six A regions (all MMAs have retired):
// X[2] = stage 0 (A_hi, A_lo), Z ring[4] = stages 1, 2 (A_hi, A_lo).
// =====================================================================================================
template <class L, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_tc2_dx(TcDxArgs g) {
  static_assert(L::kStatic, "pair kernels are specialised for the static jet layouts");
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.Nout, NH = N / 2;
  const int stage_bytes = tc2_stage_bytes(N);
  const uint32_t bars_off = T2_NSTAGE * stage_bytes;
  const uint32_t bars = base + bars_off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(2 * N);
  constexpr int CS = L::CS;
  constexpr int TP = 128 / CS;
  constexpr int rows_used = CS * TP;
  constexpr int NG = 2, WG = T2_NPW / NG;  // two producer groups own alternate chunks (see k_tc2_fwd)
  const uint32_t rank = cluster_ctarank();
  // full[s]: producer warps of the chunk + weight streamer (+ the peer's relay on the leader)
  const uint32_t acc0 = tc2_setup(base, base_ptr, bars_off, stage_bytes, ncols, WG + 1 + (rank == 0 ? 1 : 0)), acc1 = acc0 + (uint32_t)N;
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int nchunks = g.Kdim / KCH;
  const int act = act_id<ACT>(g.act);
  const int n_tile_pairs = (g.num_tiles + 1) / 2;
  const int my_tp = pair < n_tile_pairs ? (n_tile_pairs - 1 - pair) / npairs + 1 : 0;
  const uint32_t total_it = (uint32_t)my_tp * (uint32_t)nchunks;
  const bool dbg0 = g.dbg && blockIdx.x == 0;
#ifdef PPSCI_B200_TIMELINE
#define DBG_STAMP(cond, slot) do { if (dbg0 && (cond) && it < 48) g.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define DBG_STAMP(cond, slot) do { } while (0)
#endif

  if (warp == T2_TMA_WARP) {
    if (lane == 0) {
      const uint32_t half_bytes = (uint32_t)(NH * KCH * 4);
      T2Stage st;
      int j = 0;
      for (uint32_t it = 0; it < total_it; ++it) {
        if (it >= T2_NSTAGE) mbar_wait(bars + 24 + 8 * st.s, st.par ^ 1u);
        const float* img = g.Wimg + (long long)j * 2 * N * KCH + (long long)rank * NH * KCH;
        const uint32_t dst = base + st.s * stage_bytes + 2 * A_TILE_BYTES;
        mbar_expect_tx(bars + 48 + 8 * st.s, 2 * half_bytes);
        bulk_g2s(dst, img, half_bytes, bars + 48 + 8 * st.s);
        bulk_g2s(dst + half_bytes, img + (long long)N * KCH, half_bytes, bars + 48 + 8 * st.s);
        DBG_STAMP(true, 7);
        st.next();
        if (++j == nchunks) j = 0;
      }
    }
  } else if (warp == T2_MMA_WARP) {
    if (lane == 0) {
      T2Stage st;
      if (rank == 0) {
        const uint32_t idesc = make_idesc_tf32(256, N);
        const uint64_t d_a_hi = make_smem_desc(base), d_a_lo = make_smem_desc(base + A_TILE_BYTES);
        const uint64_t d_b_hi = make_smem_desc(base + 2 * A_TILE_BYTES);
        const uint64_t d_b_lo = make_smem_desc(base + 2 * A_TILE_BYTES + (uint32_t)(NH * KCH * 4));
        const uint64_t stage_inc = (uint64_t)(stage_bytes >> 4);
        int j = 0;
        for (uint32_t it = 0; it < total_it; ++it) {
          DBG_STAMP(true, 8);
          mbar_wait_cluster(bars + 48 + 8 * st.s, st.par);  // both A tiles and both weight halves are in place
          DBG_STAMP(true, 15);
          tc_fence_after();
          const uint64_t so = (uint64_t)st.s * stage_inc;
          issue_chunk_mmas_2(acc0, acc1, d_a_hi + so, d_a_lo + so, d_b_hi + so, d_b_lo + so, idesc, j == 0);
          mma_commit_2(bars + 24 + 8 * st.s);
          DBG_STAMP(true, 10);
          st.next();
          if (++j == nchunks) j = 0;
        }
      } else {
        for (uint32_t it = 0; it < total_it; ++it) {
          mbar_wait(bars + 48 + 8 * st.s, st.par);
          mbar_remote_arrive(bars + 48 + 8 * st.s, 0);
          st.next();
        }
      }
    }
  } else {
    constexpr int RPP = WG * 4;
    constexpr int MAXR = (128 + RPP - 1) / RPP;
    const int group = warp / WG, wg = warp - group * WG;
    const int kq = lane & 7, psub = lane >> 3;
    auto valid_pts = [&](long long p0r) {
      const long long vp = g.Np - p0r;
      return vp >= TP ? TP : (vp > 0 ? (int)vp : 0);
    };
    auto prefetch = [&](float4 (&buf)[MAXR], uint32_t itn) {
      const uint32_t tn = itn / (uint32_t)nchunks;
      const int jr = (int)(itn - tn * (uint32_t)nchunks);
      const long long p0r = (2LL * (pair + (long long)tn * npairs) + rank) * TP;
      const int col = jr * KCH + 4 * kq;
#pragma unroll
      for (int i = 0; i < MAXR; ++i) {
        const int r = wg * 4 + psub + i * RPP;
        const int c = r / TP, pl = r - c * TP;
        const bool ok = r < rows_used && p0r + pl < g.Np;
        buf[i] = ok ? __ldg(reinterpret_cast<const float4*>(g.A.Z + (long long)c * g.A.plane + (p0r + pl) * g.A.ld + col))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    T2RowPieces zpcs;
    zpcs.init(TP, rows_used, g.zplane, g.ldz);
    uint32_t it = 0;
    auto own_step = [&](float4 (&buf)[MAXR]) {  // see k_tc2_fwd: two register buffers, refilled after the hand-off
      DBG_STAMP(tid == 0, 0);
      const uint32_t s = it % T2_NSTAGE;
      unsigned char* stage_ptr = base_ptr + s * stage_bytes;
      if (it >= T2_NSTAGE) t2_wait_chunk_done(bars, it - T2_NSTAGE);  // stage free
      DBG_STAMP(tid == 0, 3);
#pragma unroll
      for (int i = 0; i < MAXR; ++i) {
        const int r = wg * 4 + psub + i * RPP;
        if (r < rows_used) {
          const float v[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
          store_split4_at(stage_ptr, sw128_q(r, kq), v);
        }
      }
      DBG_STAMP(tid == 0, 4);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 48 + 8 * s);
      DBG_STAMP(tid == 0, 5);
      if (it + 2 * NG < total_it) prefetch(buf, it + 2 * NG);
      DBG_STAMP(tid == 0, 1);
    };
    float4 zA[MAXR], zB[MAXR];
    if ((uint32_t)group < total_it) prefetch(zA, (uint32_t)group);
    if ((uint32_t)(group + NG) < total_it) prefetch(zB, (uint32_t)(group + NG));
    for (int tp = pair; tp < n_tile_pairs; tp += npairs) {
      const long long tile = 2LL * tp + rank;
      const long long p0 = tile * TP;
      const int vpts = valid_pts(p0);
      for (int j = 0; j < nchunks; ++j, ++it) {
        if ((int)(it % NG) != group) continue;  // the other group's chunk
        if ((it / NG) & 1u) own_step(zB);
        else own_step(zA);
      }
      // ---- epilogue: two 32-column blocks of Abar per step, all 14 warps on the adjoint ----
      // Scratch = the A regions of the three stages.  The stage of the tile's LAST chunk holds the two exchange
      // tiles X0 / X1; the other two stages are free one chunk earlier (mma_done of the second-to-last chunk) and
      // hold a 4-slot ring of Z_{l-1} blocks, so the first two block pairs are already in flight while the
      // last chunk's MMAs run.  Step i: warps 0..3 / 7..10 move blocks 2i / 2i+1 from TMEM to X0 / X1, then warps
      // 0..6 run the adjoint of block 2i and warps 7..13 of block 2i+1; the Z pair i+1 is requested at step i.
      {
        const uint32_t lastc = it - 1;
        const int ncb = N / 32;
        const int nsteps = (ncb + 1) / 2;
        const uint32_t sc = lastc % T2_NSTAGE, sa = (lastc + 1) % T2_NSTAGE, sb = (lastc + 2) % T2_NSTAGE;
        auto xbuf = [&](int i) { return base_ptr + sc * stage_bytes + (i & 1) * A_TILE_BYTES; };
        auto zslot = [&](int i) { return (uint32_t)(((i & 2) ? sb : sa) * stage_bytes + (i & 1) * A_TILE_BYTES); };
        auto issue_pair = [&](int i) {  // Z blocks 2i, 2i+1 -> ring slots 2(i&1), 2(i&1)+1; one cp.async group
          if (i < nsteps) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int cb = 2 * i + h;
              if (cb < ncb) zpcs.issue(base + zslot(2 * (i & 1) + h), g.Zprev, p0 * g.ldz + cb * 32, vpts);
            }
          }
          cp_async_commit();
        };
        if (lastc >= 1) {  // the two other stages are no longer read once the second-to-last chunk has retired
          if ((int)((lastc - 1) % NG) == group) t2_wait_chunk_done(bars, lastc - 1);  // (waited by its producers,
          t2_prod_sync();                                                              //  see t2_wait_chunk_done)
        }
        issue_pair(0);
        issue_pair(1);
        if ((int)(lastc % NG) == group) t2_wait_chunk_done(bars, lastc);
        t2_prod_sync();
        tc_fence_after();
        if (dbg0 && tid == 0 && lastc < 48) g.dbg[lastc * 16 + 13] = clock64();
        const int grp = warp / 7, wg7 = warp - grp * 7;  // adjoint group (block parity) and warp within it
        for (int i = 0; i < nsteps; ++i) {
          if (i == 0) cp_async_wait<1>();  // Z pair i has landed (at step 0 pair 1 may still be in flight;
          else cp_async_wait<0>();         //  later pairs are requested one step ahead)
          t2_prod_sync();      // ... for every thread; all warps are done with step i-1 (X and its Z slots are free)
          if (i >= 1) issue_pair(i + 1);  // -> the slots of pair i-1
          {
            const int fw = grp == 0 ? warp : warp - 7;  // warps 0..3 and 7..10 fill X0 / X1
            const int cb = 2 * i + grp;
            if (fw < 4 && cb < ncb) {
              const int q = warp & 3;  // (7..10) & 3 = 3, 0, 1, 2: all four lane quadrants
              float v[32];
              load_acc_sum(acc0, acc1, q, cb * 32, v, tc_rz_comp_split(nchunks));
              unsigned char* Xb = xbuf(grp);
              const int row = q * 32 + lane;
#pragma unroll
              for (int t4 = 0; t4 < 8; ++t4)
                *reinterpret_cast<float4*>(Xb + sw128_q(row, t4)) = make_float4(v[4 * t4], v[4 * t4 + 1], v[4 * t4 + 2], v[4 * t4 + 3]);
            }
          }
          t2_prod_sync();  // X0 / X1 complete
          const int cb = 2 * i + grp;
          if (cb < ncb) {
            const unsigned char* zb = base_ptr + zslot(2 * (i & 1) + grp);
            const unsigned char* Xb = xbuf(grp);
            for (int pl = wg7 * 4 + psub; pl < vpts; pl += 28) {  // item = (point, k quad)
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
                float sc_[6];
                float y0;
                act_coef<float, L::KM + 1>(act, comp(zc[0]), y0, sc_);
                float sb_[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
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
                  jet_adj_dir<float, L::KM>(sc_, zz, yb, zbv, sb_);
#pragma unroll
                  for (int o = 0; o < L::KM; ++o)
                    if (o < K && cbs + o < CS) ob[cbs + o][t] = zbv[o];
                }
                ob[0][t] = jet_adj_z0<float, L::KM>(sc_, comp(xc[0]), sb_);
              }
              float* out = g.Out + (p0 + pl) * g.ldo + cb * 32 + 4 * kq;
#pragma unroll
              for (int c = 0; c < CS; ++c)
                *reinterpret_cast<float4*>(out + (long long)c * g.oplane) = make_float4(ob[c][0], ob[c][1], ob[c][2], ob[c][3]);
            }
          }
        }
        cp_async_wait<0>();
        if (dbg0 && tid == 0 && lastc < 48) g.dbg[lastc * 16 + 14] = clock64();
        tc_fence_before();
        t2_prod_sync();  // all scratch reads done before the A regions are produced into again
      }
    }
  }
#undef DBG_STAMP
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(acc0, ncols);
}


// =====================================================================================================
// Backward dW on a CTA pair:  dW_l[k][n] += sum_rows A_{l-1}[row][k] Zbar_l[row][n]  for a 256 x 256 block of dW
// (M = 256 = the block's k rows, 128 per CTA; N = 256, each CTA stages the 128 n rows of its half) over this
// pair's range of 32-row reduction chunks.  Producers exactly as in k_tc_dw (16 warps, one 4x4-block task each:
// 8 tasks transpose-split the a-stash block, 8 the Zbar block; registers double-buffered two chunks ahead), but
// each CTA now feeds twice the flops per byte it pulls from L2.  17th warp: MMA issue (leader) / relay (peer).
// grid (2 * K/256, splits, N/256), cluster (2, 1, 1).
// =====================================================================================================
template <class L>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DW_THREADS, 1) k_tc2_dw(TcDwArgs g) {
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  constexpr int N = 256, NH = 128;
  const int stage_bytes = tc2_stage_bytes(N);
  const uint32_t bars_off = T2_NSTAGE * stage_bytes;
  const uint32_t bars = base + bars_off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = 512;
  const int PT = L::pt(g.PT);
  const int C = L::nchan(g.J);
  const int rows_used = C * PT;
  const uint32_t rank = cluster_ctarank();
  const uint32_t acc0 = tc2_setup(base, base_ptr, bars_off, stage_bytes, ncols, DW_NPW + (rank == 0 ? 1 : 0)), acc1 = acc0 + (uint32_t)N;
  const int k0 = (int)(blockIdx.x >> 1) * 256 + (int)rank * 128;  // this CTA's dW rows
  const int n0 = (int)blockIdx.z * 256;
  const int n0h = n0 + (int)rank * NH;                             // this CTA's half of the B operand
  const long long total_chunks = (g.Np + PT - 1) / PT;
  const long long ch_begin = (long long)blockIdx.y * g.chunks_per_split;
  long long ch_end = ch_begin + g.chunks_per_split;
  if (ch_end > total_chunks) ch_end = total_chunks;
  const uint32_t n_it = ch_begin < ch_end ? (uint32_t)(ch_end - ch_begin) : 0u;
  auto valid_pts = [&](long long ch) {
    const long long vp = g.Np - ch * PT;
    return vp >= PT ? PT : (vp > 0 ? (int)vp : 0);
  };
  const bool is_mma = (warp == DW_MMA_WARP);
  const bool dbg0 = g.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#ifdef PPSCI_B200_TIMELINE
#define DBG_STAMP(cond, slot) do { if (dbg0 && (cond) && it < 48) g.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define DBG_STAMP(cond, slot) do { } while (0)
#endif

  if (is_mma) {
    if (lane == 0) {
      T2Stage st;
      if (rank == 0) {
        const uint32_t idesc = make_idesc_tf32(256, N);
        const uint64_t d_a_hi = make_smem_desc(base), d_a_lo = make_smem_desc(base + A_TILE_BYTES);
        const uint64_t d_b_hi = make_smem_desc(base + 2 * A_TILE_BYTES);
        const uint64_t d_b_lo = make_smem_desc(base + 2 * A_TILE_BYTES + (uint32_t)(NH * KCH * 4));
        const uint64_t stage_inc = (uint64_t)(stage_bytes >> 4);
        for (uint32_t it = 0; it < n_it; ++it) {
          DBG_STAMP(true, 8);
          mbar_wait_cluster(bars + 48 + 8 * st.s, st.par);  // both A tiles and both weight halves are in place
          DBG_STAMP(true, 15);
          tc_fence_after();
          const uint64_t so = (uint64_t)st.s * stage_inc;
          issue_chunk_mmas_2(acc0, acc1, d_a_hi + so, d_a_lo + so, d_b_hi + so, d_b_lo + so, idesc, it == 0);
          mma_commit_2(bars + 24 + 8 * st.s);
          DBG_STAMP(true, 10);
          st.next();
        }
      } else {
        for (uint32_t it = 0; it < n_it; ++it) {
          mbar_wait(bars + 48 + 8 * st.s, st.par);
          mbar_remote_arrive(bars + 48 + 8 * st.s, 0);
          st.next();
        }
      }
    }
    __syncwarp();
  } else {
    // task geometry (see k_tc_dw): warps 0..7 -> A' (this CTA's 128 k rows), warps 8..15 -> B' (its 128 n rows)
    const bool t_isA = warp < 8;
    uint32_t g_off[4], d_off[4];
    uint32_t g_plb = 0xFFFFFFFFu;
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
        if (rr < rows_used) {
          const int c = rr / PT, pl = rr - c * PT;
          b8 = (uint32_t)pl;
          g_off[e] = t_isA ? (uint32_t)((long long)c * g.aplane + (long long)pl * g.lda + k0 + row0)
                           : (uint32_t)((long long)c * g.zbplane + (long long)pl * g.ldzb + n0h + row0);
        }
        plb |= b8 << (8 * e);
      }
      g_plb = plb;
#pragma unroll
      for (int i = 0; i < 4; ++i) d_off[i] = sw128_q(row0 + i, q) + (t_isA ? 0u : (uint32_t)(2 * A_TILE_BYTES));
    }
    const int d_lo = t_isA ? A_TILE_BYTES : NH * KCH * 4;
    // fused bias gradient: B' tasks of the first k block whose 4 reduction rows include value-channel rows (rr < PT)
    const int db_q0 = 4 * ((warp & 1) * 4 + ((lane >> 1) & 3));
    const bool db_on = g.db != nullptr && !t_isA && (blockIdx.x >> 1) == 0 && db_q0 < PT;
    const int db_col = n0h + ((warp - 8) >> 1) * 32 + ((((lane >> 3) << 1) | (lane & 1)) * 4);
    float4 db_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto prefetch = [&](float4 (&buf)[4], long long ch) {
      const uint32_t vp = (uint32_t)valid_pts(ch);
      const float* bp = t_isA ? g.Aact + ch * PT * (long long)g.lda : g.Zbar + ch * PT * (long long)g.ldzb;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = ((g_plb >> (8 * e)) & 255u) < vp;
        buf[e] = ok ? __ldg(reinterpret_cast<const float4*>(bp + g_off[e])) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // Rows of chunk `ch` pulled towards L2 (one request per 128-byte line: the lanes whose 16-byte piece starts it).
    // The register prefetch runs two chunks ahead, i.e. about one chunk time (~2.5 k cycles) between the request and
    // the first use — less than a loaded HBM round trip: the ncu source page of round 2 had 19 % of ALL warp samples of
    // this kernel on the first use of the prefetched registers (long scoreboard).  With the lines already in L2 the
    // same request returns in a fraction of that.
    const bool l2_lane = (((lane >> 3) << 1) | (lane & 1)) == 0;
    auto prefetch_far = [&](long long ch) {
      if (!l2_lane || ch >= ch_end) return;
      const uint32_t vp = (uint32_t)valid_pts(ch);
      const float* bp = t_isA ? g.Aact + ch * PT * (long long)g.lda : g.Zbar + ch * PT * (long long)g.ldzb;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (((g_plb >> (8 * e)) & 255u) < vp) prefetch_l2(bp + g_off[e]);
    };
    auto step = [&](float4 (&buf)[4], long long ch, uint32_t it) {
      const uint32_t s = it % T2_NSTAGE;
      DBG_STAMP(tid == 0, 0);
      unsigned char* stage_ptr = base_ptr + s * stage_bytes;
      if (it >= T2_NSTAGE) t2_wait_chunk_done(bars, it - T2_NSTAGE);  // MMAs that read this stage have retired
      DBG_STAMP(tid == 0, 3);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v[4] = {f4comp(buf[0], i), f4comp(buf[1], i), f4comp(buf[2], i), f4comp(buf[3], i)};
        float4 h, l;
        h.x = tf32_rn(v[0]); h.y = tf32_rn(v[1]); h.z = tf32_rn(v[2]); h.w = tf32_rn(v[3]);
        l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
        *reinterpret_cast<float4*>(stage_ptr + d_off[i]) = h;
        *reinterpret_cast<float4*>(stage_ptr + d_off[i] + d_lo) = l;
      }
      if (db_on) {  // value-channel rows of the Zbar block are exactly the bias-gradient terms of my 4 columns
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (db_q0 + e < PT) {
            db_acc.x += buf[e].x; db_acc.y += buf[e].y; db_acc.z += buf[e].z; db_acc.w += buf[e].w;
          }
      }
      DBG_STAMP(tid == 0, 4);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 48 + 8 * s);
      DBG_STAMP(tid == 0, 5);
      if (ch + 2 < ch_end) prefetch(buf, ch + 2);
      prefetch_far(ch + 5);
      DBG_STAMP(tid == 0, 6);
    };
    float4 bufA[4], bufB[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) bufA[e] = bufB[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch_begin < ch_end) prefetch(bufA, ch_begin);
    if (ch_begin + 1 < ch_end) prefetch(bufB, ch_begin + 1);
    for (int a = 2; a < 5; ++a) prefetch_far(ch_begin + a);
    uint32_t it = 0;
    for (long long ch = ch_begin; ch < ch_end; ch += 2, it += 2) {
      step(bufA, ch, it);
      if (ch + 1 < ch_end) step(bufB, ch + 1, it + 1);
    }
    if (db_on && n_it > 0) {
      atomicAdd(g.db + db_col, db_acc.x);
      atomicAdd(g.db + db_col + 1, db_acc.y);
      atomicAdd(g.db + db_col + 2, db_acc.z);
      atomicAdd(g.db + db_col + 3, db_acc.w);
    }
    // ---- flush: this CTA's 128 x 256 block of partial dW ----
    if (n_it > 0) {
      t2_wait_chunk_done(bars, n_it - 1);
      tc_fence_after();
      const int q = warp & 3, part = warp >> 2;
      const int k = k0 + q * 32 + lane;
      float* dw_row = g.dW + (long long)k * g.ldw + n0;
      for (int cb = part; cb < N / 32; cb += DW_NPW / 4) {
        float v[32];
        load_acc_sum(acc0, acc1, q, cb * 32, v, tc_rz_comp_split(n_it));
        if (k < g.Kdim) {
#pragma unroll
          for (int t = 0; t < 32; ++t) atomicAdd(dw_row + cb * 32 + t, v[t]);
        }
      }
    }
  }
#undef DBG_STAMP
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(acc0, ncols);
}

}  // namespace tc
}  // namespace ppsci

#define PPSCI_TC2_LAUNCH_L(KERNEL, lay, grid, smem, stream, args, err_expr)                                         \
  do {                                                                                                              \
    void (*kfn_)(decltype(args)) = nullptr;                                                                         \
    switch (lay) {                                                                                                  \
      case ppsci::TC_LAY_22: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<2, 2, 0, 0>>; break;                          \
      case ppsci::TC_LAY_12: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<1, 2, 0, 0>>; break;                          \
      case ppsci::TC_LAY_222: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<2, 2, 2, 0>>; break;                         \
      case ppsci::TC_LAY_VALUE: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<0, 0, 0, 0>>; break;                       \
      default: kfn_ = ppsci::tc::KERNEL<ppsci::tc::DLay<4>>;                                                        \
    }                                                                                                               \
    cudaError_t e_ = cudaFuncSetAttribute(kfn_, cudaFuncAttributeMaxDynamicSharedMemorySize, (smem));               \
    if (e_ != cudaSuccess) { err_expr; }                                                                            \
    PPSCI_KLAUNCH(kfn_, (grid), dim3(ppsci::tc::DW_THREADS), (smem), (stream), 2, args);                            \
  } while (0)

// static jet layouts only (callers check lay != TC_LAY_DYN); tanh activation
#define PPSCI_TC2_LAUNCH(KERNEL, lay, grid, smem, stream, args, err_expr)                                           \
  do {                                                                                                              \
    void (*kfn_)(decltype(args)) = nullptr;                                                                         \
    switch (lay) {                                                                                                  \
      case ppsci::TC_LAY_22: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<2, 2, 0, 0>, PPSCI_ACT_TANH>; break;          \
      case ppsci::TC_LAY_12: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<1, 2, 0, 0>, PPSCI_ACT_TANH>; break;          \
      case ppsci::TC_LAY_222: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<2, 2, 2, 0>, PPSCI_ACT_TANH>; break;         \
      default: kfn_ = ppsci::tc::KERNEL<ppsci::tc::SLay<0, 0, 0, 0>, PPSCI_ACT_TANH>;                               \
    }                                                                                                               \
    cudaError_t e_ = cudaFuncSetAttribute(kfn_, cudaFuncAttributeMaxDynamicSharedMemorySize, (smem));               \
    if (e_ != cudaSuccess) { err_expr; }                                                                            \
    PPSCI_KLAUNCH(kfn_, (grid), dim3(ppsci::tc::THREADS), (smem), (stream), 2, args);                               \
  } while (0)
