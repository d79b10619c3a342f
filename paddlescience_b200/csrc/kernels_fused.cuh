// kernels_fused.cuh — layer-FUSED tensor-core kernels (CTA pairs, tcgen05 cta_group::2).
//
// What this replaces in the reference: MLP.forward_tensor's loop over hidden layers (ppsci/arch/mlp.py:281-296),
// which — like the round-1 layer-at-a-time kernels here — materialises every [N, H] activation in HBM and re-reads
// it for the next layer.  A tile's jets now stay ON CHIP from layer to layer:
//
//   TMEM accumulator of layer l  --tcgen05.ld-->  registers  --(+bias, rz compensation)-->  exchange tile in SMEM
//     --(point, k-quad) items-->  act_jets  -->  hi/lo tf32 split  -->  swizzled A operand of layer l+1 (SMEM ring)
//
// and only the stash the adjoint needs goes to HBM, written ONCE by the same epilogue pass: Z_l (pre-activations)
// and a_l = act_jets(Z_l) (the dW operand).  Nothing is re-read on the forward.
//
// Overlap.  Two 256-column accumulators (single fp32 accumulator per layer, see below) alternate between layers, so
// the epilogue of layer l — which PRODUCES the A chunks of layer l+1 one 32-wide K chunk at a time — runs while the
// MMAs of layer l+1 consume them: chunk j of layer l+1 is issued as soon as block j of layer l has been drained.
// The tensor pipe only idles while the first K chunk of a layer is being produced.
//
// Accuracy.  One accumulator per layer means 12 round-toward-zero accumulations per 32-wide chunk at full magnitude;
// the measured shrink (c0 = 3.35e-8 per accumulation, tests/microbench/tc_numerics.cu) is a data-independent factor
// and is compensated in the epilogue (tc_rz_comp_single) — what is left is the zero-mean part of the rounding.
//
// Roles per CTA (512 threads) and barriers are those of kernels_tc2.cuh: warps 0..13 workers (operand producers +
// epilogue), warp 14 lane 0 weight streamer (three chunks ahead, across layers and tiles), warp 15 lane 0 MMA issuer
// (leader) / relay (peer); full[s] collects 14 worker arrivals + the weight bytes (+ the peer's relay on the leader),
// mma_done[s] is the multicast tcgen05.commit.  EVERY worker warp arrives on EVERY chunk, and only after it has seen
// mma_done of the chunk three before (= the previous use of that stage and of its full barrier): a warp can never
// arrive into a stale phase, and no parity wait is ever more than one phase away from its target.
#pragma once
#include "kernels_tc2.cuh"

namespace ppsci {
namespace tc {

constexpr int FUSE_MAXL = 8;
constexpr int FUSE_X_BYTES = 128 * 32 * 4;  // one [128 x 32] fp32 exchange tile

struct FusedFwdArgs {
  JetLayout J;
  int act;
  int H;                        // hidden width: K = N = H for every fused layer
  int n_fused;                  // number of fused hidden -> hidden layers
  const float* Zin;             // Z of the layer below the first fused one: [C][nc_max][ld]
  int ld;                       // row pitch of every plane set (round4(H) = H)
  long long plane;              // plane stride (nc_max * ld)
  const float* Wimg[FUSE_MAXL];  // swizzled hi / lo weight images (k_tc_prep_w), forward orientation
  const float* bias[FUSE_MAXL];
  float* Zout[FUSE_MAXL];       // pre-activations of fused layer i: the adjoint's stash (and the next kernel's input)
  float* Astash[FUSE_MAXL];     // act_jets(input of fused layer i) for the dW kernels; null: forward only
  long long Np;
  int num_tiles;
  long long* dbg;
};

// single accumulator: 12 accumulations per 32-wide chunk, all at the full magnitude of the result
__host__ __device__ __forceinline__ float tc_rz_comp_single(long long nchunks) { return 1.f + TC_RZ_C0 * 6.f * (float)nchunks; }

// 12 MMAs (3xTF32) of one K chunk into ONE accumulator, M = 256 over the CTA pair
__device__ __forceinline__ void issue_chunk_mmas_2_single(uint32_t acc, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                          uint32_t idesc, bool first_chunk) {
#pragma unroll
  for (int ks = 0; ks < KCH / 8; ++ks) {
    const uint64_t inc = (uint64_t)(2 * ks);
    mma_tf32_2(acc, a_hi + inc, b_hi + inc, idesc, (first_chunk && ks == 0) ? 0u : 1u);
    mma_tf32_2(acc, a_lo + inc, b_hi + inc, idesc, 1u);
    mma_tf32_2(acc, a_hi + inc, b_lo + inc, idesc, 1u);
  }
}

// Shared-memory map of the fused kernels (offsets from the 1024-aligned base).  The A operand ring is FOUR stages deep
// and separate from the TWO weight slots: an epilogue step produces two K chunks at once, so with the two chunks of the
// previous step still being read by the tensor cores the producers need two MORE free stages to run a full step ahead
// (timeline of the first version, three combined stages: the workers idled ~2,600 cycles per step waiting for a stage
// and the MMAs started ~1,700 cycles after the hand-off — nothing overlapped).  Weights only need double buffering:
// a 32 KB half-chunk arrives from L2 in ~800 cycles, a chunk's MMAs take ~1,900.
// Second version (4 A stages + 2 weight slots + 2 exchange tiles): the stage wait disappeared but the MMAs still
// started 2,500 - 3,800 cycles after the hand-off — with two weight slots a chunk's weights can only be requested when
// the chunk two before has retired, one chunk time (1,900 cycles) before they are needed, and the loaded L2 -> SMEM
// round trip is longer than that (round-1 finding: weights must be three chunks ahead).  Three weight slots need the
// 32 KB of the exchange tiles back, so the exchange now happens IN PLACE: the raw fp32 accumulator block of 32
// columns is written into the A_lo region of the very stage its operand chunk will occupy, and every (point, k-quad)
// item reads exactly the 16-byte cells it later overwrites with its own hi / lo output.
//   [0, 4 * 32 KB)            A stages: A_hi | A_lo of chunk it % 4   (A_lo doubles as the exchange tile)
//   then 3 weight slots       B_hi half | B_lo half of chunk it % 3   (N/2 rows each)
//   then barriers: mma_done[4] +0.., full[4] +64.., TMEM base slot +192
constexpr int FA = 4;
constexpr int FB = 3;
__host__ __device__ inline int fused_bslot_bytes(int N) { return 2 * (N / 2) * KCH * 4; }
__host__ __device__ inline int fused_fwd_smem_bytes(int N) {
  return FA * 2 * A_TILE_BYTES + FB * fused_bslot_bytes(N) + 1024 + 256;
}
// chunk c has retired (its A stage, its weight slot and its full barrier may be reused)
__device__ __forceinline__ void f_wait_done(uint32_t bars, uint32_t c) { mbar_wait_warp(bars + 8 * (c % FA), (c / FA) & 1u); }
__device__ __forceinline__ void f_wait_done_lane(uint32_t bars, uint32_t c) { mbar_wait(bars + 8 * (c % FA), (c / FA) & 1u); }
// Workers keep a watermark `seen`: every chunk below it is known to have retired (tcgen05.commit arrives in issue
// order, so observing chunk c covers all earlier ones).  A stage wait whose chunk lies below the watermark is skipped:
// the two-CTA timeline of round 2 showed 350 - 900 cycles per poll even for a long-completed phase (divergent lane-0
// poll + reconvergence), on the workers' critical path four times per layer.  Polls only ever move forward by at
// most one phase per barrier (the chunk FA before the awaited one is always below the watermark), so the parity
// test cannot alias.
__device__ __forceinline__ void f_wait_done_seen(uint32_t bars, uint32_t c, uint32_t& seen) {
  if (c >= seen) {
    f_wait_done(bars, c);
    seen = c + 1;
  }
}
// Layer-start wait (the accumulator of the layer is final): every worker warp is idle until then, so lane 0 probes
// with the non-suspending test_wait instead of try_wait, whose suspended wait is woken noticeably later
// (the timeline showed ~1.5 k cycles between the last MMA retiring and the workers resuming).
__device__ __forceinline__ void f_spin_done_seen(uint32_t bars, uint32_t c, uint32_t& seen) {
#ifdef PPSCI_EMUL
  f_wait_done_seen(bars, c, seen);  // OS threads: the yielding wait
  return;
#endif
  if (c >= seen) {
    if ((threadIdx.x & 31) == 0) {
      uint32_t spins = 0;
      while (!mbar_test_wait(bars + 8 * (c % FA), (c / FA) & 1u)) {
        if (++spins > (1u << 26)) __trap();
      }
    }
    __syncwarp();
    seen = c + 1;
  }
}

// Shared prologue: barriers, pair-wide TMEM allocation, zeroed A stages, cluster rendezvous.
__device__ __forceinline__ uint32_t fused_setup(uint32_t base, unsigned char* base_ptr, uint32_t bars_off, uint32_t ncols, int full_count) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bars = base + bars_off;
  if (tid == 0) {
    for (int i = 0; i < FA; ++i) {
      mbar_init(bars + 8 * i, 1);                // mma_done
      mbar_init(bars + 64 + 8 * i, full_count);  // full
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc2(bars + 192, ncols);
    tmem_relinquish2();
  }
  float4* az = reinterpret_cast<float4*>(base_ptr);
  for (int i = tid; i < FA * 2 * A_TILE_BYTES / 16; i += (int)blockDim.x) az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(base_ptr + bars_off + 192);
}

// weight streamer lane: this CTA's N/2 rows of W_hi and W_lo of every (tile, layer, chunk), three chunks ahead
__device__ __forceinline__ void fused_stream_weights(uint32_t base, uint32_t bars, uint32_t b_off, const float* const* Wimg, int my_tp,
                                                     int NLf, int nchunks, int N, uint32_t rank) {
  const int NH = N / 2;
  const uint32_t half_bytes = (uint32_t)(NH * KCH * 4);
  uint32_t it = 0;
  for (int t = 0; t < my_tp; ++t)
    for (int l = 0; l < NLf; ++l) {
      const float* wl = Wimg[l] + (long long)rank * NH * KCH;
      for (int j = 0; j < nchunks; ++j, ++it) {
        if (it >= FB) f_wait_done_lane(bars, it - FB);  // the MMAs that read this weight slot have retired
        const float* img = wl + (long long)j * 2 * N * KCH;
        const uint32_t dst = base + b_off + (it % FB) * (2 * half_bytes);
        const uint32_t fb = bars + 64 + 8 * (it % FA);
        mbar_expect_tx(fb, 2 * half_bytes);
        bulk_g2s(dst, img, half_bytes, fb);
        bulk_g2s(dst + half_bytes, img + (long long)N * KCH, half_bytes, fb);
      }
    }
}

// MMA issuer (leader) / relay (peer) lane
template <class DBG>
__device__ __forceinline__ void fused_issue_mmas(uint32_t base, uint32_t bars, uint32_t b_off, uint32_t acc_base, int my_tp, int NLf,
                                                 int nchunks, int N, uint32_t rank, DBG&& dbg) {
  const int NH = N / 2;
  const uint32_t total_it = (uint32_t)my_tp * (uint32_t)(NLf * nchunks);
  if (rank == 0) {
    const uint32_t idesc = make_idesc_tf32(256, N);
    const uint64_t d_a_hi = make_smem_desc(base), d_a_lo = make_smem_desc(base + A_TILE_BYTES);
    const uint64_t d_b_hi = make_smem_desc(base + b_off);
    const uint64_t d_b_lo = make_smem_desc(base + b_off + (uint32_t)(NH * KCH * 4));
    const uint64_t a_inc = (uint64_t)((2 * A_TILE_BYTES) >> 4), b_inc = (uint64_t)(fused_bslot_bytes(N) >> 4);
    uint32_t lay = 0, it = 0;
    for (int t = 0; t < my_tp; ++t)
      for (int l = 0; l < NLf; ++l, ++lay) {
        const uint32_t acc = acc_base + (lay & 1u) * (uint32_t)N;
        for (int j = 0; j < nchunks; ++j, ++it) {
          dbg(it, 8);
          mbar_wait_cluster(bars + 64 + 8 * (it % FA), (it / FA) & 1u);  // both A tiles and both weight halves are in place
          dbg(it, 9);
          tc_fence_after();
          const uint64_t ao = (uint64_t)(it % FA) * a_inc, bo = (uint64_t)(it % FB) * b_inc;
          issue_chunk_mmas_2_single(acc, d_a_hi + ao, d_a_lo + ao, d_b_hi + bo, d_b_lo + bo, idesc, j == 0);
          mma_commit_2(bars + 8 * (it % FA));
          dbg(it, 10);
        }
      }
  } else {
    for (uint32_t it = 0; it < total_it; ++it) {  // relay: tell the leader when this CTA's operands of a chunk are in place
      mbar_wait(bars + 64 + 8 * (it % FA), (it / FA) & 1u);
      mbar_remote_arrive(bars + 64 + 8 * (it % FA), 0);
    }
  }
}

// activation jets of 4 consecutive hidden units of one point: y[c][t] from z[c] (float4 along k); zeros if !valid
template <class L>
__device__ __forceinline__ void act_jets4(int act, const JetLayout& J, const float4 (&z)[L::CS], bool valid, float (&yout)[L::CS][4]) {
  constexpr int CS = L::CS;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float sc[6];
    float y0;
    act_coef<float, L::KM>(act, f4comp(z[0], t), y0, sc);
    yout[0][t] = valid ? y0 : 0.f;
#pragma unroll
    for (int d = 0; d < L::ND; ++d) {
      const int K = L::order(J, d), cb = L::cbase(J, d);
      float zz[4], yy[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) zz[o] = (o < L::KM && o < K) ? f4comp(z[(cb + o) < CS ? (cb + o) : 0], t) : 0.f;
      jet_fwd_dir<float, L::KM>(sc, zz, yy);
#pragma unroll
      for (int o = 0; o < L::KM; ++o)
        if (o < K && cb + o < CS) yout[cb + o][t] = valid ? yy[o] : 0.f;
    }
  }
}

// =====================================================================================================
// Fused forward over the hidden -> hidden layers of one point chunk.  CTA rank r of pair q handles point tiles
// 2 t + r for t = q, q + pairs, ...  Per tile:  P0 (operand of the first fused layer from Zin in HBM, register
// prefetch two K chunks ahead)  ->  for every fused layer: epilogue = stash Z_l (+ a_l) and produce the next layer's
// operand chunks.  Static jet layouts, tanh.
// =====================================================================================================
template <class L, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_fused_fwd(FusedFwdArgs g) {
  static_assert(L::kStatic, "fused kernels are specialised for the static jet layouts");
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.H;
  const uint32_t b_off = (uint32_t)(FA * 2 * A_TILE_BYTES);
  const uint32_t bars_off = b_off + (uint32_t)(FB * fused_bslot_bytes(N));
  const uint32_t bars = base + bars_off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(2 * N);
  constexpr int CS = L::CS;
  constexpr int TP = 128 / CS;
  const uint32_t rank = cluster_ctarank();
  // full[s]: every worker warp + weight streamer (+ the peer's relay on the leader)
  const uint32_t acc_base = fused_setup(base, base_ptr, bars_off, ncols, T2_NPW + 1 + (rank == 0 ? 1 : 0));
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int nchunks = N / KCH;  // K = N = H
  const int ncb = N / 32;
  const int NLf = g.n_fused;
  const int act = act_id<ACT>(g.act);
  const int n_tile_pairs = (g.num_tiles + 1) / 2;
  const int my_tp = pair < n_tile_pairs ? (n_tile_pairs - 1 - pair) / npairs + 1 : 0;
  const bool dbg0 = g.dbg && blockIdx.x == 0;
#ifdef PPSCI_B200_TIMELINE
#define FDBG(cond, row, slot) do { if (dbg0 && (cond) && (row) < 48u) g.dbg[(row) * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define FDBG(cond, row, slot) do { } while (0)
#endif

  if (warp == T2_TMA_WARP) {
    if (lane == 0) fused_stream_weights(base, bars, b_off, g.Wimg, my_tp, NLf, nchunks, N, rank);
  } else if (warp == T2_MMA_WARP) {
    if (lane == 0)
      fused_issue_mmas(base, bars, b_off, acc_base, my_tp, NLf, nchunks, N, rank, [&](uint32_t row, int slot) { FDBG(true, row, slot); });
  } else {
    // ---- workers: two groups of 7 warps; a step = two K chunks, group g owns the odd / even one ----
    const int grp = warp / 7, wg7 = warp - grp * 7;
    const int kq = lane & 7, psub = lane >> 3;
    const int pl0 = wg7 * 4 + psub;               // this thread's point in the first item pass (28 points per pass)
    constexpr int MAXI = (TP + 27) / 28;
    const float comp = tc_rz_comp_single(nchunks);
    uint32_t it = 0, lay = 0, seen = 0;
    // hand-off of the two chunks of a step: every warp arrives on both full barriers (see the file header)
    auto hand_off = [&](uint32_t it0, bool has1) {
      fence_proxy_async();
      tc_fence_before();  // this warp's accumulator reads are ordered before the MMAs its arrivals release
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bars + 64 + 8 * (it0 % FA));
        if (has1) mbar_arrive(bars + 64 + 8 * ((it0 + 1) % FA));
      }
    };
    // both stages of a step are free once the LATER of the two chunks that used them has retired (tcgen05.commit
    // completes in issue order): one barrier poll per step instead of two (each costs ~250 cycles even when complete)
    auto wait_stages = [&](uint32_t it0, bool has1) {
      const uint32_t last = has1 ? it0 + 1 : it0;
      if (last >= FA) f_wait_done_seen(bars, last - FA, seen);
    };
    for (int tp = pair; tp < n_tile_pairs; tp += npairs) {
      const long long tile = 2LL * tp + rank;
      const long long p0 = tile * TP;
      const long long vleft = g.Np - p0;
      const int vpts = vleft >= TP ? TP : (vleft > 0 ? (int)vleft : 0);
      // ================= P0: operand of the first fused layer = act_jets(Zin) =================
      {
        auto prefetch = [&](float4 (&buf)[MAXI][CS], int j) {
          const int col = j * KCH + 4 * kq;
#pragma unroll
          for (int i = 0; i < MAXI; ++i) {
            const int pl = pl0 + i * 28;
            const bool ok = pl < vpts;
            const float* src = g.Zin + (p0 + pl) * g.ld + col;
#pragma unroll
            for (int c = 0; c < CS; ++c)
              buf[i][c] = ok ? __ldg(reinterpret_cast<const float4*>(src + (long long)c * g.plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        // one step: chunks j0, j0 + 1; this group's chunk (j0 + grp) comes out of `buf`, which is then refilled with
        // the group's chunk after next (4 chunks = two steps ahead; fence.proxy.async waits for outstanding loads,
        // so the refill is issued after the hand-off)
        auto step = [&](float4 (&buf)[MAXI][CS], int j0) {
          const bool has1 = j0 + 1 < nchunks;
          const int j = j0 + grp;
          FDBG(tid == 0, it, 0);
          wait_stages(it, has1);
          FDBG(tid == 0, it, 4);
          if (j < nchunks) {
            unsigned char* stage_ptr = base_ptr + ((it + (uint32_t)grp) % FA) * (2 * A_TILE_BYTES);
#pragma unroll
            for (int i = 0; i < MAXI; ++i) {
              const int pl = pl0 + i * 28;
              if (pl >= TP) continue;
              const bool valid = pl < vpts;
              float yout[CS][4];
              act_jets4<L>(act, g.J, buf[i], valid, yout);
              float* ast = (g.Astash[0] && valid) ? g.Astash[0] + (p0 + pl) * g.ld + j * KCH + 4 * kq : nullptr;
#pragma unroll
              for (int c = 0; c < CS; ++c) {
                store_split4_at(stage_ptr, sw128_q(c * TP + pl, kq), yout[c]);
                if (ast) *reinterpret_cast<float4*>(ast + (long long)c * g.plane) = make_float4(yout[c][0], yout[c][1], yout[c][2], yout[c][3]);
              }
            }
          }
          FDBG(tid == 0, it, 5);
          hand_off(it, has1);
          FDBG(tid == 0, it, 6);
          if (j + 4 < nchunks) prefetch(buf, j + 4);
          it += has1 ? 2u : 1u;
        };
        float4 zA[MAXI][CS], zB[MAXI][CS];
        if (grp < nchunks) prefetch(zA, grp);
        if (grp + 2 < nchunks) prefetch(zB, grp + 2);
        for (int j0 = 0; j0 < nchunks; j0 += 4) {
          step(zA, j0);
          if (j0 + 2 < nchunks) step(zB, j0 + 2);
        }
      }
      // ================= fused layers: epilogue of layer l produces the operand of layer l + 1 =================
      for (int l = 0; l < NLf; ++l, ++lay) {
        const bool produce_next = l + 1 < NLf;
        const uint32_t acc = acc_base + (lay & 1u) * (uint32_t)N;
        f_spin_done_seen(bars, it - 1, seen);  // the layer's last chunk: its accumulator is final
        tc_fence_after();
        const float* bias = g.bias[l];
        float* zout = g.Zout[l];
        float* ast_next = produce_next ? g.Astash[l + 1] : nullptr;
        const int nsteps = (ncb + 1) / 2;
        for (int i = 0; i < nsteps; ++i) {
          const uint32_t it0 = it + 2 * (uint32_t)i;
          const bool has1 = 2 * i + 1 < ncb;
          const uint32_t drow = produce_next ? it0 : 47u;
          FDBG(tid == 0, drow, 0);
          const int cb = 2 * i + grp;
          // this item's bias quad: requested before the stage wait / fill / barrier so that its latency is hidden
          const int col = (cb < ncb ? cb : 0) * 32 + 4 * kq;
          const float4 b4 = bias ? make_float4(__ldg(bias + col), __ldg(bias + col + 1), __ldg(bias + col + 2), __ldg(bias + col + 3))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          // the two stages of this step (operand chunks AND exchange tiles) have been released; the last layer's
          // epilogue only borrows them as scratch (every MMA issued so far has retired, nothing to wait for)
          if (produce_next) wait_stages(it0, has1);
          FDBG(tid == 0, drow, 1);
          // exchange tile of block cb = the A_lo region of the stage its operand chunk goes to
          unsigned char* stage_ptr = base_ptr + ((it + (uint32_t)cb) % FA) * (2 * A_TILE_BYTES);
          unsigned char* Xb = stage_ptr + A_TILE_BYTES;
          if (wg7 < 4 && cb < ncb) {  // warps 0..3 and 7..10 move blocks 2i / 2i+1 out of TMEM
            const int q = warp & 3;   // (7..10) & 3 = 3, 0, 1, 2: all four lane quadrants
            uint32_t v[32];
            tmem_ld32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
            tmem_ld_wait();
            const int row = q * 32 + lane;
#pragma unroll
            for (int t4 = 0; t4 < 8; ++t4)
              *reinterpret_cast<float4*>(Xb + sw128_q(row, t4)) =
                  make_float4(__uint_as_float(v[4 * t4]) * comp, __uint_as_float(v[4 * t4 + 1]) * comp,
                              __uint_as_float(v[4 * t4 + 2]) * comp, __uint_as_float(v[4 * t4 + 3]) * comp);
          }
          FDBG(tid == 0, drow, 2);
          t2_prod_sync();  // exchange tiles complete
          FDBG(tid == 0, drow, 3);
          FDBG(tid == 0, drow, 4);
          if (cb < ncb) {
            for (int pl = pl0; pl < TP; pl += 28) {
              const bool valid = pl < vpts;
              float4 z[CS];
#pragma unroll
              for (int c = 0; c < CS; ++c) z[c] = *reinterpret_cast<const float4*>(Xb + sw128_q(c * TP + pl, kq));
              z[0].x += b4.x; z[0].y += b4.y; z[0].z += b4.z; z[0].w += b4.w;
              const long long goff = (p0 + pl) * g.ld + col;
              if (valid) {
#pragma unroll
                for (int c = 0; c < CS; ++c) *reinterpret_cast<float4*>(zout + (long long)c * g.plane + goff) = z[c];
              }
              if (produce_next) {
                float yout[CS][4];
                act_jets4<L>(act, g.J, z, valid, yout);
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                  store_split4_at(stage_ptr, sw128_q(c * TP + pl, kq), yout[c]);
                  if (ast_next && valid)
                    *reinterpret_cast<float4*>(ast_next + (long long)c * g.plane + goff) =
                        make_float4(yout[c][0], yout[c][1], yout[c][2], yout[c][3]);
                }
              }
            }
          }
          FDBG(tid == 0, drow, 5);
          if (produce_next) hand_off(it0, has1);
          FDBG(tid == 0, drow, 6);
        }
        if (produce_next) {
          it += (uint32_t)ncb;
        } else {
          tc_fence_before();
          // the last layer's epilogue borrowed the stages as scratch without any barrier protocol: nobody may start
          // producing the next tile's operand chunks into them while a slower warp is still reading its exchange cells
          // (found by the CPU emulation as a run-to-run varying Zbar_1: the race window is a few hundred cycles)
          t2_prod_sync();
        }
      }
    }
  }
#undef FDBG
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal / read this CTA
  if (warp == 1) tmem_dealloc2(acc_base, ncols);
}

// =====================================================================================================
// Fused forward, fp16 hi/lo operands (kind::f16) for every fused layer after the first.
//
// Why.  A fused layer has ONE 256-column accumulator (the other 256 columns belong to the layer being drained), so the
// 3xTF32 scheme puts all 12 round-toward-zero accumulations of a 32-wide chunk into the big accumulator: 96 per
// 256-deep contraction, residual error 1.3e-5 on cfg3 — over the 1e-5 bar.  fp16 has the SAME 11-bit significand as
// tf32, so x = hi + lo with fp16 pieces is just as exact a split, but kind::f16 contracts K = 16 per instruction:
// half the accumulations (48), measured error of a 256-deep layer 9.8e-7 of which 8.3e-7 is the compensated bias
// (tests/microbench/tc_numerics.cu, scheme 4) — better than the round-1 TF32 split with two accumulators.  It also
// halves the tensor time per layer (6 instead of 12 kind::tf32-equivalents per 32 k).
//
// Range.  fp16 spans 2^-24 .. 65504.  Every operand ROW (one channel of one point) gets its own power-of-two scale so
// that a rigorous bound of its largest entry lands in [2^13, 2^14): rows are independent in A W, so the scale is
// undone per row in the next epilogue.  The bounds come from the accumulator itself (a pre-pass takes the row maxima
// m_c = max_h |z_c[h]| straight out of TMEM) and the tanh jet formulas:  |y_0| <= 1,  |y_1| <= m_1,
// |y_2| <= m_2 + 0.385 m_1^2  (|s_1| <= 1, |s_2| = |t (1 - t^2)| <= 2 / (3 sqrt 3)).  Entries far below their row
// maximum keep full relative precision down to 2^-24 absolute (fp16 subnormals are honoured by the tensor cores), i.e.
// 2^-38 of the row maximum.  Weights are scaled once per layer (k_w16_scale).  The first fused layer still takes its
// operand from HBM and runs in 3xTF32 (no scale is known before its rows have been seen).
//
// A K chunk is 64 wide (one 128-byte fp16 row): an epilogue step (two 32-column blocks, one per worker group) produces
// exactly ONE chunk, the two groups filling the two halves of every row; stages, weight slots, descriptors and the 12
// MMAs per chunk are identical to the tf32 kernel.
// =====================================================================================================
struct FusedFwd16Args {
  FusedFwdArgs f;               // Wimg[0]: tf32 image of the first fused layer; Wimg[i >= 1]: fp16 hi / lo images
  const float* wscale;          // [n_fused] power-of-two weight scales (entry 0 unused)
};

__host__ __device__ inline int fused_fwd16_smem_bytes(int N) { return fused_fwd_smem_bytes(N) + 1536; }
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // fp16 A / B, fp32 accumulate, K-major
}
// single accumulator, K = 16 per instruction: 12 accumulations per 64-wide chunk at full magnitude
__host__ __device__ __forceinline__ float tc_rz_comp_single16(long long nchunks) { return 1.f + TC_RZ_C0 * 6.f * (float)nchunks; }
// 8-byte cell of (row, group g, k-quad kq) inside a [128 x 64 fp16] K-major SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t cell16(int row, int g, int kq) {
  return (uint32_t)(row * 128 + ((((g << 2) | (kq >> 1)) ^ row) & 7) * 16 + ((kq & 1) << 3));
}
// largest power of two s with bound * s < 2^14  (bound > 0); 1 for a zero row
__device__ __forceinline__ float pow2_scale_for(float bound) {
  if (!(bound > 0.f)) return 1.f;
  int ex = (int)((__float_as_uint(bound) >> 23) & 255u) - 127;  // floor(log2(bound)) for normal numbers
  ex = ex < -100 ? -100 : (ex > 100 ? 100 : ex);
  return __uint_as_float((uint32_t)(127 + 13 - ex) << 23);
}

// weight scale of one layer: 2^(13 - floor(log2(max |W|)))
__global__ void k_w16_absmax(const float* __restrict__ W, long long n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(W[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
__global__ void k_w16_scale(const unsigned* __restrict__ absmax, float* __restrict__ scale) {
  const float m = __uint_as_float(absmax[0]);
  float s = 1.f;
  if (m > 0.f) {
    const int ex = (int)((absmax[0] >> 23) & 255u) - 127;
    s = __uint_as_float((uint32_t)(127 + 13 - ex) << 23);
  }
  scale[0] = s;
}
// fp16 weight image: Wimg16[j][piece][cell(n, kk)] = split(scale * W[(64 j + kk) * N + n]), 64-wide chunks
__global__ void k_tc_prep_w16(const float* __restrict__ W, unsigned short* __restrict__ img, int K, int N, const float* __restrict__ scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K * N) return;
  const int n = (int)(i / K), k = (int)(i % K);
  const float w = W[(long long)k * N + n] * scale[0];
  const uint32_t hb = f32_to_f16_bits(w);
  const uint32_t lb = f32_to_f16_bits(w - f16_bits_to_f32(hb));
  const int j = k / 64, kk = k % 64;
  unsigned short* blk = img + (long long)j * 2 * N * 64;
  const uint32_t off = (uint32_t)(n * 64 + ((((kk >> 3) ^ n) & 7) << 3) + (kk & 7));  // in halfs
  blk[off] = (unsigned short)hb;
  blk[(long long)N * 64 + off] = (unsigned short)lb;
}

template <class L, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_fused_fwd16(FusedFwd16Args ga) {
  static_assert(L::kStatic && L::KM <= 2, "fp16 fused forward: static jet layouts of order <= 2");
  static_assert(ACT == PPSCI_ACT_TANH, "the row-scale bounds are those of tanh");
  const FusedFwdArgs& g = ga.f;
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.H;
  const uint32_t b_off = (uint32_t)(FA * 2 * A_TILE_BYTES);
  const uint32_t bars_off = b_off + (uint32_t)(FB * fused_bslot_bytes(N));
  const uint32_t bars = base + bars_off;
  unsigned* rowmax = reinterpret_cast<unsigned*>(base_ptr + bars_off + 256);   // [128] row maxima (float bits, >= 0)
  float* rs_buf = reinterpret_cast<float*>(base_ptr + bars_off + 256 + 512);    // [2][128] operand row scales by layer parity
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(2 * N);
  constexpr int CS = L::CS;
  constexpr int TP = 128 / CS;
  constexpr int rows_used = CS * TP;
  const uint32_t rank = cluster_ctarank();
  const uint32_t acc_base = fused_setup(base, base_ptr, bars_off, ncols, T2_NPW + 1 + (rank == 0 ? 1 : 0));
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int nch32 = N / KCH;   // tf32 chunks of the first fused layer
  const int nch64 = N / 64;    // fp16 chunks of every later layer
  const int ncb = N / 32;
  const int NLf = g.n_fused;
  const int act = act_id<ACT>(g.act);
  const int n_tile_pairs = (g.num_tiles + 1) / 2;
  const int my_tp = pair < n_tile_pairs ? (n_tile_pairs - 1 - pair) / npairs + 1 : 0;
  const uint32_t chunks_per_tile = (uint32_t)(nch32 + (NLf - 1) * nch64);
  // timeline rows: CTA 0 (leader) rows 0..47, CTA 1 (its peer) rows 48..95; each SM has its own cycle counter, the two are
  // aligned offline on the multicast mma_done events (epi_start)
  const bool dbg0 = g.dbg && blockIdx.x < 2;
  const uint32_t dbg_r0 = blockIdx.x * 48u;
#ifdef PPSCI_B200_TIMELINE
#define FDBG(cond, row, slot) do { if (dbg0 && (cond) && (row) < 48u) g.dbg[(dbg_r0 + (row)) * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define FDBG(cond, row, slot) do { } while (0)
#endif
#ifdef PPSCI_B200_TIMELINE
#define FDBG_MAX(cond, row, slot) do { if (dbg0 && (cond) && (row) < 48u) atomicMax((unsigned long long*)&g.dbg[(dbg_r0 + (row)) * 16 + (slot)], (unsigned long long)clock64()); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define FDBG_MAX(cond, row, slot) do { } while (0)
#endif

  if (warp == T2_TMA_WARP) {
    if (lane == 0) {  // weight streamer (see fused_stream_weights): layer 0 has nch32 chunks, the others nch64; same bytes per chunk
      const int NH = N / 2;
      const uint32_t half_bytes = (uint32_t)(NH * KCH * 4);
      uint32_t it = 0;
      for (int t = 0; t < my_tp; ++t)
        for (int l = 0; l < NLf; ++l) {
          const float* wl = g.Wimg[l] + (long long)rank * NH * KCH;
          const int nch = l == 0 ? nch32 : nch64;
          for (int j = 0; j < nch; ++j, ++it) {
            if (it >= FB) f_wait_done_lane(bars, it - FB);
            const float* img = wl + (long long)j * 2 * N * KCH;
            const uint32_t dst = base + b_off + (it % FB) * (2 * half_bytes);
            const uint32_t fb = bars + 64 + 8 * (it % FA);
            mbar_expect_tx(fb, 2 * half_bytes);
            bulk_g2s(dst, img, half_bytes, fb);
            bulk_g2s(dst + half_bytes, img + (long long)N * KCH, half_bytes, fb);
          }
        }
    }
  } else if (warp == T2_MMA_WARP) {
    if (lane == 0) {
      const int NH = N / 2;
      const uint32_t total_it = (uint32_t)my_tp * chunks_per_tile;
      if (rank == 0) {
        const uint32_t idesc32 = make_idesc_tf32(256, N), idesc16 = make_idesc_f16(256, N);
        const uint64_t d_a_hi = make_smem_desc(base), d_a_lo = make_smem_desc(base + A_TILE_BYTES);
        const uint64_t d_b_hi = make_smem_desc(base + b_off);
        const uint64_t d_b_lo = make_smem_desc(base + b_off + (uint32_t)(NH * KCH * 4));
        const uint64_t a_inc = (uint64_t)((2 * A_TILE_BYTES) >> 4), b_inc = (uint64_t)(fused_bslot_bytes(N) >> 4);
        uint32_t lay = 0, it = 0;
        for (int t = 0; t < my_tp; ++t)
          for (int l = 0; l < NLf; ++l, ++lay) {
            const uint32_t acc = acc_base + (lay & 1u) * (uint32_t)N;
            const int nch = l == 0 ? nch32 : nch64;
            for (int j = 0; j < nch; ++j, ++it) {
              FDBG(true, it, 8);
              mbar_wait_cluster(bars + 64 + 8 * (it % FA), (it / FA) & 1u);
              FDBG(true, it, 9);
              tc_fence_after();
              const uint64_t ao = (uint64_t)(it % FA) * a_inc, bo = (uint64_t)(it % FB) * b_inc;
              if (l == 0) {
                issue_chunk_mmas_2_single(acc, d_a_hi + ao, d_a_lo + ao, d_b_hi + bo, d_b_lo + bo, idesc32, j == 0);
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {  // K = 16 fp16 = 32 bytes per instruction: same descriptor stepping
                  const uint64_t inc = (uint64_t)(2 * ks);
                  mma_f16_2(acc, d_a_hi + ao + inc, d_b_hi + bo + inc, idesc16, (j == 0 && ks == 0) ? 0u : 1u);
                  mma_f16_2(acc, d_a_lo + ao + inc, d_b_hi + bo + inc, idesc16, 1u);
                  mma_f16_2(acc, d_a_hi + ao + inc, d_b_lo + bo + inc, idesc16, 1u);
                }
              }
              mma_commit_2(bars + 8 * (it % FA));
              FDBG(true, it, 10);
            }
          }
      } else {
        for (uint32_t it = 0; it < total_it; ++it) {
          FDBG(true, it, 8);
          mbar_wait(bars + 64 + 8 * (it % FA), (it / FA) & 1u);
          FDBG(true, it, 9);
          mbar_remote_arrive(bars + 64 + 8 * (it % FA), 0);
        }
      }
    }
  } else {
    const int grp = warp / 7, wg7 = warp - grp * 7;
    const int kq = lane & 7, psub = lane >> 3;
    // Point of this thread's item.  The fp16 operand cells are 8 bytes, so one shared-memory wavefront serves a HALF
    // warp = two points: with consecutive points their rows (c TP + pl) mostly share bit 2 of (row & 7), the 128-byte
    // swizzle then maps both onto the same 16 banks and every LDS.64 / STS.64 of the item loop took ~1.9x its ideal
    // wavefronts (ncu source page, round 2).  For TP <= 25 the two points of a half warp are 4 apart instead:
    // warp pairs cover 8 points as {0,4,1,5} | {2,6,3,7}, the seventh warp takes point 24.
    const int pl0 = (TP <= 25) ? (8 * (wg7 >> 1) + 2 * (wg7 & 1) + (psub >> 1) + 4 * (psub & 1)) : wg7 * 4 + psub;
    constexpr int MAXI = (TP + 27) / 28;
    uint32_t it = 0, lay = 0, seen = 0;
    // exchange / operand cells of this thread's first-pass item (constant over steps, layers and tiles)
    uint32_t coff0[CS];
#pragma unroll
    for (int c = 0; c < CS; ++c) coff0[c] = cell16(c * TP + (pl0 < TP ? pl0 : 0), grp, kq);
    auto hand_off = [&](uint32_t it0, bool has1) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bars + 64 + 8 * (it0 % FA));
        if (has1) mbar_arrive(bars + 64 + 8 * ((it0 + 1) % FA));
      }
    };
    for (int tp = pair; tp < n_tile_pairs; tp += npairs) {
      const long long tile = 2LL * tp + rank;
      const long long p0 = tile * TP;
      const long long vleft = g.Np - p0;
      const int vpts = vleft >= TP ? TP : (vleft > 0 ? (int)vleft : 0);
      // ================= P0 (3xTF32): operand of the first fused layer = act_jets(Zin), unscaled =================
      {
        auto prefetch = [&](float4 (&buf)[MAXI][CS], int j) {
          const int col = j * KCH + 4 * kq;
#pragma unroll
          for (int i = 0; i < MAXI; ++i) {
            const int pl = pl0 + i * 28;
            const bool ok = pl < vpts;
            const float* src = g.Zin + (p0 + pl) * g.ld + col;
#pragma unroll
            for (int c = 0; c < CS; ++c)
              buf[i][c] = ok ? __ldg(reinterpret_cast<const float4*>(src + (long long)c * g.plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        auto step = [&](float4 (&buf)[MAXI][CS], int j0) {
          const bool has1 = j0 + 1 < nch32;
          const int j = j0 + grp;
          {
            const uint32_t last = has1 ? it + 1 : it;
            if (last >= FA) f_wait_done_seen(bars, last - FA, seen);
          }
          if (j < nch32) {
            unsigned char* stage_ptr = base_ptr + ((it + (uint32_t)grp) % FA) * (2 * A_TILE_BYTES);
#pragma unroll
            for (int i = 0; i < MAXI; ++i) {
              const int pl = pl0 + i * 28;
              if (pl >= TP) continue;
              const bool valid = pl < vpts;
              float yout[CS][4];
              act_jets4<L>(act, g.J, buf[i], valid, yout);
              float* ast = (g.Astash[0] && valid) ? g.Astash[0] + (p0 + pl) * g.ld + j * KCH + 4 * kq : nullptr;
#pragma unroll
              for (int c = 0; c < CS; ++c) {
                store_split4_at(stage_ptr, sw128_q(c * TP + pl, kq), yout[c]);
                if (ast) *reinterpret_cast<float4*>(ast + (long long)c * g.plane) = make_float4(yout[c][0], yout[c][1], yout[c][2], yout[c][3]);
              }
            }
          }
          hand_off(it, has1);
          if (j + 4 < nch32) prefetch(buf, j + 4);
          it += has1 ? 2u : 1u;
        };
        float4 zA[MAXI][CS], zB[MAXI][CS];
        if (grp < nch32) prefetch(zA, grp);
        if (grp + 2 < nch32) prefetch(zB, grp + 2);
        if (tid < 128) rs_buf[tid] = 1.f;  // the first fused layer's operand rows are unscaled (ordered by the syncs of E(0))
        for (int j0 = 0; j0 < nch32; j0 += 4) {
          step(zA, j0);
          if (j0 + 2 < nch32) step(zB, j0 + 2);
        }
      }
      // ================= fused layers =================
      for (int l = 0; l < NLf; ++l, ++lay) {
        const bool produce_next = l + 1 < NLf;
        const uint32_t acc = acc_base + (lay & 1u) * (uint32_t)N;
        f_spin_done_seen(bars, it - 1, seen);  // the layer's last chunk: its accumulator is final
        tc_fence_after();
        const float* bias = g.bias[l];
        float* zout = g.Zout[l];
        float* ast_next = produce_next ? g.Astash[l + 1] : nullptr;
        const float* rs_cur = rs_buf + (l & 1) * 128;
        float* rs_next = rs_buf + ((l + 1) & 1) * 128;
        // accumulator -> true pre-activation: rz compensation, weight scale, operand row scale
        const float unw = l == 0 ? tc_rz_comp_single(nch32) : tc_rz_comp_single16(nch64) / __ldg(ga.wscale + l);
        const int nsteps = ncb / 2;
        FDBG(tid == 0, produce_next ? it : 46u, 11);
        if (tid < 128) rowmax[tid] = 0u;
        t2_prod_sync();  // rs_cur (written by other warps in P0 / the previous epilogue) and the cleared maxima are visible
        if (produce_next) {
          // ---- pre-pass: row maxima of |z| straight out of TMEM -> rigorous bounds of the next operand rows -> scales ----
          if (wg7 < 4) {  // warps 0..3 / 7..10: quadrant warp & 3, column half grp
            const int q = warp & 3, row = q * 32 + lane;
            float m = 0.f;
            for (int cb = grp * (ncb / 2); cb < (grp + 1) * (ncb / 2); ++cb) {
              uint32_t v[32];
              tmem_ld32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
              tmem_ld_wait();
#pragma unroll
              for (int t = 0; t < 32; ++t) m = fmaxf(m, fabsf(__uint_as_float(v[t])));
            }
            atomicMax(&rowmax[row], __float_as_uint(m * unw / rs_cur[row]));
          }
          t2_prod_sync();
          if (tid < 128) {
            float s = 1.f;
            if (tid < rows_used) {
              const int c = tid / TP, pl = tid - c * TP;
              float bound = 1.f;  // value channel: |tanh| <= 1
#pragma unroll
              for (int d = 0; d < L::ND; ++d) {
                const int K = L::order(g.J, d), cbs = L::cbase(g.J, d);
                const float m1 = __uint_as_float(rowmax[cbs * TP + pl]);
                if (c == cbs) bound = m1;
                if (K >= 2 && c == cbs + 1) bound = __uint_as_float(rowmax[(cbs + 1) * TP + pl]) + 0.385f * m1 * m1;
              }
              s = pow2_scale_for(bound);
            }
            rs_next[tid] = s;
          }
          t2_prod_sync();
        }
        for (int i = 0; i < nsteps; ++i) {
          const uint32_t itc = it + (uint32_t)i;  // the ONE 64-wide chunk this step produces (both groups, one row half each)
          const uint32_t drow = produce_next ? itc : 47u;
          FDBG(tid == 0, drow, 0);
          const int cb = 2 * i + grp;
          // this item's bias quad: requested before the stage wait / fill / barrier so that its latency is hidden
          const int col = cb * 32 + 4 * kq;
          const float4 b4 = bias ? make_float4(__ldg(bias + col), __ldg(bias + col + 1), __ldg(bias + col + 2), __ldg(bias + col + 3))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          if (produce_next && itc >= FA) f_wait_done_seen(bars, itc - FA, seen);
          FDBG(tid == 0, drow, 1);
          unsigned char* stage_ptr = base_ptr + (itc % FA) * (2 * A_TILE_BYTES);
          if (wg7 < 4) {  // raw block -> the 8-byte cells its items will overwrite: (v0, v1) -> hi cell, (v2, v3) -> lo cell
            const int q = warp & 3, row = q * 32 + lane;
            uint32_t v[32];
            tmem_ld32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
            tmem_ld_wait();
            const float f = unw / rs_cur[row];
#pragma unroll
            for (int t4 = 0; t4 < 8; ++t4) {
              const uint32_t off = cell16(row, grp, t4);
              *reinterpret_cast<float2*>(stage_ptr + off) = make_float2(__uint_as_float(v[4 * t4]) * f, __uint_as_float(v[4 * t4 + 1]) * f);
              *reinterpret_cast<float2*>(stage_ptr + A_TILE_BYTES + off) =
                  make_float2(__uint_as_float(v[4 * t4 + 2]) * f, __uint_as_float(v[4 * t4 + 3]) * f);
            }
          }
          FDBG(tid == 0, drow, 2);
          t2_prod_sync();  // exchange cells complete
          FDBG(tid == 0, drow, 3);
          {
            for (int pl = pl0; pl < TP; pl += 28) {
              const bool valid = pl < vpts;
              float4 z[CS];
#pragma unroll
              for (int c = 0; c < CS; ++c) {
                const uint32_t off = (TP <= 28) ? coff0[c] : cell16(c * TP + pl, grp, kq);  // one item pass: hoisted offsets
                const float2 a = *reinterpret_cast<const float2*>(stage_ptr + off);
                const float2 b = *reinterpret_cast<const float2*>(stage_ptr + A_TILE_BYTES + off);
                z[c] = make_float4(a.x, a.y, b.x, b.y);
              }
              z[0].x += b4.x; z[0].y += b4.y; z[0].z += b4.z; z[0].w += b4.w;
              const long long goff = (p0 + pl) * g.ld + col;
              if (valid) {
#pragma unroll
                for (int c = 0; c < CS; ++c) *reinterpret_cast<float4*>(zout + (long long)c * g.plane + goff) = z[c];
              }
              if (produce_next) {
                float yout[CS][4];
                act_jets4<L>(act, g.J, z, true, yout);
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                  // rows of points beyond the batch are written as zeros: one select on the row scale instead of 20
                  const float sc = valid ? rs_next[c * TP + pl] : 0.f;
                  const float y0 = yout[c][0] * sc, y1 = yout[c][1] * sc, y2 = yout[c][2] * sc, y3 = yout[c][3] * sc;
                  const uint32_t h01 = f32x2_to_f16x2_bits(y0, y1), h23 = f32x2_to_f16x2_bits(y2, y3);  // packed conversions
                  float b0, b1, b2, b3;
                  f16x2_bits_to_f32x2(h01, b0, b1);
                  f16x2_bits_to_f32x2(h23, b2, b3);
                  const uint32_t l01 = f32x2_to_f16x2_bits(y0 - b0, y1 - b1), l23 = f32x2_to_f16x2_bits(y2 - b2, y3 - b3);
                  const uint32_t off = (TP <= 28) ? coff0[c] : cell16(c * TP + pl, grp, kq);
                  *reinterpret_cast<uint2*>(stage_ptr + off) = make_uint2(h01, h23);
                  *reinterpret_cast<uint2*>(stage_ptr + A_TILE_BYTES + off) = make_uint2(l01, l23);
                  if (ast_next && valid)
                    *reinterpret_cast<float4*>(ast_next + (long long)c * g.plane + goff) =
                        make_float4(yout[c][0], yout[c][1], yout[c][2], yout[c][3]);
                }
              }
            }
          }
          FDBG(tid == 0, drow, 5);
          FDBG_MAX(lane == 0, drow, 12);  // the slowest warp's items_done
          if (produce_next) hand_off(itc, false);
          else t2_prod_sync();  // scratch stage reuse two steps later is ordered by the next step's sync; keep groups together
          FDBG(tid == 0, drow, 6);
          FDBG_MAX(lane == 0, drow, 7);   // the slowest warp's arrival
        }
        if (produce_next) it += (uint32_t)nsteps;
        else tc_fence_before();
      }
    }
  }
#undef FDBG
#undef FDBG_MAX
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(acc_base, ncols);
}

// =====================================================================================================
// Fused backward dx chain over the hidden -> hidden layers of one point chunk (the "dX" half of
// total_loss.backward(), ppsci/solver/train.py:158, for layers L-1 .. 2 in ONE launch):
//   P0: operand of the first dx GEMM = hi/lo split of Zbar_{L-1} rows from HBM;
//   fused step i (layer l = L-1-i):  Abar_{l-1} = Zbar_l W_l^T  on the tensor cores (transposed weight image), then
//   the epilogue applies the activation adjoint  Zbar_{l-1} = adj(Abar_{l-1}, Z_{l-1})  with Z_{l-1} read from the
//   forward stash, writes Zbar_{l-1} to HBM ONCE (the dW kernels' operand) and splits it straight into the A operand
//   chunks of the next dx GEMM — Zbar never comes back from HBM inside the chain.
// Same roles, ring and barrier protocol as k_fused_fwd.
// =====================================================================================================
struct FusedDxArgs {
  JetLayout J;
  int act;
  int H;
  int n_fused;                    // fused dx layers: i = 0 .. n_fused-1  <->  network layer l = L-1-i
  const float* ZbarIn;            // Zbar_{L-1}: [C][nc_max][ld]
  int ld;
  long long plane;
  const float* WimgT[FUSE_MAXL];  // transposed weight image of layer l (k_tc_prep_w, transposed = 1)
  const float* Zprev[FUSE_MAXL];  // Z_{l-1}: pre-activations below layer l (forward stash)
  float* ZbarOut[FUSE_MAXL];      // Zbar_{l-1}
  long long Np;
  int num_tiles;
  long long* dbg;
};


template <class L, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_fused_dx(FusedDxArgs g) {
  static_assert(L::kStatic, "fused kernels are specialised for the static jet layouts");
  PPSCI_DYN_SMEM(smem_dyn);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
  const int N = g.H;
  const uint32_t b_off = (uint32_t)(FA * 2 * A_TILE_BYTES);
  const uint32_t bars_off = b_off + (uint32_t)(FB * fused_bslot_bytes(N));
  const uint32_t bars = base + bars_off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = tc_pow2_cols(2 * N);
  constexpr int CS = L::CS;
  constexpr int TP = 128 / CS;
  constexpr int rows_used = CS * TP;
  const uint32_t rank = cluster_ctarank();
  const uint32_t acc_base = fused_setup(base, base_ptr, bars_off, ncols, T2_NPW + 1 + (rank == 0 ? 1 : 0));
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int nchunks = N / KCH;
  const int ncb = N / 32;
  const int NLf = g.n_fused;
  const int act = act_id<ACT>(g.act);
  const int n_tile_pairs = (g.num_tiles + 1) / 2;
  const int my_tp = pair < n_tile_pairs ? (n_tile_pairs - 1 - pair) / npairs + 1 : 0;
  const bool dbg0 = g.dbg && blockIdx.x == 0;
#ifdef PPSCI_B200_TIMELINE
#define FDBG(cond, row, slot) do { if (dbg0 && (cond) && (row) < 48u) g.dbg[(row) * 16 + (slot)] = clock64(); } while (0)
#else  // product build: no stamp code at all (even predicated off it costs issue slots in the item loops)
#define FDBG(cond, row, slot) do { } while (0)
#endif

  if (warp == T2_TMA_WARP) {
    if (lane == 0) fused_stream_weights(base, bars, b_off, g.WimgT, my_tp, NLf, nchunks, N, rank);
  } else if (warp == T2_MMA_WARP) {
    if (lane == 0)
      fused_issue_mmas(base, bars, b_off, acc_base, my_tp, NLf, nchunks, N, rank, [&](uint32_t row, int slot) { FDBG(true, row, slot); });
  } else {
    // ---- workers (see k_fused_fwd) ----
    const int grp = warp / 7, wg7 = warp - grp * 7;
    const int kq = lane & 7, psub = lane >> 3;
    const int pl0 = wg7 * 4 + psub;
    constexpr int MAXR = (128 + 27) / 28;  // P0 item = (row, 4 consecutive k), 28 rows per pass and group
    const float comp = tc_rz_comp_single(nchunks);
    // swizzled 16-byte cells of this thread's first-pass item (constant over steps, layers and tiles)
    uint32_t coff0[CS];
#pragma unroll
    for (int c = 0; c < CS; ++c) coff0[c] = sw128_q(c * TP + (pl0 < TP ? pl0 : 0), kq);
    uint32_t it = 0, lay = 0, seen = 0;
    auto hand_off = [&](uint32_t it0, bool has1) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bars + 64 + 8 * (it0 % FA));
        if (has1) mbar_arrive(bars + 64 + 8 * ((it0 + 1) % FA));
      }
    };
    // both stages of a step are free once the LATER of the two chunks that used them has retired (tcgen05.commit
    // completes in issue order): one barrier poll per step instead of two, and none when the watermark covers it
    auto wait_stages = [&](uint32_t it0, bool has1) {
      const uint32_t last = has1 ? it0 + 1 : it0;
      if (last >= FA) f_wait_done_seen(bars, last - FA, seen);
    };
    for (int tp = pair; tp < n_tile_pairs; tp += npairs) {
      const long long tile = 2LL * tp + rank;
      const long long p0 = tile * TP;
      const long long vleft = g.Np - p0;
      const int vpts = vleft >= TP ? TP : (vleft > 0 ? (int)vleft : 0);
      // ================= P0: hi/lo split of the Zbar_{L-1} rows =================
      {
        auto prefetch = [&](float4 (&buf)[MAXR], int j) {
          const int col = j * KCH + 4 * kq;
#pragma unroll
          for (int i = 0; i < MAXR; ++i) {
            const int r = pl0 + i * 28;
            const int c = r / TP, pl = r - c * TP;
            const bool ok = r < rows_used && pl < vpts;
            buf[i] = ok ? __ldg(reinterpret_cast<const float4*>(g.ZbarIn + (long long)c * g.plane + (p0 + pl) * g.ld + col))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        auto step = [&](float4 (&buf)[MAXR], int j0) {
          const bool has1 = j0 + 1 < nchunks;
          const int j = j0 + grp;
          wait_stages(it, has1);
          if (j < nchunks) {
            unsigned char* stage_ptr = base_ptr + ((it + (uint32_t)grp) % FA) * (2 * A_TILE_BYTES);
#pragma unroll
            for (int i = 0; i < MAXR; ++i) {
              const int r = pl0 + i * 28;
              if (r < rows_used) {
                const float v[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
                store_split4_at(stage_ptr, sw128_q(r, kq), v);
              }
            }
          }
          hand_off(it, has1);
          if (j + 4 < nchunks) prefetch(buf, j + 4);
          it += has1 ? 2u : 1u;
        };
        float4 zA[MAXR], zB[MAXR];
        if (grp < nchunks) prefetch(zA, grp);
        if (grp + 2 < nchunks) prefetch(zB, grp + 2);
        for (int j0 = 0; j0 < nchunks; j0 += 4) {
          step(zA, j0);
          if (j0 + 2 < nchunks) step(zB, j0 + 2);
        }
      }
      // ================= fused dx layers =================
      for (int l = 0; l < NLf; ++l, ++lay) {
        const bool produce_next = l + 1 < NLf;
        const uint32_t acc = acc_base + (lay & 1u) * (uint32_t)N;
        const float* zprev = g.Zprev[l];
        float* zbout = g.ZbarOut[l];
        const int nsteps = (ncb + 1) / 2;
        auto load_z = [&](float4 (&zc)[CS], int pl, int cb) {
          const float* src = zprev + (p0 + pl) * g.ld + cb * 32 + 4 * kq;
#pragma unroll
          for (int c = 0; c < CS; ++c) zc[c] = __ldg(reinterpret_cast<const float4*>(src + (long long)c * g.plane));
        };
        // the Z_{l-1} block of the first step is requested before the accumulator is even final
        float4 zc0[CS];
#pragma unroll
        for (int c = 0; c < CS; ++c) zc0[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grp < ncb && pl0 < vpts) load_z(zc0, pl0, grp);
        f_spin_done_seen(bars, it - 1, seen);  // the layer's last chunk: Abar_{l-1} is final
        tc_fence_after();
        FDBG(tid == 0, produce_next ? it : 46u, 11);
        for (int i = 0; i < nsteps; ++i) {
          const uint32_t it0 = it + 2 * (uint32_t)i;
          const bool has1 = 2 * i + 1 < ncb;
          const uint32_t drow = produce_next ? it0 : 47u;
          FDBG(tid == 0, drow, 0);
          // the two stages of this step (operand chunks AND exchange tiles) have been released; the last layer's
          // epilogue only borrows them as scratch (every MMA issued so far has retired, nothing to wait for)
          if (produce_next) wait_stages(it0, has1);
          FDBG(tid == 0, drow, 1);
          const int cb = 2 * i + grp;
          unsigned char* stage_ptr = base_ptr + ((it + (uint32_t)cb) % FA) * (2 * A_TILE_BYTES);
          unsigned char* Xb = stage_ptr + A_TILE_BYTES;  // exchange tile = the A_lo region of the destination stage
          if (cb + 2 < ncb && pl0 < vpts) {  // pull the next step's Z block towards L2 while this step runs
#pragma unroll
            for (int c = 0; c < CS; ++c) prefetch_l2(zprev + (long long)c * g.plane + (p0 + pl0) * g.ld + (cb + 2) * 32 + 4 * kq);
          }
          if (wg7 < 4 && cb < ncb) {
            const int q = warp & 3;
            uint32_t v[32];
            tmem_ld32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
            tmem_ld_wait();
            const int row = q * 32 + lane;
#pragma unroll
            for (int t4 = 0; t4 < 8; ++t4)
              *reinterpret_cast<float4*>(Xb + sw128_q(row, t4)) =
                  make_float4(__uint_as_float(v[4 * t4]) * comp, __uint_as_float(v[4 * t4 + 1]) * comp,
                              __uint_as_float(v[4 * t4 + 2]) * comp, __uint_as_float(v[4 * t4 + 3]) * comp);
          }
          FDBG(tid == 0, drow, 2);
          t2_prod_sync();  // exchange tiles complete
          FDBG(tid == 0, drow, 3);
          if (cb < ncb) {
            for (int pl = pl0; pl < TP; pl += 28) {
              const bool valid = pl < vpts;
              float4 zc[CS], xc[CS];
              if (pl == pl0) {
#pragma unroll
                for (int c = 0; c < CS; ++c) zc[c] = zc0[c];
              } else if (valid) {
                load_z(zc, pl, cb);
              } else {
#pragma unroll
                for (int c = 0; c < CS; ++c) zc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int c = 0; c < CS; ++c)
                xc[c] = *reinterpret_cast<const float4*>(Xb + ((TP <= 28) ? coff0[c] : sw128_q(c * TP + pl, kq)));
              float ob[CS][4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float sc_[6];
                float y0;
                act_coef<float, L::KM + 1>(act, f4comp(zc[0], t), y0, sc_);
                float sb_[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int d = 0; d < L::ND; ++d) {
                  const int K = L::order(g.J, d), cbs = L::cbase(g.J, d);
                  float zz[4], yb[4], zbv[4];
#pragma unroll
                  for (int o = 0; o < 4; ++o) {
                    const bool on = (o < L::KM && o < K && cbs + o < CS);
                    zz[o] = on ? f4comp(zc[on ? cbs + o : 0], t) : 0.f;
                    yb[o] = on ? f4comp(xc[on ? cbs + o : 0], t) : 0.f;
                    zbv[o] = 0.f;
                  }
                  jet_adj_dir<float, L::KM>(sc_, zz, yb, zbv, sb_);
#pragma unroll
                  for (int o = 0; o < L::KM; ++o)
                    if (o < K && cbs + o < CS) ob[cbs + o][t] = zbv[o];
                }
                // points beyond the batch: their operand rows were zero, so Abar = 0 exactly and (with z = 0) every
                // adjoint below is zero without a select
                ob[0][t] = jet_adj_z0<float, L::KM>(sc_, f4comp(xc[0], t), sb_);
              }
              if (valid) {
                float* out = zbout + (p0 + pl) * g.ld + cb * 32 + 4 * kq;
#pragma unroll
                for (int c = 0; c < CS; ++c)
                  *reinterpret_cast<float4*>(out + (long long)c * g.plane) = make_float4(ob[c][0], ob[c][1], ob[c][2], ob[c][3]);
              }
              if (produce_next) {
#pragma unroll
                for (int c = 0; c < CS; ++c) store_split4_at(stage_ptr, (TP <= 28) ? coff0[c] : sw128_q(c * TP + pl, kq), ob[c]);
              }
            }
          }
          FDBG(tid == 0, drow, 5);
          if (produce_next) hand_off(it0, has1);
          FDBG(tid == 0, drow, 6);
          // this thread's Z block of the next step (its block index advances by 2); requested AFTER the hand-off:
          // fence.proxy.async also waits for the thread's outstanding global loads
          if (cb + 2 < ncb && pl0 < vpts) load_z(zc0, pl0, cb + 2);
        }
        if (produce_next) {
          it += (uint32_t)ncb;
        } else {
          tc_fence_before();
          // the last layer's epilogue borrowed the stages as scratch without any barrier protocol: nobody may start
          // producing the next tile's operand chunks into them while a slower warp is still reading its exchange cells
          // (found by the CPU emulation as a run-to-run varying Zbar_1: the race window is a few hundred cycles)
          t2_prod_sync();
        }
      }
    }
  }
#undef FDBG
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(acc_base, ncols);
}

}  // namespace tc
}  // namespace ppsci

// static jet layouts only (callers check lay != TC_LAY_DYN); tanh activation
#define PPSCI_FUSED_LAUNCH(KERNEL, lay, grid, smem, stream, args, err_expr)                                         \
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
