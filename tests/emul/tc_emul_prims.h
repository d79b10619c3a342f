// tc_emul_prims.h — TEST INFRASTRUCTURE ONLY (included by kernels_tc.cuh / kernels_tc2.cuh under -DPPSCI_EMUL).
// The PTX wrappers of the tensor-core kernels, same names and signatures, mapped onto the CPU emulation in
// cuda_emul.h.  See that file for what is modelled (descriptor decoding, swizzle, RZ accumulation, deferred MMA
// execution, mbarrier phases, lane-quadrant checks) and what is not (timing, proxy fences).
#pragma once
#ifndef PPSCI_EMUL
#error "tc_emul_prims.h is only for -DPPSCI_EMUL builds"
#endif
#include "cuda_emul.h"

namespace ppsci {
namespace tc {

inline uint32_t smem_u32(const void* p) { return emul::smem_addr_of(p); }
inline float tf32_rn(float x) { return emul::tf32_rna(x); }

// ---- mbarrier ----
inline void mbar_init(uint32_t bar, uint32_t count) { emul::mbar_init(bar, count); }
inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) { emul::mbar_arrive_on(emul::t_cta, bar, bytes); }
inline void mbar_arrive(uint32_t bar) { emul::mbar_arrive_on(emul::t_cta, bar, 0); }
inline bool mbar_test_wait(uint32_t bar, uint32_t parity) { return emul::mbar_test(bar, parity); }
inline bool mbar_try_wait(uint32_t bar, uint32_t parity) { return emul::mbar_test(bar, parity); }
inline void mbar_wait(uint32_t bar, uint32_t parity) { emul::mbar_wait_block(bar, parity); }
inline void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  if ((emul::linear_tid() & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}

// ---- copies ----
inline void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) { emul::bulk_g2s(dst_smem, src, bytes, bar); }
inline void cp_async16(uint32_t dst_smem, const void* src, bool valid) { emul::cp_async16(dst_smem, src, valid); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}

// ---- TMEM ----
inline void tmem_alloc(uint32_t dst_smem, uint32_t ncols) { emul::tmem_alloc(dst_smem, ncols); __syncwarp(); }
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) { emul::tmem_ld32(taddr, r); }
inline void tmem_ld_wait() {}

// ---- UMMA ----
inline void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { emul::mma_issue(1, 0, d, a, b, idesc, acc); }
inline void mma_commit(uint32_t bar) { emul::mma_commit(1, bar, 1u); }

// ---- cluster / pair (kernels_tc2.cuh) ----
inline uint32_t cluster_ctarank() { return emul::t_cta->rank; }
inline void cluster_sync_all() { emul::t_cta->cluster->sync.wait(); }
inline void mbar_remote_arrive(uint32_t local_bar, uint32_t rank) { emul::mbar_arrive_on(&emul::t_cta->cluster->ctas[rank], local_bar, 0); }
inline bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) { return emul::mbar_test(bar, parity); }
inline void mbar_wait_cluster(uint32_t bar, uint32_t parity) { emul::mbar_wait_block(bar, parity); }
inline void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) { emul::tmem_alloc(dst_smem, ncols); __syncwarp(); }
inline void tmem_relinquish2() {}
inline void tmem_dealloc2(uint32_t, uint32_t) {}
inline void mma_tf32_2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { emul::mma_issue(2, 0, d, a, b, idesc, acc); }
inline void mma_f16_2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  emul::mma_issue(2, ((idesc >> 7) & 7u) == 1u ? 2 : 1, d, a, b, idesc, acc);
}
inline void mma_commit_2(uint32_t bar) { emul::mma_commit(2, bar, 3u); }
inline uint32_t f32_to_f16_bits(float x) {
  const _Float16 h = (_Float16)x;  // round to nearest even
  uint16_t u;
  memcpy(&u, &h, 2);
  return (uint32_t)u;
}
inline float f16_bits_to_f32(uint32_t b) {
  const uint16_t u = (uint16_t)b;
  _Float16 h;
  memcpy(&h, &u, 2);
  return (float)h;
}
inline uint32_t f32x2_to_f16x2_bits(float a, float b) { return f32_to_f16_bits(a) | (f32_to_f16_bits(b) << 16); }
inline void f16x2_bits_to_f32x2(uint32_t h, float& a, float& b) {
  a = f16_bits_to_f32(h & 0xFFFFu);
  b = f16_bits_to_f32(h >> 16);
}
// bar.sync id, nthreads
inline void named_bar_sync(int id, int nthreads) { emul::t_cta->named[id].wait_n((unsigned)nthreads); }

}  // namespace tc
}  // namespace ppsci
