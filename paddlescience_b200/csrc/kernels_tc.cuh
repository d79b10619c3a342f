// kernels_tc.cuh — tcgen05 / TMEM / TMA kernels for the wide fp32 tanh layers (sm_100a only).
// (placeholder until the tensor-core forward lands; the SIMT kernels serve every plan)
#pragma once
#include <string>
#include "kernels_simt.cuh"

namespace ppsci {
inline bool tc_plan_supported(const ppsci_plan_spec&, int, int) { return false; }
inline size_t tc_scratch_bytes_impl(const ppsci_plan_spec&, int, int64_t) { return 0; }
template <int KMAX>
inline int tc_forward(const ppsci_plan_spec&, const JetLayout&, const int64_t*, const int64_t*, const int*,
                      const float*, const void* const*, int64_t, int64_t, int64_t, unsigned char*, const size_t*,
                      size_t, size_t, cudaStream_t, int64_t*, std::string* err) {
  *err = "tcgen05 backend not built";
  return 1;
}
}  // namespace ppsci
