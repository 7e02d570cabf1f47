// tcgen05 scoring engine (placeholder until the UMMA pipeline lands in this file).
#include "internal.h"
namespace bx {
bool score_tc_supported(const bx_factor_t& f) { (void)f; return false; }
int32_t launch_score_tc(cudaStream_t, const ScoreArgs&, const SamplerPack&) {
  set_error("tcgen05 engine not built");
  return BX_E_UNSUPPORTED;
}
}  // namespace bx
extern "C" int32_t bx_selftest_umma(const float*, const float*, float*, int32_t, int32_t, void*) {
  bx::set_error("tcgen05 engine not built");
  return BX_E_UNSUPPORTED;
}
