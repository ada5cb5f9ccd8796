// xcorr_tc.cu - tensor-core (tcgen05) correlator for 8-bit IQ.  Placeholder until the kernel lands:
// plans report tc_ready=false and LCS_KERNEL_AUTO resolves to the FP32 kernel.
#include "lcs_ctx.hpp"

namespace lcs {
lcs_status tc_plan_setup(lcs_xcorr_plan* p) {
  p->tc_ready = false;
  return LCS_OK;
}
int launch_xcorr_fold_tc(lcs_xcorr_plan*, const void*, uint32_t, float*, cudaStream_t) { return 0; }
}  // namespace lcs
