// dcn_umma_bwd.cu — tcgen05 backward (placeholder: not yet covering any shape).
#include "dcn_umma.h"
#include "dcn_umma_common.cuh"

namespace dcn {
bool umma_bwd_supported(const Geo&, int) { return false; }
size_t umma_bwd_workspace(const Geo&) { return 0; }
int umma_backward_fp32(const Geo&, int, const float*, const float*, const float*, const float*, float*,
                       float*, float*, float*, void*, cudaStream_t) {
  set_error("tcgen05 backward not built");
  return DCN_ERR_UNSUPPORTED;
}
}  // namespace dcn
