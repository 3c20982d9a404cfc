// placeholder until the tcgen05 family lands: nothing is supported, so dcn_api.cu routes
// every shape to the generic kernels.
#include "dcn_umma.h"
namespace dcn {
bool umma_supported(const Geo&, int, int) { return false; }
size_t umma_workspace_bytes(const Geo&, int, int) { return 0; }
int umma_forward(const Geo&, int, int, const void*, const float*, const void*, const float*, void*,
                 void*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return DCN_ERR_UNSUPPORTED;
}
int umma_backward(const Geo&, int, int, const void*, const float*, const void*, const void*, float*,
                  float*, float*, float*, void*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return DCN_ERR_UNSUPPORTED;
}
}  // namespace dcn
