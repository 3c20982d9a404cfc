// dcn_umma_host.cu — dispatch of the tcgen05 kernel family behind dcn_umma.h.
#include "dcn_umma.h"
#include "dcn_umma_common.cuh"

namespace dcn {

bool umma_fwd_supported(const Geo& g, int operand);
size_t umma_fwd_workspace(const Geo& g, int operand);
int umma_forward_any(const Geo& g, int operand, const void* x, const float* off, const void* wt,
                     const float* bias, float* out, void* workspace, cudaStream_t st);
bool umma_bwd_supported(const Geo& g, int operand);
size_t umma_bwd_workspace(const Geo& g, int operand);
int umma_backward_any(const Geo& g, int operand, int flags, const void* x, const float* off, const void* wt,
                      const void* gout, float* gx, float* goff, float* gw, float* gb, void* workspace,
                      cudaStream_t st);

bool umma_supported(const Geo& g, int operand, int phase) {
  if (phase == DCN_PHASE_FORWARD) return umma_fwd_supported(g, operand);
  if (phase == DCN_PHASE_BACKWARD) return umma_bwd_supported(g, operand);
  return false;
}

size_t umma_workspace_bytes(const Geo& g, int operand, int phase) {
  return phase == DCN_PHASE_FORWARD ? umma_fwd_workspace(g, operand) : umma_bwd_workspace(g, operand);
}

int umma_forward(const Geo& g, int operand, int flags, const void* x, const float* off, const void* wt,
                 const float* bias, void* out, void* workspace, cudaStream_t st) {
  (void)flags;
  return umma_forward_any(g, operand, x, off, wt, bias, (float*)out, workspace, st);
}

int umma_backward(const Geo& g, int operand, int flags, const void* x, const float* off, const void* wt,
                  const void* gout, float* gx, float* goff, float* gw, float* gb, void* workspace,
                  cudaStream_t st) {
  return umma_backward_any(g, operand, flags, x, off, wt, gout, gx, goff, gw, gb, workspace, st);
}

}  // namespace dcn
