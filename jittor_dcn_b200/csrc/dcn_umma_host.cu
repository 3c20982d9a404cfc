// dcn_umma_host.cu — dispatch of the tcgen05 kernel family behind dcn_umma.h.
#include "dcn_umma.h"
#include "dcn_umma_common.cuh"

namespace dcn {

bool umma_fwd_supported(const Geo& g, int operand);
size_t umma_fwd_workspace(const Geo& g);
int umma_forward_fp32(const Geo& g, const float* x, const float* off, const float* wt, const float* bias,
                      float* out, void* workspace, cudaStream_t st);
bool umma_bwd_supported(const Geo& g, int operand);
size_t umma_bwd_workspace(const Geo& g);
int umma_backward_fp32(const Geo& g, int flags, const float* x, const float* off, const float* wt,
                       const float* gout, float* gx, float* goff, float* gw, float* gb, void* workspace,
                       cudaStream_t st);

bool umma_supported(const Geo& g, int operand, int phase) {
  if (phase == DCN_PHASE_FORWARD) return umma_fwd_supported(g, operand);
  if (phase == DCN_PHASE_BACKWARD) return umma_bwd_supported(g, operand);
  return false;
}

size_t umma_workspace_bytes(const Geo& g, int operand, int phase) {
  (void)operand;
  return phase == DCN_PHASE_FORWARD ? umma_fwd_workspace(g) : umma_bwd_workspace(g);
}

int umma_forward(const Geo& g, int operand, int flags, const void* x, const float* off, const void* wt,
                 const float* bias, void* out, void* workspace, cudaStream_t st) {
  (void)flags;
  if (operand != DCN_OPERAND_FP32) {
    set_error("umma forward: operand mode %d not implemented", operand);
    return DCN_ERR_UNSUPPORTED;
  }
  return umma_forward_fp32(g, (const float*)x, off, (const float*)wt, bias, (float*)out, workspace, st);
}

int umma_backward(const Geo& g, int operand, int flags, const void* x, const float* off, const void* wt,
                  const void* gout, float* gx, float* goff, float* gw, float* gb, void* workspace,
                  cudaStream_t st) {
  if (operand != DCN_OPERAND_FP32) {
    set_error("umma backward: operand mode %d not implemented", operand);
    return DCN_ERR_UNSUPPORTED;
  }
  return umma_backward_fp32(g, flags, (const float*)x, off, (const float*)wt, (const float*)gout, gx, goff,
                            gw, gb, workspace, st);
}

}  // namespace dcn
