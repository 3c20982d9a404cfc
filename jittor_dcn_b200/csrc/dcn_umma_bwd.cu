// dcn_umma_bwd.cu — backward pass on the tensor path.
//
//   grad_weight   tcgen05 GEMM gW = gout^T * S with S re-sampled by the forward's plan / gather
//                 warps (dcn_umma_fwd.cu, MODE_WGRAD); nothing is materialised.
//   grad_x, grad_offset   (this revision) generic CUDA-core kernel: gA = gout * Wm, bilinear
//                 col2im scatter with red.global.add.f32 and coordinate-gradient reduction
//                 (dcn_simt.cu:bwd_data_kernel).
//   grad_bias     column sums of gout (dcn_simt.cu:bias_grad_kernel).
#include "dcn_umma.h"
#include "dcn_umma_common.cuh"

namespace dcn {

bool umma_wgrad_supported(const Geo& g, int operand);
size_t umma_xt_bytes(const Geo& g);
int umma_wgrad_fp32(const Geo& g, const float* xt, const float* off, const float* gout, float* gw,
                    cudaStream_t st);

static size_t plan_bytes(const Geo& g) { return align_up(sizeof(Tap) * (size_t)g.B * g.P, 1024); }

bool umma_bwd_supported(const Geo& g, int operand) { return umma_wgrad_supported(g, operand); }

size_t umma_bwd_workspace(const Geo& g) { return plan_bytes(g) + umma_xt_bytes(g); }

int umma_backward_fp32(const Geo& g, int flags, const float* x, const float* off, const float* wt,
                       const float* gout, float* gx, float* goff, float* gw, float* gb, void* workspace,
                       cudaStream_t st) {
  Tap* plan = (Tap*)workspace;
  float* xt = (float*)((uint8_t*)workspace + plan_bytes(g));
  Tiling t;
  if (!make_tiling(g, &t)) {
    set_error("umma backward: shape not tileable");
    return DCN_ERR_UNSUPPORTED;
  }
  int rc;
  if ((rc = launch_plan(g, off, plan, st))) return rc;
  if ((rc = simt_backward(g, flags, x, plan, wt, gout, gx, goff, gw, gb, st, SIMT_BWD_DATA | SIMT_BWD_BIAS)))
    return rc;
  if ((rc = launch_nchw_to_nhwc(g, t, x, xt, st))) return rc;
  return umma_wgrad_fp32(g, xt, off, gout, gw, st);
}

}  // namespace dcn
