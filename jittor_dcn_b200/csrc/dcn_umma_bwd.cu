// dcn_umma_bwd.cu — backward pass on the tensor path: orchestration.
//
//   grad_x, grad_offset   dcn_umma_bwd_data.cu: tcgen05 GEMM gA = gout * Wm whose TMEM accumulator is
//                 consumed in place by the bilinear col2im scatter (red.global.add.f32 into a
//                 channels-last grad copy) and the warp-shuffle coordinate-gradient reduction
//                 (both column layouts; shapes it cannot tile use dcn_simt.cu:bwd_data_kernel).
//   grad_weight   Torch layout, O <= 128 (bf16 operands: O <= 256): FUSED into the kernel above — its scatter warps already
//                 hold every sample's four corner values, so they also emit the blended sample as
//                 a bf16 hi/lo operand and a second TMEM accumulator set collects gW = gout^T * S:
//                 the whole backward touches x once.  Otherwise dcn_umma_fwd.cu (MODE_WGRAD):
//                 S re-sampled by the forward's plan / gather warps; nothing is materialised.
//   grad_bias     column sums of gout (dcn_simt.cu:bias_grad_kernel).
#include <cstdlib>

#include "dcn_umma.h"
#include "dcn_umma_common.cuh"

namespace dcn {

bool umma_wgrad_supported(const Geo& g, int operand);
size_t umma_xt_bytes(const Geo& g, int operand);
int umma_wgrad_any(const Geo& g, int operand, const void* xt, const float* off, const void* gout, float* gw,
                   uint8_t* gtiles, cudaStream_t st);
size_t umma_wgrad_gtile_bytes(const Geo& g, int operand);
bool umma_bwd_data_supported(const Geo& g, int operand);
size_t umma_bwd_data_wtile_bytes(const Geo& g, int operand);
bool umma_bwd_data_fuses_wgrad(const Geo& g, int operand);
int umma_bwd_data_any(const Geo& g, int operand, const void* xt, float* gxt, const float* off, const void* wt,
                      const void* gout, float* goff, float* gw, uint8_t* wtiles, uint8_t* gtiles,
                      cudaStream_t st, float* gb_fused = nullptr);
size_t umma_bwd_data_gtile_bytes(const Geo& g, int operand);
bool o_groups(const Geo& g, int* size);
Geo o_group_geo(const Geo& g, int o0, int size);

static size_t plan_bytes(const Geo& g) { return align_up(sizeof(Tap) * (size_t)g.B * g.P, 1024); }

static bool use_umma_data(const Geo& g, int operand) {
  if (knobs().bwd_data_simt) return false;
  return umma_bwd_data_supported(g, operand);
}

bool umma_bwd_supported(const Geo& g, int operand) {
  if (!umma_wgrad_supported(g, operand)) return false;
  // the generic data-gradient kernel is fp32 only: bf16 needs the tensor-path one
  return operand == DCN_OPERAND_FP32 || umma_bwd_data_supported(g, operand);
}

// shifted-view convolution kernels (dcn_conv.cu): preferred when they cover the shape
bool conv_offset_bwd_supported(const Geo& g);
size_t conv_offset_wtile_bytes(const Geo& g);
int conv_offset_backward(const Geo& g, const float* xt, float* gxt, const float* goff, const float* woff, float* gwoff,
                         uint8_t* wtiles, cudaStream_t st);

// ---- companion offset convolution (PLAIN problem) inside the layer's backward pass ---------------------------------
static bool plain_bwd_ok(const Geo& g) {
  Tiling t;
  if (!make_tiling(g, &t)) return false;
  if (conv_offset_bwd_supported(g)) return true;
  const Geo gp = plain_geo(g, t, false);
  return umma_bwd_data_supported(gp, DCN_OPERAND_FP32) && umma_bwd_data_fuses_wgrad(gp, DCN_OPERAND_FP32);
}

// Whole-layer backward (DCN span + offset conv) on the tensor path: fp32 operands, one output-channel group or more,
// data gradient on the tcgen05 kernel (so that the channels-last grad_x accumulator exists for both passes).
bool umma_layer_bwd_supported(const Geo& g) {
  int gsize;
  if (!o_groups(g, &gsize)) return false;
  const Geo g0 = o_group_geo(g, 0, gsize);
  return umma_bwd_supported(g0, DCN_OPERAND_FP32) && use_umma_data(g0, DCN_OPERAND_FP32) && plain_bwd_ok(g);
}

static size_t goff_bytes(const Geo& g) { return align_up(sizeof(float) * (size_t)g.B * 2 * g.N * g.HW, 1024); }

// [xt][rest as umma_bwd_workspace, its tile regions sized for the larger of the two passes][grad_offset]
size_t umma_layer_bwd_workspace(const Geo& g) {
  int gsize;
  if (!o_groups(g, &gsize)) return 0;
  const Geo g0 = o_group_geo(g, 0, gsize);
  Tiling t;
  if (!make_tiling(g, &t)) return 0;
  const Geo gp = plain_geo(g, t, false);
  const size_t tiles_dcn = umma_bwd_data_wtile_bytes(g0, DCN_OPERAND_FP32) + umma_bwd_data_gtile_bytes(g0, DCN_OPERAND_FP32);
  size_t tiles_pln = umma_bwd_data_wtile_bytes(gp, DCN_OPERAND_FP32) + umma_bwd_data_gtile_bytes(gp, DCN_OPERAND_FP32);
  if (conv_offset_bwd_supported(g)) tiles_pln = conv_offset_wtile_bytes(g);
  const size_t a = umma_xt_bytes(g, DCN_OPERAND_FP32) + (tiles_dcn > tiles_pln ? tiles_dcn : tiles_pln);
  const size_t c = umma_wgrad_gtile_bytes(g0, DCN_OPERAND_FP32);
  return umma_xt_bytes(g, DCN_OPERAND_FP32) + (a > c ? a : c) + goff_bytes(g);
}

size_t umma_bwd_workspace(const Geo& g, int operand) {
  // [xt] then either [gxt | Wm^T tiles | grad_out tiles] (tensor-path data gradient) or [sampling plan]
  // (generic); the unfused weight-gradient pass runs last and re-uses that region for its own staged
  // grad_out tiles
  const size_t a = umma_xt_bytes(g, DCN_OPERAND_FP32) + umma_bwd_data_wtile_bytes(g, operand) +
                   umma_bwd_data_gtile_bytes(g, operand);
  const size_t b = operand == DCN_OPERAND_FP32 ? plan_bytes(g) : 0;
  const size_t c = umma_wgrad_gtile_bytes(g, operand);
  const size_t m = a > b ? a : b;
  return umma_xt_bytes(g, operand) + (m > c ? m : c);
}

// g is the whole layer; the kernels run per output-channel group (dcn_umma_host.cu) — one group unless
// O > 256.  grad_x (via gxt) and grad_offset accumulate over the groups.
// woff != nullptr: whole-layer backward — the companion offset convolution's backward runs as a PLAIN pass between the
// DCN data gradient and the final transposition: grad_offset (goff) is its grad_out, its data gradient accumulates
// into the SAME channels-last grad_x buffer, gwoff / gboff receive its parameter gradients.
int umma_backward_any(const Geo& g, int operand, int flags, const void* xv, const float* off, const void* wtv,
                      const void* goutv, float* gx, float* goff, float* gw, float* gb, void* workspace,
                      cudaStream_t st, const float* woff, float* gwoff, float* gboff) {
  int gsize;
  if (!o_groups(g, &gsize)) {
    set_error("umma backward: O = %d cannot be split into groups", g.O);
    return DCN_ERR_UNSUPPORTED;
  }
  const Geo g0 = o_group_geo(g, 0, gsize);  // the largest group: sizes every scratch region
  const size_t esz = operand == DCN_OPERAND_BF16 ? 2 : 4;
  const void* x = xv;
  void* xt = workspace;
  uint8_t* rest = (uint8_t*)workspace + umma_xt_bytes(g, operand);
  Tiling t;
  if (!make_tiling(g, &t)) {
    set_error("umma backward: shape not tileable");
    return DCN_ERR_UNSUPPORTED;
  }
  int rc;
  // DCN_FLAG_XT_STAGED: the head of the workspace still holds the staged copy of x that dcn_forward
  // wrote for this very shape / operand (same layout in both phases) — skip the transpose
  if (!(flags & DCN_FLAG_XT_STAGED) && (rc = launch_nchw_to_nhwc(g, t, x, xt, operand, st))) return rc;
  const bool want_gx = !(flags & DCN_FLAG_NO_GRAD_X) && gx != nullptr;
  bool all_fused = true;
  if (operand != DCN_OPERAND_FP32 || use_umma_data(g0, operand)) {
    // DCN_FLAG_GRAD_X_FRAMED: grad_x IS the framed channels-last accumulator (the producer-side post-op reads it in
    // that layout, dcn_bn_relu_backward_staged): scatter straight into the caller's buffer, no transposition at the end
    const bool gx_framed = (flags & DCN_FLAG_GRAD_X_FRAMED) != 0;
    float* gxt = gx_framed ? gx : (float*)rest;
    uint8_t* wtiles = rest + umma_xt_bytes(g, DCN_OPERAND_FP32);
    uint8_t* gtiles = wtiles + umma_bwd_data_wtile_bytes(g0, operand);
    if (want_gx) DCN_CUDA_TRY(cudaMemsetAsync(gxt, 0, sizeof(float) * (size_t)g.B * xt_image_stride(g), st));
    DCN_CUDA_TRY(cudaMemsetAsync(goff, 0, sizeof(float) * (size_t)g.B * 2 * g.N * g.HW, st));
    // Torch layout: the staging pass of grad_out sums grad_bias on its way (every element passes through it once)
    const bool gb_rides = gb != nullptr && g.variant == DCN_VARIANT_TORCH;
    if (gb_rides) DCN_CUDA_TRY(cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)g.O, st));
    for (int o0 = 0; o0 < g.O; o0 += gsize) {
      const Geo gc = o_group_geo(g, o0, gsize);
      const bool fused = umma_bwd_data_fuses_wgrad(gc, operand);  // one pass over the samples yields gW as well
      all_fused = all_fused && fused;
      float* gwc = gw + (size_t)o0 * g.K;
      if (fused) DCN_CUDA_TRY(cudaMemsetAsync(gwc, 0, sizeof(float) * (size_t)gc.O * g.K, st));
      if ((rc = umma_bwd_data_any(gc, operand, xt, want_gx ? gxt : nullptr, off,
                                  (const uint8_t*)wtv + (size_t)o0 * g.K * esz,
                                  (const uint8_t*)goutv + (size_t)o0 * g.HW * esz, goff, gwc, wtiles, gtiles, st,
                                  gb_rides ? gb + o0 : nullptr)))
        return rc;
    }
    if (woff && conv_offset_bwd_supported(g)) {
      const Geo gp = plain_geo(g, t, false);
      uint8_t* wtiles2 = rest + umma_xt_bytes(g, DCN_OPERAND_FP32);
      if ((rc = conv_offset_backward(g, (const float*)xt, want_gx ? gxt : nullptr, goff, woff, gwoff, wtiles2, st)))
        return rc;
      if ((rc = launch_bias_grad(gp, goff, DCN_OPERAND_FP32, gboff, st))) return rc;
    } else if (woff) {
      const Geo gp = plain_geo(g, t, false);
      uint8_t* wtiles2 = rest + umma_xt_bytes(g, DCN_OPERAND_FP32);
      uint8_t* gtiles2 = wtiles2 + umma_bwd_data_wtile_bytes(gp, DCN_OPERAND_FP32);
      DCN_CUDA_TRY(cudaMemsetAsync(gwoff, 0, sizeof(float) * (size_t)gp.O * g.K, st));
      if ((rc = umma_bwd_data_any(gp, DCN_OPERAND_FP32, xt, want_gx ? gxt : nullptr, nullptr, woff, goff, nullptr, gwoff,
                                  wtiles2, gtiles2, st)))
        return rc;
      if ((rc = launch_bias_grad(gp, goff, DCN_OPERAND_FP32, gboff, st))) return rc;
    }
    if (want_gx && !gx_framed && (rc = launch_nhwc_to_nchw_add(g, t, gxt, gx, (flags & DCN_FLAG_ACCUM_GRAD_X) ? 1 : 0, st)))
      return rc;
    if (!gb_rides && (rc = launch_bias_grad(g, goutv, operand, gb, st))) return rc;
    if (all_fused) return DCN_OK;
  } else {
    if (woff) {
      set_error("layer backward: the data gradient of this shape does not run on the tensor path");
      return DCN_ERR_UNSUPPORTED;
    }
    all_fused = false;
    Tap* plan = (Tap*)rest;
    if ((rc = launch_plan(g, off, plan, st))) return rc;
    if ((rc = simt_backward(g, flags, (const float*)x, plan, (const float*)wtv, (const float*)goutv, gx, goff, gw,
                            gb, st, SIMT_BWD_DATA | SIMT_BWD_BIAS)))
      return rc;
  }
  // everything that lived in `rest` (gxt, tiles, plan) is dead by now: the unfused weight-gradient
  // pass of the remaining groups stages its grad_out tiles there
  for (int o0 = 0; o0 < g.O; o0 += gsize) {
    const Geo gc = o_group_geo(g, o0, gsize);
    const bool data_on_umma = operand != DCN_OPERAND_FP32 || use_umma_data(g0, operand);
    if (data_on_umma && umma_bwd_data_fuses_wgrad(gc, operand)) continue;
    if ((rc = umma_wgrad_any(gc, operand, xt, off, (const uint8_t*)goutv + (size_t)o0 * g.HW * esz,
                             gw + (size_t)o0 * g.K, rest, st)))
      return rc;
  }
  return DCN_OK;
}

}  // namespace dcn

namespace dcn {

// Whole-layer backward: see umma_backward_any.  The internal grad_offset buffer sits at the tail of the workspace.
int umma_layer_backward(const Geo& g, int flags, const void* x, const float* off, const float* woff, const void* wt,
                        const void* gout, float* gx, float* gwoff, float* gboff, float* gw, float* gb,
                        void* workspace, size_t workspace_bytes, cudaStream_t st) {
  float* goff = (float*)((uint8_t*)workspace + workspace_bytes - goff_bytes(g));
  return umma_backward_any(g, DCN_OPERAND_FP32, flags, x, off, wt, gout, gx, goff, gw, gb, workspace, st, woff, gwoff,
                           gboff);
}

}  // namespace dcn
