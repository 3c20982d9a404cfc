// dcn_umma.h — interface of the tcgen05 / TMEM kernel family (dcn_umma_*.cu).
#pragma once
#include "dcn_common.cuh"

namespace dcn {

// Does the tensor-core path cover this problem?  (channel / output counts must tile the
// UMMA shapes; everything else goes to the generic kernels in dcn_simt.cu.)
bool umma_supported(const Geo& g, int operand, int phase);
size_t umma_workspace_bytes(const Geo& g, int operand, int phase);

int umma_forward(const Geo& g, int operand, int flags, const void* x, const float* off,
                 const void* wt, const float* bias, void* out, void* workspace, cudaStream_t st);
int umma_backward(const Geo& g, int operand, int flags, const void* x, const float* off,
                  const void* wt, const void* gout, float* gx, float* goff, float* gw, float* gb,
                  void* workspace, cudaStream_t st);

// ---- whole layer (companion offset conv as a PLAIN problem + DCN span), fp32 operands
bool umma_offset_conv_fwd_supported(const Geo& g);
size_t umma_offset_conv_fwd_workspace(const Geo& g);
int umma_offset_conv_forward(const Geo& g, const void* x, const float* woff, const float* boff, float* offset_out,
                             void* workspace, cudaStream_t st, bool stage_x);
bool umma_layer_bwd_supported(const Geo& g);
size_t umma_layer_bwd_workspace(const Geo& g);
int umma_layer_backward(const Geo& g, int flags, const void* x, const float* off, const float* woff, const void* wt,
                        const void* gout, float* gx, float* gwoff, float* gboff, float* gw, float* gb,
                        void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace dcn
