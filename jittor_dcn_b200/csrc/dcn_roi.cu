// dcn_roi.cu — DeformRoIPool / DeformPSRoIPool (deform_conv.py:83-157, 160-241) for the one output size
// the reference's code can execute: pooled 1 x 1.  Both modules sum the per-bin values over the bin axis
// (`.sum(dim=2)`, :137-157 / :236-240) and then reshape [num_rois, C] to [num_rois, C, pooled_h, pooled_w], which only
// type-checks for a single bin.  With one bin both reduce to ONE bilinear-like sample per (roi, channel):
//   scaled roi  (x1, y1, x2, y2) = rois[:, 1:5] * spatial_scale;  roi_w = max(x2 - x1, 1e-6), roi_h likewise   (:96-100)
//   bin centre  cx = x1 + 0.5 * roi_w + off_x * roi_w [* trans_std],  cy alike                                  (:108-118, :206-214)
//   corners     x0 = clamp(floor(cx)), x1' = clamp(floor(cx) + 1), dx = cx - x0 (x0 AFTER the clamp, so the weights
//               leave [0, 1] for centres outside the map — reproduced as written)                                 (:120-136)
//   out[r, c]   = f[y0,x0](1-dx)(1-dy) + f[y1,x0](1-dx)dy + f[y0,x1]dx(1-dy) + f[y1,x1]dx dy                       (:138-156)
// The PS variant reads channel c*1*1 + 0 = c (:226) and offsets[:, 0] / offsets[:, 1] (:209-210).
// One block per roi, threads over channels; the backward pass scatters into grad_features with red.global.add and
// block-reduces the two offset gradients.  These are latency-sized ops (R x C x 4 loads); no tensor path applies.
#include <cuda_runtime.h>

#include "dcn_common.cuh"

namespace dcn {

namespace roi {

struct Sample {
  int b, y0, x0, y1, x1;
  float dx, dy, roi_w, roi_h;
};

// the reference's float32 op chain, one rounding per op
__device__ __forceinline__ Sample make_sample(const float* __restrict__ rois, const float* __restrict__ offsets, int r,
                                              int H, int W, float spatial_scale, float off_scale, int no_trans) {
  Sample s;
  const float* q = rois + (size_t)r * 5;
  s.b = (int)q[0];                                           // rois[:, 0].long(): truncation
  const float x1 = __fmul_rn(q[1], spatial_scale), y1 = __fmul_rn(q[2], spatial_scale);
  const float x2 = __fmul_rn(q[3], spatial_scale), y2 = __fmul_rn(q[4], spatial_scale);
  s.roi_w = fmaxf(__fsub_rn(x2, x1), 1e-6f);
  s.roi_h = fmaxf(__fsub_rn(y2, y1), 1e-6f);
  // pooled / part size 1: bin_w = roi_w / 1, centre = x1 + (0 + 0.5) * bin_w
  float cx = __fadd_rn(x1, __fmul_rn(0.5f, s.roi_w)), cy = __fadd_rn(y1, __fmul_rn(0.5f, s.roi_h));
  if (!no_trans) {
    float ox = __fmul_rn(offsets[(size_t)r * 2], s.roi_w), oy = __fmul_rn(offsets[(size_t)r * 2 + 1], s.roi_h);
    if (off_scale != 1.0f) {                                 // DeformPSRoIPool: * trans_std
      ox = __fmul_rn(ox, off_scale);
      oy = __fmul_rn(oy, off_scale);
    }
    cx = __fadd_rn(cx, ox);
    cy = __fadd_rn(cy, oy);
  }
  const int fx = sat_int(floorf(cx)), fy = sat_int(floorf(cy));
  s.x0 = min(max(fx, 0), W - 1);
  s.x1 = min(max(fx + 1, 0), W - 1);
  s.y0 = min(max(fy, 0), H - 1);
  s.y1 = min(max(fy + 1, 0), H - 1);
  s.dx = __fsub_rn(cx, (float)s.x0);
  s.dy = __fsub_rn(cy, (float)s.y0);
  return s;
}

__global__ void __launch_bounds__(256) roi_fwd_kernel(int B, int C, int H, int W, const float* __restrict__ f,
                                                      const float* __restrict__ rois, const float* __restrict__ offsets,
                                                      float spatial_scale, float off_scale, int no_trans,
                                                      float* __restrict__ out) {
  const int r = blockIdx.x;
  const Sample s = make_sample(rois, offsets, r, H, W, spatial_scale, off_scale, no_trans);
  if (s.b < 0 || s.b >= B) {                                 // an index the reference would fault on: zeros
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)r * C + c] = 0.f;
    return;
  }
  const float ex = __fsub_rn(1.0f, s.dx), ey = __fsub_rn(1.0f, s.dy);
  const float w00 = __fmul_rn(ex, ey), w01 = __fmul_rn(ex, s.dy), w10 = __fmul_rn(s.dx, ey), w11 = __fmul_rn(s.dx, s.dy);
  const size_t HW = (size_t)H * W;
  const float* img = f + (size_t)s.b * C * HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = img + (size_t)c * HW;
    const float v00 = __ldg(p + (size_t)s.y0 * W + s.x0), v01 = __ldg(p + (size_t)s.y1 * W + s.x0);
    const float v10 = __ldg(p + (size_t)s.y0 * W + s.x1), v11 = __ldg(p + (size_t)s.y1 * W + s.x1);
    // val00 + val01 + val10 + val11, left to right (:156 / :240)
    out[(size_t)r * C + c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, w00), __fmul_rn(v01, w01)), __fmul_rn(v10, w10)),
                                       __fmul_rn(v11, w11));
  }
}

__global__ void __launch_bounds__(256) roi_bwd_kernel(int B, int C, int H, int W, const float* __restrict__ f,
                                                      const float* __restrict__ rois, const float* __restrict__ offsets,
                                                      float spatial_scale, float off_scale, int no_trans,
                                                      const float* __restrict__ gout, float* __restrict__ gf,
                                                      float* __restrict__ goffsets) {
  __shared__ float red[2][8];
  const int r = blockIdx.x;
  const Sample s = make_sample(rois, offsets, r, H, W, spatial_scale, off_scale, no_trans);
  float gdx = 0.f, gdy = 0.f;
  if (s.b >= 0 && s.b < B) {
    const float ex = 1.0f - s.dx, ey = 1.0f - s.dy;
    const float w00 = ex * ey, w01 = ex * s.dy, w10 = s.dx * ey, w11 = s.dx * s.dy;
    const size_t HW = (size_t)H * W;
    const float* img = f + (size_t)s.b * C * HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float g = gout[(size_t)r * C + c];
      const size_t o00 = (size_t)c * HW + (size_t)s.y0 * W + s.x0, o01 = (size_t)c * HW + (size_t)s.y1 * W + s.x0;
      const size_t o10 = (size_t)c * HW + (size_t)s.y0 * W + s.x1, o11 = (size_t)c * HW + (size_t)s.y1 * W + s.x1;
      if (gf) {
        float* gimg = gf + (size_t)s.b * C * HW;
        atomicAdd(gimg + o00, g * w00);
        atomicAdd(gimg + o01, g * w01);
        atomicAdd(gimg + o10, g * w10);
        atomicAdd(gimg + o11, g * w11);
      }
      const float v00 = __ldg(img + o00), v01 = __ldg(img + o01), v10 = __ldg(img + o10), v11 = __ldg(img + o11);
      gdx += g * ((v10 - v00) * ey + (v11 - v01) * s.dy);
      gdy += g * ((v01 - v00) * ex + (v11 - v10) * s.dx);
    }
  }
  if (!goffsets) return;
  // block reduction of the two coordinate gradients
  for (int o = 16; o; o >>= 1) {
    gdx += __shfl_xor_sync(0xffffffffu, gdx, o);
    gdy += __shfl_xor_sync(0xffffffffu, gdy, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = gdx;
    red[1][warp] = gdy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sx = 0.f, sy = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      sx += red[0][w];
      sy += red[1][w];
    }
    // cx = ... + off_x * roi_w [* trans_std]
    goffsets[(size_t)r * 2] = no_trans ? 0.f : sx * s.roi_w * off_scale;
    goffsets[(size_t)r * 2 + 1] = no_trans ? 0.f : sy * s.roi_h * off_scale;
  }
}

}  // namespace roi

}  // namespace dcn

using namespace dcn;

extern "C" {

static int roi_check(int kind, int B, int C, int H, int W, int R) {
  if (kind != DCN_ROI_POOL && kind != DCN_PSROI_POOL) {
    set_error("roi pool: unknown kind %d", kind);
    return DCN_ERR_BAD_SHAPE;
  }
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || R < 0) {
    set_error("roi pool: bad extents B=%d C=%d H=%d W=%d R=%d", B, C, H, W, R);
    return DCN_ERR_BAD_SHAPE;
  }
  return DCN_OK;
}

int dcn_roi_pool_forward(int kind, int B, int C, int H, int W, int R, const void* features, const void* rois,
                         const void* offsets, float spatial_scale, float trans_std, int no_trans, void* out,
                         void* stream) {
  int rc = roi_check(kind, B, C, H, W, R);
  if (rc) return rc;
  if (!features || !rois || !out || (!offsets && !no_trans)) return DCN_ERR_NULL_POINTER;
  if (R == 0) return DCN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float off_scale = kind == DCN_PSROI_POOL ? trans_std : 1.0f;
  // DeformRoIPool has no no_trans switch (deform_conv.py:84): offsets always apply
  const int nt = kind == DCN_PSROI_POOL ? no_trans : 0;
  KernelScope scope("roi_pool_fwd_kernel", st);
  roi::roi_fwd_kernel<<<R, C < 256 ? (C + 31) / 32 * 32 : 256, 0, st>>>(B, C, H, W, (const float*)features, (const float*)rois,
                                                                        (const float*)offsets, spatial_scale, off_scale, nt,
                                                                        (float*)out);
  DCN_KERNEL_CHECK("roi_pool_fwd_kernel");
  return DCN_OK;
}

int dcn_roi_pool_backward(int kind, int B, int C, int H, int W, int R, const void* features, const void* rois,
                          const void* offsets, float spatial_scale, float trans_std, int no_trans, const void* grad_out,
                          void* grad_features, void* grad_offsets, void* stream) {
  int rc = roi_check(kind, B, C, H, W, R);
  if (rc) return rc;
  if (!features || !rois || !grad_out || (!offsets && !no_trans)) return DCN_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_features) DCN_CUDA_TRY(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * H * W, st));
  if (R == 0) return DCN_OK;
  const float off_scale = kind == DCN_PSROI_POOL ? trans_std : 1.0f;
  const int nt = kind == DCN_PSROI_POOL ? no_trans : 0;
  KernelScope scope("roi_pool_bwd_kernel", st);
  roi::roi_bwd_kernel<<<R, C < 256 ? (C + 31) / 32 * 32 : 256, 0, st>>>(B, C, H, W, (const float*)features, (const float*)rois,
                                                                        (const float*)offsets, spatial_scale, off_scale, nt,
                                                                        (const float*)grad_out, (float*)grad_features,
                                                                        (float*)grad_offsets);
  DCN_KERNEL_CHECK("roi_pool_bwd_kernel");
  return DCN_OK;
}

}  // extern "C"
