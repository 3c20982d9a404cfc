// dcn_umma_prep.cu — layout staging for the tcgen05 kernels.
//
//   nchw_to_nhwc      x[B,C,H,W] -> xt[B,H,W,C'] (channels-last, channel-permuted for the Torch
//                     column layout) so that one bilinear corner of Gt channels is one
//                     contiguous, 16-byte-vectorisable run.
//   nhwc_to_nchw      the inverse, for grad_x (optionally accumulating).
//   weight tiles      weight.reshape(O,K) (deform_conv.py:74 / train.py:133) split into bf16
//                     hi/lo and laid out exactly as the UMMA shared-memory images, so that the
//                     kernels fetch a K block with one linear bulk copy.
#include "dcn_umma_common.cuh"

namespace dcn {

// Both transposes move 32-channel x 128-pixel panels per block: 16 independent 128-byte-coalesced
// loads per thread are in flight before the single barrier (4 sub-tiles of 32 x 32 through
// padded shared memory), then 16 coalesced stores.
constexpr int kSub = 4;  // 32-pixel sub-tiles per block

template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(Geo g, int variant, int G, int Cs,
                                                           const T* __restrict__ x, T* __restrict__ xt) {
  __shared__ T tile[kSub][32][33];
  const int HWi = g.H * g.W;
  const int b = blockIdx.z, p0 = blockIdx.x * (32 * kSub), d0 = blockIdx.y * 32;  // d = destination channel
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < kSub; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = d0 + ty + 8 * i, p = p0 + 32 * s + tx;
      if (d < g.C && p < HWi) {
        const int c = variant == DCN_VARIANT_TORCH ? (d % G) * Cs + d / G : d;  // inverse permutation
        tile[s][ty + 8 * i][tx] = x[((size_t)b * g.C + c) * HWi + p];
      }
    }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kSub; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + 32 * s + ty + 8 * i, d = d0 + tx;
      if (d < g.C && p < HWi) xt[((size_t)b * (HWi + 1) + p) * g.C + d] = tile[s][tx][ty + 8 * i];
    }
  // the zero pad pixel that closes the image (target of out-of-image corners)
  if (blockIdx.x == 0 && threadIdx.x < 32 && d0 + tx < g.C)
    xt[((size_t)b * (HWi + 1) + HWi) * g.C + d0 + tx] = T(0.f);
}

int launch_nchw_to_nhwc(const Geo& g, const Tiling& t, const void* x, void* xt, int operand, cudaStream_t st) {
  const int HWi = g.H * g.W;
  dim3 grid((HWi + 32 * kSub - 1) / (32 * kSub), (g.C + 31) / 32, g.B);
  KernelScope scope("nchw_to_nhwc_kernel", st);
  if (operand == DCN_OPERAND_BF16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, (const __nv_bfloat16*)x,
                                                            (__nv_bfloat16*)xt);
  else
    nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, (const float*)x, (float*)xt);
  DCN_KERNEL_CHECK("nchw_to_nhwc_kernel");
  return DCN_OK;
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(Geo g, int variant, int G, int Cs,
                                                           int accumulate,
                                                           const float* __restrict__ gxt,
                                                           float* __restrict__ gx) {
  __shared__ float tile[kSub][32][33];
  const int HWi = g.H * g.W;
  const int b = blockIdx.z, p0 = blockIdx.x * (32 * kSub), d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < kSub; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + 32 * s + ty + 8 * i, d = d0 + tx;
      if (d < g.C && p < HWi) tile[s][ty + 8 * i][tx] = __ldg(gxt + ((size_t)b * (HWi + 1) + p) * g.C + d);
    }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kSub; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = d0 + ty + 8 * i, p = p0 + 32 * s + tx;
      if (d < g.C && p < HWi) {
        const int c = variant == DCN_VARIANT_TORCH ? (d % G) * Cs + d / G : d;
        float* dst = gx + ((size_t)b * g.C + c) * HWi + p;
        const float v = tile[s][tx][ty + 8 * i];
        *dst = accumulate ? *dst + v : v;
      }
    }
}

int launch_nhwc_to_nchw_add(const Geo& g, const Tiling& t, const float* gxt, float* gx, int accumulate,
                            cudaStream_t st) {
  const int HWi = g.H * g.W;
  dim3 grid((HWi + 32 * kSub - 1) / (32 * kSub), (g.C + 31) / 32, g.B);
  KernelScope scope("nhwc_to_nchw_kernel", st);
  nhwc_to_nchw_kernel<<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, accumulate, gxt, gx);
  DCN_KERNEL_CHECK("nhwc_to_nchw_kernel");
  return DCN_OK;
}

// tiles[kb][hl][K-major SW128 image of O rows x 64 k]; columns j >= K are zero.  bf16 weights
// have no lo image (NIMG = 1).
template <typename T>
__global__ void __launch_bounds__(256) weight_tiles_fwd_kernel(Geo g, int KB, const T* __restrict__ wt,
                                                               uint8_t* __restrict__ tiles) {
  constexpr int NIMG = sizeof(T) == 2 ? 1 : 2;
  const int total = g.O * KB * 64;
  const uint32_t tile_bytes = (uint32_t)g.O * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int o = i / (KB * 64), jj = i - o * (KB * 64), kb = jj >> 6, kk = jj & 63;
    const float v = jj < g.K ? (float)wt[wt_index(g, o, jj)] : 0.f;
    __nv_bfloat16 hi, lo;
    ptx::split_bf16(v, hi, lo);
    uint8_t* base = tiles + (size_t)kb * NIMG * tile_bytes + ptx::kmajor_sw128_off(o, kk);
    *reinterpret_cast<__nv_bfloat16*>(base) = hi;
    if (NIMG == 2) *reinterpret_cast<__nv_bfloat16*>(base + tile_bytes) = lo;
  }
}

int launch_weight_tiles_fwd(const Geo& g, const Tiling& t, const void* wt, uint8_t* tiles, int operand,
                            cudaStream_t st) {
  const int total = g.O * t.KB * 64;
  KernelScope scope("weight_tiles_fwd_kernel", st);
  if (operand == DCN_OPERAND_BF16)
    weight_tiles_fwd_kernel<__nv_bfloat16><<<min((total + 255) / 256, 2048), 256, 0, st>>>(
        g, t.KB, (const __nv_bfloat16*)wt, tiles);
  else
    weight_tiles_fwd_kernel<float><<<min((total + 255) / 256, 2048), 256, 0, st>>>(g, t.KB, (const float*)wt, tiles);
  DCN_KERNEL_CHECK("weight_tiles_fwd_kernel");
  return DCN_OK;
}

}  // namespace dcn
