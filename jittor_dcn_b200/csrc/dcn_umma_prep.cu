// dcn_umma_prep.cu — layout staging for the tcgen05 kernels.
//
//   nchw_to_nhwc      x[B,C,H,W] -> xt[B,H+3,W+2,C'] (channels-last inside an all-zero frame,
//                     dcn_umma_common.cuh; channel-permuted for the Torch column layout) so that
//                     one bilinear corner of Gt channels is one contiguous, 16-byte-vectorisable
//                     run and zero padding needs no validity test.
//   nhwc_to_nchw      the inverse, for grad_x (optionally accumulating).
//   weight tiles      weight.reshape(O,K) (deform_conv.py:74 / train.py:133) split into bf16
//                     hi/lo and laid out exactly as the UMMA shared-memory images, so that the
//                     kernels fetch a K block with one linear bulk copy.
#include "dcn_umma_common.cuh"

namespace dcn {

// Both transposes are register-only: a thread owns a 4-channel x 4-pixel patch, reads it with four
// 4-element vector loads along one axis and writes it with four vector stores along the other (the
// 4 x 4 transpose is a renaming of registers).  Lanes: 8 along the channels x 4 along the pixels, so a
// warp reads 64-byte runs of 8 planes and writes 128-byte runs of 4 pixels — full sectors both ways,
// no shared memory, 16 KB in flight per block.  Block = 8 warps = 32 channels x 128 pixels.
template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<__nv_bfloat16> { typedef uint2 type; };

template <typename T>
__device__ __forceinline__ void load4(const T* p, bool vec_ok, int n_ok, T v[4]) {
  if (vec_ok) {
    typename Vec4<T>::type t = __ldg(reinterpret_cast<const typename Vec4<T>::type*>(p));
    memcpy(v, &t, sizeof(t));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = i < n_ok ? p[i] : T(0.f);
  }
}

// frame position (element offset of channel 0) of linear image pixel p
__device__ __forceinline__ size_t frame_px(const Geo& g, const FastDiv& divW, int p) {
  uint32_t y, xx;
  divW.divmod((uint32_t)p, y, xx);
  return (size_t)((y + 1) * (g.W + 2) + xx + 1) * g.C;
}

template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(Geo g, int variant, int G, int Cs, FastDiv divW,
                                                           const T* __restrict__ x, T* __restrict__ xt) {
  const int HWi = g.H * g.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = blockIdx.y * 32 + (lane & 7) * 4;              // destination channels d..d+3
  const int p = blockIdx.x * 128 + warp * 16 + (lane >> 3) * 4;  // pixels p..p+3
  if (d >= g.C || p >= HWi) return;
  const int b = blockIdx.z;
  const bool vec_ok = (HWi & 3) == 0;  // p % 4 == 0 always: plane rows are then 16-byte aligned
  T v[4][4];                           // [channel][pixel]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dd = d + k;
    const int c = variant == DCN_VARIANT_TORCH ? (dd % G) * Cs + dd / G : dd;  // inverse permutation
    load4(x + ((size_t)b * g.C + c) * HWi + p, vec_ok, HWi - p, v[k]);
  }
  T* img = xt + (size_t)b * xt_image_stride(g) + d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (p + i >= HWi) break;
    T o[4] = {v[0][i], v[1][i], v[2][i], v[3][i]};
    typename Vec4<T>::type t;
    memcpy(&t, o, sizeof(t));
    *reinterpret_cast<typename Vec4<T>::type*>(img + frame_px(g, divW, p + i)) = t;
  }
}

// zero the frame of every image: rows 0, H+1, H+2 and columns 0, W+1 (16-byte stores; C % 4 == 0)
template <typename T>
__global__ void __launch_bounds__(256) xt_frame_zero_kernel(Geo g, T* __restrict__ xt) {
  constexpr int V = 16 / sizeof(T);
  const int cv = g.C / V;                                  // 16-byte chunks per pixel
  const int fw = g.W + 2, frame_px = 3 * fw + 2 * g.H;     // frame pixels per image
  const size_t img_stride = xt_image_stride(g);
  const long long total = (long long)g.B * frame_px * cv;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % cv);
    const long long r = i / cv;
    const int f = (int)(r % frame_px), b = (int)(r / frame_px);
    int y, xx;
    if (f < fw) { y = 0; xx = f; }
    else if (f < 3 * fw) { y = g.H + 1 + (f - fw) / fw; xx = (f - fw) % fw; }
    else { const int k = f - 3 * fw; y = 1 + (k >> 1); xx = (k & 1) ? g.W + 1 : 0; }
    *reinterpret_cast<uint4*>(xt + (size_t)b * img_stride + (size_t)(y * fw + xx) * g.C + c * V) =
        make_uint4(0, 0, 0, 0);
  }
}

// zero the frame of a float staging copy (the interior is written by someone else: bn_relu_stage_kernel)
int launch_xt_frame_zero(const Geo& g, float* xt, cudaStream_t st) {
  KernelScope scope("xt_frame_zero_kernel", st);
  const long long chunks = (long long)g.B * (3 * (g.W + 2) + 2 * g.H) * (g.C / 4);
  const int blocks = (int)((chunks + 255) / 256 < 4096 ? (chunks + 255) / 256 : 4096);
  xt_frame_zero_kernel<float><<<blocks, 256, 0, st>>>(g, xt);
  DCN_KERNEL_CHECK("xt_frame_zero_kernel");
  return DCN_OK;
}

int launch_nchw_to_nhwc(const Geo& g, const Tiling& t, const void* x, void* xt, int operand, cudaStream_t st) {
  const int HWi = g.H * g.W;
  dim3 grid((HWi + 127) / 128, (g.C + 31) / 32, g.B);
  const FastDiv divW = FastDiv::make(g.W);
  {
    KernelScope scope("xt_frame_zero_kernel", st);
    const long long chunks = (long long)g.B * (3 * (g.W + 2) + 2 * g.H) * (g.C / (operand == DCN_OPERAND_BF16 ? 8 : 4));
    const int blocks = (int)((chunks + 255) / 256 < 4096 ? (chunks + 255) / 256 : 4096);
    if (operand == DCN_OPERAND_BF16)
      xt_frame_zero_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(g, (__nv_bfloat16*)xt);
    else
      xt_frame_zero_kernel<float><<<blocks, 256, 0, st>>>(g, (float*)xt);
    DCN_KERNEL_CHECK("xt_frame_zero_kernel");
  }
  KernelScope scope("nchw_to_nhwc_kernel", st);
  if (operand == DCN_OPERAND_BF16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, divW, (const __nv_bfloat16*)x,
                                                            (__nv_bfloat16*)xt);
  else
    nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, divW, (const float*)x, (float*)xt);
  DCN_KERNEL_CHECK("nchw_to_nhwc_kernel");
  return DCN_OK;
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(Geo g, int variant, int G, int Cs, FastDiv divW,
                                                           int accumulate,
                                                           const float* __restrict__ gxt,
                                                           float* __restrict__ gx) {
  const int HWi = g.H * g.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = blockIdx.y * 32 + (lane & 7) * 4;
  const int p = blockIdx.x * 128 + warp * 16 + (lane >> 3) * 4;
  if (d >= g.C || p >= HWi) return;
  const int b = blockIdx.z;
  const float* img = gxt + (size_t)b * xt_image_stride(g) + d;
  float v[4][4];  // [pixel][channel]
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p + i < HWi) t = __ldg(reinterpret_cast<const float4*>(img + frame_px(g, divW, p + i)));
    v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
  }
  const bool vec_ok = (HWi & 3) == 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dd = d + k;
    const int c = variant == DCN_VARIANT_TORCH ? (dd % G) * Cs + dd / G : dd;
    float* dst = gx + ((size_t)b * g.C + c) * HWi + p;
    if (vec_ok) {
      float4 o = make_float4(v[0][k], v[1][k], v[2][k], v[3][k]);
      if (accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(dst);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *reinterpret_cast<float4*>(dst) = o;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p + i < HWi) dst[i] = accumulate ? dst[i] + v[i][k] : v[i][k];
    }
  }
}

int launch_nhwc_to_nchw_add(const Geo& g, const Tiling& t, const float* gxt, float* gx, int accumulate,
                            cudaStream_t st) {
  const int HWi = g.H * g.W;
  dim3 grid((HWi + 127) / 128, (g.C + 31) / 32, g.B);
  KernelScope scope("nhwc_to_nchw_kernel", st);
  nhwc_to_nchw_kernel<<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, FastDiv::make(g.W), accumulate, gxt, gx);
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
    const float v = (jj < g.K && o < g.o_valid) ? (float)wt[wt_index(g, o, jj)] : 0.f;
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
