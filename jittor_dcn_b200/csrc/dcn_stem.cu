// dcn_stem.cu — the producer of the first DCN layer's input in the detector (train.py:145,166 / :307,328:
// `conv1 = Conv2d(1, 16, 3, 1, 1)`): a 3 x 3, stride 1, padding 1 convolution with a handful of input channels.
//
// Why it is here: with the DCN layers on the engine the framework's kernels for this one layer were 4.7 ms of the
// 24 ms detector step at batch 1024 (profiles/r2_detector_kernels_before_stem.txt: an NHWC implicit-GEMM forward between
// two layout passes, a separate bias add, a 1.4 ms weight-gradient kernel and a 0.5 ms bias-gradient reduction), for an
// op whose HBM floor is 0.17 ms each way (one pass over the [B,16,H,W] tensor).  There is nothing to put on tensor cores
// (K = 9 * Cin); both kernels are plain streaming kernels:
//   forward   thread = 4 consecutive output pixels of one row, all O channels: 3 x 6 input values in registers, weights
//             broadcast from shared memory, O float4 stores (coalesced along w)
//   backward  warp = (row, group of 4 output channels), lane = 4 consecutive pixels: 4 float4 loads of grad_out per
//             row, 36 + 4 accumulators per thread; shuffle + shared-memory reduction, then one atomicAdd per block and
//             gradient entry.  No input gradient: the input is the network input.
// fp32 throughout, FMA in the written order (tap-major), so the result differs from cuDNN's only by summation order.
#include <cuda_runtime.h>

#include <algorithm>

#include "dcn_common.cuh"

namespace dcn {

namespace stem {

constexpr int kMaxCin = 4;

// x values a thread needs for 4 consecutive output pixels of row h: rows h-1..h+1, columns w0-1..w0+4 (zeros outside)
__device__ __forceinline__ void load_window(const float* __restrict__ xp, int H, int W, int h, int w0, float (&v)[3][6]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int y = h + r - 1;
    if (y < 0 || y >= H) {
#pragma unroll
      for (int i = 0; i < 6; ++i) v[r][i] = 0.f;
      continue;
    }
    const float* row = xp + (size_t)y * W;
    const float4 m = *reinterpret_cast<const float4*>(row + w0);
    v[r][0] = w0 > 0 ? row[w0 - 1] : 0.f;
    v[r][1] = m.x;
    v[r][2] = m.y;
    v[r][3] = m.z;
    v[r][4] = m.w;
    v[r][5] = w0 + 4 < W ? row[w0 + 4] : 0.f;
  }
}

template <int O>
__global__ void __launch_bounds__(128) stem_fwd_kernel(int B, int Cin, int H, int W, const float* __restrict__ x,
                                                       const float* __restrict__ wt, const float* __restrict__ bias,
                                                       float* __restrict__ out) {
  __shared__ float sw[O * kMaxCin * 9 + O];
  for (int i = threadIdx.x; i < O * Cin * 9; i += blockDim.x) sw[i] = wt[i];
  for (int i = threadIdx.x; i < O; i += blockDim.x) sw[O * Cin * 9 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int W4 = W >> 2;
  const long long total = (long long)B * H * W4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int w0 = (int)(t % W4) * 4;
    const long long bh = t / W4;
    const int h = (int)(bh % H), b = (int)(bh / H);
    float acc[O][4];
#pragma unroll
    for (int o = 0; o < O; ++o) {
      const float bv = sw[O * Cin * 9 + o];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[o][i] = bv;
    }
    for (int ci = 0; ci < Cin; ++ci) {
      float v[3][6];
      load_window(x + ((size_t)b * Cin + ci) * H * W, H, W, h, w0, v);
#pragma unroll
      for (int o = 0; o < O; ++o) {
        const float* wo = sw + (o * Cin + ci) * 9;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float wv = wo[r * 3 + c];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[o][i] = fmaf(wv, v[r][c + i], acc[o][i]);
          }
      }
    }
    float* op = out + (((size_t)b * O) * H + h) * W + w0;
#pragma unroll
    for (int o = 0; o < O; ++o)
      *reinterpret_cast<float4*>(op + (size_t)o * H * W) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
  }
}

// grad_weight [O, Cin, 3, 3] and grad_bias [O], both zeroed by the caller
template <int O>
__global__ void __launch_bounds__(256) stem_bwd_kernel(int B, int Cin, int H, int W, const float* __restrict__ x,
                                                       const float* __restrict__ gout, float* __restrict__ gw,
                                                       float* __restrict__ gb) {
  constexpr int OG = O / 4;                       // groups of 4 output channels
  constexpr int kWarps = 8;
  static_assert(kWarps % OG == 0 || OG % kWarps == 0, "warps <-> channel groups");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W4 = W >> 2, segs = (W4 + 31) / 32;   // 128-pixel segments per row
  // unit of work = (b, h, segment, channel group); consecutive warps take the groups of one (b, h, segment), so the
  // x window they share comes out of L1
  const long long units = (long long)B * H * segs * OG;
  __shared__ float red[kWarps][40];
  for (int ci = 0; ci < Cin; ++ci) {
    float aw[4][9], ab[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ab[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) aw[j][k] = 0.f;
    }
    const int og = warp % OG;                     // gridDim.x * kWarps is a multiple of OG: a warp keeps its group
    for (long long u = (long long)blockIdx.x * kWarps + warp; u < units; u += (long long)gridDim.x * kWarps) {
      long long r = u / OG;
      const int seg = (int)(r % segs);
      r /= segs;
      const int h = (int)(r % H), b = (int)(r / H);
      const int w0 = (seg * 32 + lane) * 4;
      if (w0 >= W) continue;
      float v[3][6];
      load_window(x + ((size_t)b * Cin + ci) * H * W, H, W, h, w0, v);
      const float* gp = gout + (((size_t)b * O + og * 4) * H + h) * W + w0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 g4 = *reinterpret_cast<const float4*>(gp + (size_t)j * H * W);
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
        if (ci == 0) ab[j] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) aw[j][rr * 3 + c] = fmaf(gv[i], v[rr][c + i], aw[j][rr * 3 + c]);
      }
    }
    // lanes -> lane 0, then the block's warps of one channel group -> one atomicAdd per entry
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float s = aw[j][k];
#pragma unroll
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        aw[j][k] = s;
      }
      float s = ab[j];
#pragma unroll
      for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      ab[j] = s;
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < 9; ++k) red[warp][j * 9 + k] = aw[j][k];
        red[warp][36 + j] = ab[j];
      }
    }
    __syncthreads();
    // warps w and w + OG hold the same channel group; a warp that never ran holds zeros
    if (warp < OG) {
      for (int e = lane; e < 40; e += 32) {
        float s = 0.f;
        for (int w2 = warp; w2 < kWarps; w2 += OG) s += red[w2][e];
        if (e < 36) {
          const int j = e / 9, k = e - j * 9;
          atomicAdd(gw + ((size_t)(og * 4 + j) * Cin + ci) * 9 + k, s);
        } else if (ci == 0) {
          atomicAdd(gb + og * 4 + (e - 36), s);
        }
      }
    }
  }
}

}  // namespace stem

static int stem_check(int B, int Cin, int O, int H, int W) {
  if (B <= 0 || Cin <= 0 || O <= 0 || H <= 0 || W <= 0) {
    set_error("stem conv: bad extents B=%d Cin=%d O=%d H=%d W=%d", B, Cin, O, H, W);
    return DCN_ERR_BAD_SHAPE;
  }
  if (Cin > stem::kMaxCin || (O != 16 && O != 32) || (W & 3)) {
    set_error("stem conv: needs Cin <= %d, O in {16, 32}, W %% 4 == 0 (got Cin=%d O=%d W=%d)", stem::kMaxCin, Cin, O, W);
    return DCN_ERR_UNSUPPORTED;
  }
  return DCN_OK;
}

}  // namespace dcn

using namespace dcn;

extern "C" {

int dcn_stem_conv_forward(int B, int Cin, int O, int H, int W, const void* x, const void* weight, const void* bias,
                          void* out, void* stream) {
  int rc = stem_check(B, Cin, O, H, W);
  if (rc) return rc;
  if (!x || !weight || !out) return DCN_ERR_NULL_POINTER;
  if (((uintptr_t)x | (uintptr_t)out) & 15) return DCN_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * H * (W >> 2);
  const int blocks = (int)std::min<long long>((total + 127) / 128, (long long)148 * 16);
  KernelScope scope("stem_conv_fwd_kernel", st);
  if (O == 16)
    stem::stem_fwd_kernel<16><<<blocks, 128, 0, st>>>(B, Cin, H, W, (const float*)x, (const float*)weight,
                                                      (const float*)bias, (float*)out);
  else
    stem::stem_fwd_kernel<32><<<blocks, 128, 0, st>>>(B, Cin, H, W, (const float*)x, (const float*)weight,
                                                      (const float*)bias, (float*)out);
  DCN_KERNEL_CHECK("stem_conv_fwd_kernel");
  return DCN_OK;
}

int dcn_stem_conv_backward(int B, int Cin, int O, int H, int W, const void* x, const void* grad_out, void* grad_weight,
                           void* grad_bias, void* stream) {
  int rc = stem_check(B, Cin, O, H, W);
  if (rc) return rc;
  if (!x || !grad_out || !grad_weight || !grad_bias) return DCN_ERR_NULL_POINTER;
  if (((uintptr_t)x | (uintptr_t)grad_out) & 15) return DCN_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  DCN_CUDA_TRY(cudaMemsetAsync(grad_weight, 0, sizeof(float) * (size_t)O * Cin * 9, st));
  DCN_CUDA_TRY(cudaMemsetAsync(grad_bias, 0, sizeof(float) * (size_t)O, st));
  // gridDim.x * 8 warps must be a multiple of the channel groups (4 or 8): any block count will do
  const long long units = (long long)B * H * ((W / 4 + 31) / 32) * (O / 4);
  const int blocks = (int)std::min<long long>((units + 7) / 8, (long long)148 * 4);
  KernelScope scope("stem_conv_bwd_kernel", st);
  if (O == 16)
    stem::stem_bwd_kernel<16><<<blocks, 256, 0, st>>>(B, Cin, H, W, (const float*)x, (const float*)grad_out,
                                                      (float*)grad_weight, (float*)grad_bias);
  else
    stem::stem_bwd_kernel<32><<<blocks, 256, 0, st>>>(B, Cin, H, W, (const float*)x, (const float*)grad_out,
                                                      (float*)grad_weight, (float*)grad_bias);
  DCN_KERNEL_CHECK("stem_conv_bwd_kernel");
  return DCN_OK;
}

}  // extern "C"
