// dcn_bn.cu — the post-op that follows every DeformConv2d layer of the reference's detector:
// BatchNorm2d + ReLU (train.py:146-159 modules, train.py:167-170 / 329-332 call sites), training and
// eval mode, forward and backward, NCHW float32.  SURVEY.md 8(f) rank 2.
//
// Why it lives here: the framework's batch-norm kernels (cuDNN bn_fw_tr_1C11 / bn_bw_1C11, ATen
// batch_norm_collect_statistics / batch_norm_backward) run ONE CTA PER CHANNEL.  The detector has 16-256
// channels with up to 16.7 M elements each, so those kernels leave most of the 148 SMs idle: measured
// on B200 at batch 1024 they take 49.6 ms of the 82 ms training step (profiles/r1_detector_profile.txt).
// These kernels split every channel over many CTAs and are plain HBM streaming:
//   forward   bn_stats (read x) -> bn_finalize (C threads) -> bn_apply_relu (read x, write y)      3 passes
//   backward  bn_bwd_reduce (read x, dy) -> bn_bwd_finalize -> bn_bwd_apply (read x, dy, write dx)  5 passes
// ReLU is fused on both sides: the backward mask is recomputed from x with the SAME fmaf(x, scale, shift)
// the forward pass evaluated, so y is never re-read and the mask is bit-identical.
#include "dcn_umma_common.cuh"

namespace dcn {

namespace bn {

constexpr int kThreads = 256;

// channel of flat NCHW element index i (hw = plane size); planes are contiguous
__device__ __forceinline__ int chan_of(size_t plane, int C) { return (int)(plane % (size_t)C); }

// block reduction of two doubles, result valid in thread 0
__device__ __forceinline__ void block_reduce2(double& a, double& b) {
  __shared__ double ra[kThreads / 32], rb[kThreads / 32];
  for (int d = 16; d; d >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, d);
    b += __shfl_xor_sync(0xffffffffu, b, d);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    ra[w] = a;
    rb[w] = b;
  }
  __syncthreads();
  if (w == 0) {
    a = l < kThreads / 32 ? ra[l] : 0.0;
    b = l < kThreads / 32 ? rb[l] : 0.0;
    for (int d = 4; d; d >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, d);
      b += __shfl_xor_sync(0xffffffffu, b, d);
    }
  }
}

// Block (c, s): batch elements [s * bpb, (s + 1) * bpb) of channel c, walked as one flat index space so that
// small planes (8 x 8 in the detector's last layer) keep all threads busy.  fp32 partial sums of at most 64
// elements are flushed into double accumulators.
//   MODE 0: sums[c] += {sum (x - pivot), sum (x - pivot)^2} with pivot = the channel's first element x[0, c, 0]:
//           shifted sums, so that the variance S2/n - (S1/n)^2 does not cancel when |mean| >> std
//           (a channel at mean 1e3, std 1e-2 loses every digit of E[x^2] - E[x]^2 in fp32)
//   MODE 1: sums[c] += {sum dyr, sum dyr * (x - mean[c])}   with dyr = dy * [fmaf(x, scale, shift) > 0]
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kThreads) bn_reduce_kernel(int B, int C, int HW, int bpb,
                                                             const float* __restrict__ x,
                                                             const float* __restrict__ dy,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             const float* __restrict__ mean,
                                                             double* __restrict__ sums) {
  const int c = blockIdx.x, b0 = blockIdx.y * bpb, b1 = min(B, b0 + bpb);
  const int per_b = VEC ? (HW >> 2) : HW, total = (b1 - b0) * per_b;
  const size_t b_stride = (size_t)C * HW;
  const float* xb = x + ((size_t)b0 * C + c) * HW;
  const float* db = MODE == 1 ? dy + ((size_t)b0 * C + c) * HW : nullptr;
  float sc = 0.f, sh = 0.f, mu = 0.f;
  if (MODE == 1) {
    sc = scale[c];
    sh = shift[c];
    mu = mean[c];
  } else {
    mu = __ldg(x + (size_t)c * HW);  // pivot
  }
  double A = 0.0, Q = 0.0;
  float a = 0.f, q = 0.f;
  int pending = 0;
  for (int i = threadIdx.x; i < total; i += kThreads) {
    const int b = i / per_b, r = i - b * per_b;
    if (VEC) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)b * b_stride) + r);
      if (MODE == 0) {
        const float d0 = v.x - mu, d1 = v.y - mu, d2 = v.z - mu, d3 = v.w - mu;
        a += (d0 + d1) + (d2 + d3);
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      } else {
        const float4 g = __ldg(reinterpret_cast<const float4*>(db + (size_t)b * b_stride) + r);
        const float g0 = fmaf(v.x, sc, sh) > 0.f ? g.x : 0.f, g1 = fmaf(v.y, sc, sh) > 0.f ? g.y : 0.f;
        const float g2 = fmaf(v.z, sc, sh) > 0.f ? g.z : 0.f, g3 = fmaf(v.w, sc, sh) > 0.f ? g.w : 0.f;
        a += (g0 + g1) + (g2 + g3);
        q += (g0 * (v.x - mu) + g1 * (v.y - mu)) + (g2 * (v.z - mu) + g3 * (v.w - mu));
      }
    } else {
      const float v = __ldg(xb + (size_t)b * b_stride + r);
      if (MODE == 0) {
        a += v - mu;
        q += (v - mu) * (v - mu);
      } else {
        const float g = fmaf(v, sc, sh) > 0.f ? __ldg(db + (size_t)b * b_stride + r) : 0.f;
        a += g;
        q += g * (v - mu);
      }
    }
    if (++pending == 16) {
      A += (double)a;
      Q += (double)q;
      a = q = 0.f;
      pending = 0;
    }
  }
  A += (double)a;
  Q += (double)q;
  block_reduce2(A, Q);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, A);
    atomicAdd(sums + 2 * c + 1, Q);
  }
}

// per channel: batch statistics -> affine map of the forward pass, saved statistics, running statistics
// (nn.BatchNorm2d: biased variance normalises, unbiased variance updates running_var)
__global__ void bn_finalize_kernel(int C, int HW, const float* __restrict__ x, double count,
                                   const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd, float* __restrict__ scale,
                                   float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // shifted sums about the pivot x[0, c, 0] (bn_reduce_kernel MODE 0)
  const double dm = sums[2 * c] / count;
  const double m = (double)x[(size_t)c * HW] + dm;
  double var = sums[2 * c + 1] / count - dm * dm;
  if (var < 0.0) var = 0.0;
  const float mean = (float)m, invstd = 1.0f / sqrtf((float)var + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  save_mean[c] = mean;
  save_invstd[c] = invstd;
  scale[c] = g * invstd;
  shift[c] = fmaf(-mean, g * invstd, b);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: affine map from the running statistics
__global__ void bn_eval_affine_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean,
                                      const float* __restrict__ running_var, float eps,
                                      float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.0f / sqrtf(running_var[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  save_mean[c] = running_mean[c];
  save_invstd[c] = invstd;
  scale[c] = g * invstd;
  shift[c] = fmaf(-running_mean[c], g * invstd, b);
}

// y = max(0, fmaf(x, scale[c], shift[c]))   (RELU = false: no clamp)
template <bool VEC, bool RELU>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(size_t n_items, int C, int HW,
                                                            const float* __restrict__ x,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            float* __restrict__ y) {
  const int per_plane = VEC ? (HW >> 2) : HW;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n_items; i += (size_t)gridDim.x * kThreads) {
    const int c = chan_of(i / per_plane, C);
    const float sc = __ldg(scale + c), sh = __ldg(shift + c);
    if (VEC) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      v.x = fmaf(v.x, sc, sh);
      v.y = fmaf(v.y, sc, sh);
      v.z = fmaf(v.z, sc, sh);
      v.w = fmaf(v.w, sc, sh);
      if (RELU) {
        v.x = fmaxf(v.x, 0.f);
        v.y = fmaxf(v.y, 0.f);
        v.z = fmaxf(v.z, 0.f);
        v.w = fmaxf(v.w, 0.f);
      }
      reinterpret_cast<float4*>(y)[i] = v;
    } else {
      const float v = fmaf(__ldg(x + i), sc, sh);
      y[i] = RELU ? fmaxf(v, 0.f) : v;
    }
  }
}

// per channel: grad_gamma = invstd * sum dyr (x - mean), grad_beta = sum dyr, and the coefficients of
//   dx = a * (dyr - k1 - (x - mean) * k2),  a = gamma * invstd, k1 = grad_beta / M, k2 = invstd^2 * sum2 / M
// (eval mode: the statistics are constants, dx = a * dyr: k1 = k2 = 0)
__global__ void bn_bwd_finalize_kernel(int C, double count, int training, const double* __restrict__ sums,
                                       const float* __restrict__ save_invstd, float* __restrict__ grad_gamma,
                                       float* __restrict__ grad_beta, float* __restrict__ k1,
                                       float* __restrict__ k2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s1 = sums[2 * c], s2 = sums[2 * c + 1], is = (double)save_invstd[c];
  if (grad_gamma) grad_gamma[c] = (float)(s2 * is);
  if (grad_beta) grad_beta[c] = (float)s1;
  k1[c] = training ? (float)(s1 / count) : 0.f;
  k2[c] = training ? (float)(s2 * is * is / count) : 0.f;
}

template <bool VEC, bool RELU>
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(size_t n_items, int C, int HW,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ dy,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ k1,
                                                                const float* __restrict__ k2,
                                                                float* __restrict__ dx) {
  const int per_plane = VEC ? (HW >> 2) : HW;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n_items; i += (size_t)gridDim.x * kThreads) {
    const int c = chan_of(i / per_plane, C);
    const float sc = __ldg(scale + c), sh = __ldg(shift + c), mu = __ldg(mean + c);
    const float a1 = __ldg(k1 + c), a2 = __ldg(k2 + c);
    if (VEC) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy) + i);
      float4 r;
      r.x = sc * ((!RELU || fmaf(v.x, sc, sh) > 0.f ? g.x : 0.f) - a1 - (v.x - mu) * a2);
      r.y = sc * ((!RELU || fmaf(v.y, sc, sh) > 0.f ? g.y : 0.f) - a1 - (v.y - mu) * a2);
      r.z = sc * ((!RELU || fmaf(v.z, sc, sh) > 0.f ? g.z : 0.f) - a1 - (v.z - mu) * a2);
      r.w = sc * ((!RELU || fmaf(v.w, sc, sh) > 0.f ? g.w : 0.f) - a1 - (v.w - mu) * a2);
      reinterpret_cast<float4*>(dx)[i] = r;
    } else {
      const float v = __ldg(x + i);
      const float g = (!RELU || fmaf(v, sc, sh) > 0.f) ? __ldg(dy + i) : 0.f;
      dx[i] = sc * (g - a1 - (v - mu) * a2);
    }
  }
}

static int stream_grid(size_t n_items) {
  const size_t blocks = (n_items + kThreads - 1) / kThreads;
  const size_t cap = 148 * 16;  // 16 resident CTAs of 256 threads per SM's worth of loads in flight
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

// batch slices per channel: enough CTAs to fill the machine (~8 per SM), at least one batch element each
static void reduce_grid(int B, int C, int* slices, int* bpb) {
  int s = (148 * 8 + C - 1) / C;
  if (s > B) s = B;
  if (s < 1) s = 1;
  *bpb = (B + s - 1) / s;
  *slices = (B + *bpb - 1) / *bpb;
}

}  // namespace bn

// scratch: [sums: 2C doubles][k1: C floats][k2: C floats]
size_t bn_workspace_bytes(int C) { return align_up(sizeof(double) * 2 * (size_t)C + sizeof(float) * 2 * (size_t)C, 256); }

// saved: 4C floats [batch mean | 1/sqrt(var + eps) | scale = gamma * invstd | shift = beta - mean * scale],
// written by the forward pass, read by the backward pass (eval mode: the running statistics' values)
int bn_relu_forward(int B, int C, int HW, int training, const float* x, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, float* y, float* saved,
                    void* workspace, cudaStream_t st) {
  using namespace bn;
  const bool vec = (HW & 3) == 0;
  const size_t n = (size_t)B * C * HW, n_items = vec ? n / 4 : n;
  float *save_mean = saved, *save_invstd = saved + C, *scale = saved + 2 * C, *shift = saved + 3 * C;
  if (training) {
    double* sums = (double*)workspace;
    DCN_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st));
    int slices, bpb;
    reduce_grid(B, C, &slices, &bpb);
    {
      KernelScope scope("bn_stats_kernel", st);
      if (vec)
        bn_reduce_kernel<0, true><<<dim3(C, slices), kThreads, 0, st>>>(B, C, HW, bpb, x, nullptr, nullptr, nullptr,
                                                                     nullptr, sums);
      else
        bn_reduce_kernel<0, false><<<dim3(C, slices), kThreads, 0, st>>>(B, C, HW, bpb, x, nullptr, nullptr, nullptr,
                                                                      nullptr, sums);
      DCN_KERNEL_CHECK("bn_stats_kernel");
    }
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, HW, x, (double)B * HW, sums, gamma, beta, eps, momentum,
                                                      running_mean, running_var, save_mean, save_invstd, scale,
                                                      shift);
    DCN_KERNEL_CHECK("bn_finalize_kernel");
  } else {
    bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, gamma, beta, running_mean, running_var, eps,
                                                         save_mean, save_invstd, scale, shift);
    DCN_KERNEL_CHECK("bn_eval_affine_kernel");
  }
  KernelScope scope("bn_apply_relu_kernel", st);
  const int grid = stream_grid(n_items);
  if (vec) bn_apply_kernel<true, true><<<grid, kThreads, 0, st>>>(n_items, C, HW, x, scale, shift, y);
  else bn_apply_kernel<false, true><<<grid, kThreads, 0, st>>>(n_items, C, HW, x, scale, shift, y);
  DCN_KERNEL_CHECK("bn_apply_relu_kernel");
  return DCN_OK;
}

int bn_relu_backward(int B, int C, int HW, int training, const float* x, const float* grad_y, const float* saved,
                     float* grad_x, float* grad_gamma, float* grad_beta, void* workspace, cudaStream_t st) {
  using namespace bn;
  const bool vec = (HW & 3) == 0;
  const size_t n = (size_t)B * C * HW, n_items = vec ? n / 4 : n;
  const float *save_mean = saved, *save_invstd = saved + C, *scale = saved + 2 * C, *shift = saved + 3 * C;
  double* sums = (double*)workspace;
  float* k1 = (float*)(sums + 2 * (size_t)C);
  float* k2 = k1 + C;
  DCN_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st));
  int slices, bpb;
  reduce_grid(B, C, &slices, &bpb);
  {
    KernelScope scope("bn_bwd_reduce_kernel", st);
    if (vec)
      bn_reduce_kernel<1, true><<<dim3(C, slices), kThreads, 0, st>>>(B, C, HW, bpb, x, grad_y, scale, shift,
                                                                   save_mean, sums);
    else
      bn_reduce_kernel<1, false><<<dim3(C, slices), kThreads, 0, st>>>(B, C, HW, bpb, x, grad_y, scale, shift,
                                                                    save_mean, sums);
    DCN_KERNEL_CHECK("bn_bwd_reduce_kernel");
  }
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, (double)B * HW, training, sums, save_invstd, grad_gamma,
                                                        grad_beta, k1, k2);
  DCN_KERNEL_CHECK("bn_bwd_finalize_kernel");
  if (grad_x) {
    KernelScope scope("bn_bwd_apply_kernel", st);
    const int grid = stream_grid(n_items);
    if (vec)
      bn_bwd_apply_kernel<true, true><<<grid, kThreads, 0, st>>>(n_items, C, HW, x, grad_y, scale, shift, save_mean,
                                                               k1, k2, grad_x);
    else
      bn_bwd_apply_kernel<false, true><<<grid, kThreads, 0, st>>>(n_items, C, HW, x, grad_y, scale, shift, save_mean,
                                                                k1, k2, grad_x);
    DCN_KERNEL_CHECK("bn_bwd_apply_kernel");
  }
  return DCN_OK;
}


// ---- channels-last hand-over between stacked engine layers in TRAINING -----------------------------
// The post-op sits between two DCN layers: its input is the producer's NCHW output (batch statistics need the complete
// tensor), its output is only ever read by the consumer layer — as that layer's framed channels-last staging copy.
// So the normalise + ReLU pass writes that copy directly (it IS the staging transposition, with the affine map applied
// in registers), and the backward pass reads the consumer's channels-last grad_x accumulator directly (the un-staging
// transposition with the BatchNorm backward applied in registers).  Per layer boundary this removes two passes over
// the activation in each direction (bn_apply + nchw_to_nhwc -> one kernel; nhwc_to_nchw + bn_bwd_apply -> one kernel).
namespace bn {

__device__ __forceinline__ size_t frame_px_of(const Geo& g, const FastDiv& divW, int p) {
  uint32_t y, xx;
  divW.divmod((uint32_t)p, y, xx);
  return (size_t)((y + 1) * (g.W + 2) + xx + 1) * g.C;
}

// Register tile of the transposing kernels: a thread owns 4 staged channels x 4 pixels.  Lanes: CGL along the channels
// (4 channels each) x 32 / CGL along the pixels; CGL = 8 covers 32 channels per warp row, CGL = 4 is for 16-channel
// tensors (the detector's first DCN layer), where the 8-lane layout would leave half of every warp idle.
template <int CGL>
struct TileMap {
  static constexpr int kPixPerWarp = 4 * (32 / CGL), kPixPerBlock = 8 * kPixPerWarp, kChPerRow = 4 * CGL;
  __device__ static int chan(int by, int lane) { return by * kChPerRow + (lane % CGL) * 4; }
  __device__ static int pix(int tile, int warp, int lane) { return tile * kPixPerBlock + warp * kPixPerWarp + (lane / CGL) * 4; }
};

// xt[b, frame(p), perm(c)] = max(0, x[b, c, p] * scale[c] + shift[c]); tile structure of nchw_to_nhwc_kernel
template <int CGL>
__global__ void __launch_bounds__(256) bn_stage_kernel(Geo g, int variant, int G, int Cs, FastDiv divW,
                                                       const float* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, float* __restrict__ xt) {
  const int HWi = g.H * g.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = TileMap<CGL>::chan(blockIdx.y, lane);            // destination (staged) channels d..d+3
  const int p = TileMap<CGL>::pix(blockIdx.x, warp, lane);       // pixels p..p+3
  if (d >= g.C || p >= HWi) return;
  const int b = blockIdx.z;
  const bool vec_ok = (HWi & 3) == 0;
  float v[4][4];   // [channel][pixel]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dd = d + k;
    const int c = variant == DCN_VARIANT_TORCH ? (dd % G) * Cs + dd / G : dd;   // source channel of staged channel dd
    const float* src = x + ((size_t)b * g.C + c) * HWi + p;
    const float sc = __ldg(scale + c), sh = __ldg(shift + c);
    if (vec_ok) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src));
      v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[k][i] = p + i < HWi ? __ldg(src + i) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[k][i] = fmaxf(fmaf(v[k][i], sc, sh), 0.f);
  }
  float* img = xt + (size_t)b * xt_image_stride(g) + d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (p + i >= HWi) break;
    *reinterpret_cast<float4*>(img + frame_px_of(g, divW, p + i)) = make_float4(v[0][i], v[1][i], v[2][i], v[3][i]);
  }
}

// sums[c] += {sum dyr, sum dyr * (x - mean[c])}, dyr = gxt[b, frame(p), perm(c)] * [fmaf(x, scale, shift) > 0].
// Block (slice, channel row): walks its share of the (image, pixel block) tiles with the register tile of the
// transposing kernels and reduces once at the end.
template <int CGL>
__global__ void __launch_bounds__(256) bn_bwd_reduce_cl_kernel(Geo g, int variant, int G, int Cs, FastDiv divW, int pblocks,
                                                               int tiles_per_slice, const float* __restrict__ x,
                                                               const float* __restrict__ gxt,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const float* __restrict__ mean,
                                                               double* __restrict__ sums) {
  __shared__ double red[8][CGL][4][2];
  const int HWi = g.H * g.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = TileMap<CGL>::chan(blockIdx.y, lane);
  const bool live = d < g.C;
  int c[4];
  float sc[4], sh[4], mu[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dd = live ? d + k : 0;
    c[k] = variant == DCN_VARIANT_TORCH ? (dd % G) * Cs + dd / G : dd;
    sc[k] = scale[c[k]];
    sh[k] = shift[c[k]];
    mu[k] = mean[c[k]];
  }
  const bool vec_ok = (HWi & 3) == 0;
  double A[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
  float af[4] = {0.f, 0.f, 0.f, 0.f}, qf[4] = {0.f, 0.f, 0.f, 0.f};   // fp32 partial sums of <= 32 elements, then double
  int pending = 0;
  const int total = g.B * pblocks, t0 = blockIdx.x * tiles_per_slice, t1 = min(total, t0 + tiles_per_slice);
  // two tiles per iteration: all 16 loads of both tiles are issued before the first value is used (one tile per
  // iteration ran at half the rate of the NCHW reduction: 8 loads in flight per thread)
  for (int tile = t0; tile < t1; tile += 2) {
    float gv[2][4][4], xv[2][4][4];   // [tile][pixel][channel], [tile][channel][pixel]
    int pp[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tl = tile + u;
      const int b = tl < t1 ? tl / pblocks : 0;
      const int p = tl < t1 ? TileMap<CGL>::pix(tl - b * pblocks, warp, lane) : HWi;
      pp[u] = (live && p < HWi) ? p : HWi;
      if (pp[u] >= HWi) continue;
      const float* img = gxt + (size_t)b * xt_image_stride(g) + d;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p + i < HWi) t = __ldg(reinterpret_cast<const float4*>(img + frame_px_of(g, divW, p + i)));
        gv[u][i][0] = t.x; gv[u][i][1] = t.y; gv[u][i][2] = t.z; gv[u][i][3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* src = x + ((size_t)b * g.C + c[k]) * HWi + p;
        if (vec_ok) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src));
          xv[u][k][0] = t.x; xv[u][k][1] = t.y; xv[u][k][2] = t.z; xv[u][k][3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[u][k][i] = p + i < HWi ? __ldg(src + i) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (pp[u] >= HWi) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float gr = (pp[u] + i < HWi && fmaf(xv[u][k][i], sc[k], sh[k]) > 0.f) ? gv[u][i][k] : 0.f;
          a += gr;
          q += gr * (xv[u][k][i] - mu[k]);
        }
        af[k] += a;
        qf[k] += q;
      }
    }
    if (++pending == 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        A[k] += (double)af[k];
        Q[k] += (double)qf[k];
        af[k] = qf[k] = 0.f;
      }
      pending = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    A[k] += (double)af[k];
    Q[k] += (double)qf[k];
  }
  // lanes with equal (lane % CGL) hold the same four channels
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = CGL; o < 32; o <<= 1) {
      A[k] += __shfl_xor_sync(0xffffffffu, A[k], o);
      Q[k] += __shfl_xor_sync(0xffffffffu, Q[k], o);
    }
  }
  if (lane < CGL) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      red[warp][lane][k][0] = A[k];
      red[warp][lane][k][1] = Q[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 4 * CGL) {
    const int grp = threadIdx.x >> 2, k = threadIdx.x & 3;
    double a = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += red[w][grp][k][0];
      q += red[w][grp][k][1];
    }
    const int dd = blockIdx.y * TileMap<CGL>::kChPerRow + grp * 4 + k;
    if (dd < g.C) {
      const int cc = variant == DCN_VARIANT_TORCH ? (dd % G) * Cs + dd / G : dd;
      atomicAdd(sums + 2 * cc, a);
      atomicAdd(sums + 2 * cc + 1, q);
    }
  }
}

// dx[b, c, p] = scale[c] * (dyr - k1[c] - (x - mean[c]) * k2[c]); tile structure of nhwc_to_nchw_kernel
template <int CGL>
__global__ void __launch_bounds__(256) bn_bwd_apply_cl_kernel(Geo g, int variant, int G, int Cs, FastDiv divW,
                                                              const float* __restrict__ x, const float* __restrict__ gxt,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const float* __restrict__ mean, const float* __restrict__ k1,
                                                              const float* __restrict__ k2, float* __restrict__ dx) {
  const int HWi = g.H * g.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = TileMap<CGL>::chan(blockIdx.y, lane);
  const int p = TileMap<CGL>::pix(blockIdx.x, warp, lane);
  if (d >= g.C || p >= HWi) return;
  const int b = blockIdx.z;
  const float* img = gxt + (size_t)b * xt_image_stride(g) + d;
  float gv[4][4];   // [pixel][channel]
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p + i < HWi) t = __ldg(reinterpret_cast<const float4*>(img + frame_px_of(g, divW, p + i)));
    gv[i][0] = t.x; gv[i][1] = t.y; gv[i][2] = t.z; gv[i][3] = t.w;
  }
  const bool vec_ok = (HWi & 3) == 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dd = d + k;
    const int c = variant == DCN_VARIANT_TORCH ? (dd % G) * Cs + dd / G : dd;
    const float sc = __ldg(scale + c), sh = __ldg(shift + c), mu = __ldg(mean + c), a1 = __ldg(k1 + c), a2 = __ldg(k2 + c);
    const size_t o = ((size_t)b * g.C + c) * HWi + p;
    float xv[4], r[4];
    if (vec_ok) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + o));
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = p + i < HWi ? __ldg(x + o + i) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = sc * ((fmaf(xv[i], sc, sh) > 0.f ? gv[i][k] : 0.f) - a1 - (xv[i] - mu) * a2);
    if (vec_ok) {
      *reinterpret_cast<float4*>(dx + o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p + i < HWi) dx[o + i] = r[i];
    }
  }
}

}  // namespace bn

// g = the CONSUMER layer's geometry (its input is this post-op's output), t = its staging layout
int bn_relu_forward_staged(const Geo& g, const Tiling& t, int training, const float* x, const float* gamma,
                           const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                           float* xt, float* saved, void* workspace, cudaStream_t st) {
  using namespace bn;
  const int B = g.B, C = g.C, HW = g.H * g.W;
  const bool vec = (HW & 3) == 0;
  float *save_mean = saved, *save_invstd = saved + C, *scale = saved + 2 * C, *shift = saved + 3 * C;
  if (training) {
    double* sums = (double*)workspace;
    DCN_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st));
    int slices, bpb;
    reduce_grid(B, C, &slices, &bpb);
    {
      KernelScope scope("bn_stats_kernel", st);
      if (vec)
        bn_reduce_kernel<0, true><<<dim3(C, slices), kThreads, 0, st>>>(B, C, HW, bpb, x, nullptr, nullptr, nullptr,
                                                                     nullptr, sums);
      else
        bn_reduce_kernel<0, false><<<dim3(C, slices), kThreads, 0, st>>>(B, C, HW, bpb, x, nullptr, nullptr, nullptr,
                                                                      nullptr, sums);
      DCN_KERNEL_CHECK("bn_stats_kernel");
    }
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, HW, x, (double)B * HW, sums, gamma, beta, eps, momentum,
                                                      running_mean, running_var, save_mean, save_invstd, scale,
                                                      shift);
    DCN_KERNEL_CHECK("bn_finalize_kernel");
  } else {
    bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, gamma, beta, running_mean, running_var, eps,
                                                         save_mean, save_invstd, scale, shift);
    DCN_KERNEL_CHECK("bn_eval_affine_kernel");
  }
  // frame of the staged copy (the workspace belongs to the caller: re-zeroed on every call, as the plain staging does)
  int rc = launch_xt_frame_zero(g, xt, st);
  if (rc) return rc;
  KernelScope scope("bn_relu_stage_kernel", st);
  if (C <= 16) {
    const dim3 grid((HW + 255) / 256, (C + 15) / 16, B);
    bn_stage_kernel<4><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, FastDiv::make(g.W), x, scale, shift, xt);
  } else {
    const dim3 grid((HW + 127) / 128, (C + 31) / 32, B);
    bn_stage_kernel<8><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, FastDiv::make(g.W), x, scale, shift, xt);
  }
  DCN_KERNEL_CHECK("bn_relu_stage_kernel");
  return DCN_OK;
}

int bn_relu_backward_staged(const Geo& g, const Tiling& t, int training, const float* x, const float* gxt,
                            const float* saved, float* grad_x, float* grad_gamma, float* grad_beta, void* workspace,
                            cudaStream_t st) {
  using namespace bn;
  const int B = g.B, C = g.C, HW = g.H * g.W;
  const float *save_mean = saved, *save_invstd = saved + C, *scale = saved + 2 * C, *shift = saved + 3 * C;
  double* sums = (double*)workspace;
  float* k1 = (float*)(sums + 2 * (size_t)C);
  float* k2 = k1 + C;
  DCN_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st));
  const bool narrow = C <= 16;
  const int pblocks = narrow ? (HW + 255) / 256 : (HW + 127) / 128, cgroups = narrow ? (C + 15) / 16 : (C + 31) / 32;
  const int total = B * pblocks;
  int slices = (148 * 8 + cgroups - 1) / cgroups;
  if (slices > total) slices = total;
  const int tps = (total + slices - 1) / slices;
  slices = (total + tps - 1) / tps;
  {
    KernelScope scope("bn_bwd_reduce_kernel", st);
    if (narrow)
      bn_bwd_reduce_cl_kernel<4><<<dim3(slices, cgroups), 256, 0, st>>>(g, t.variant, t.G, t.Cs, FastDiv::make(g.W), pblocks,
                                                                      tps, x, gxt, scale, shift, save_mean, sums);
    else
      bn_bwd_reduce_cl_kernel<8><<<dim3(slices, cgroups), 256, 0, st>>>(g, t.variant, t.G, t.Cs, FastDiv::make(g.W), pblocks,
                                                                      tps, x, gxt, scale, shift, save_mean, sums);
    DCN_KERNEL_CHECK("bn_bwd_reduce_kernel");
  }
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, (double)B * HW, training, sums, save_invstd, grad_gamma,
                                                        grad_beta, k1, k2);
  DCN_KERNEL_CHECK("bn_bwd_finalize_kernel");
  if (grad_x) {
    KernelScope scope("bn_bwd_unstage_kernel", st);
    const dim3 grid(pblocks, cgroups, B);
    if (narrow)
      bn_bwd_apply_cl_kernel<4><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, FastDiv::make(g.W), x, gxt, scale, shift,
                                                     save_mean, k1, k2, grad_x);
    else
      bn_bwd_apply_cl_kernel<8><<<grid, 256, 0, st>>>(g, t.variant, t.G, t.Cs, FastDiv::make(g.W), x, gxt, scale, shift,
                                                     save_mean, k1, k2, grad_x);
    DCN_KERNEL_CHECK("bn_bwd_unstage_kernel");
  }
  return DCN_OK;
}

}  // namespace dcn
