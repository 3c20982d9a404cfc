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
#include "dcn_common.cuh"

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

}  // namespace dcn
