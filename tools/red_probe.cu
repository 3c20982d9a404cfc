// red_probe.cu — how fast can an SM add fp32 rows into an L2-resident image?  (round 2)
//
// The fused backward's scatter adds, per sample, 4 corner rows of C floats (nw|ne and sw|se are each ONE contiguous
// run of 2*C floats in the framed channels-last gradient image).  Variants measured, all doing the same logical work
// per warp iteration (two runs of ROW bytes each at random ROW-aligned spots of a `region`):
//   mode 0  red.global.add.v2.f32, lanes cover the run (what the kernel does today; ROW = 512 -> 2 instr per run)
//   mode 1  red.global.add.v4.f32
//   mode 2  products written to shared memory (STS.64), fence.proxy.async, then ONE lane issues
//           cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 per run (the TMA engine does the adds:
//           no LSU / L1 data-pipe wavefront per sector)
//   mode 3  as 2 but without the STS (buffer written once): the bulk-reduce issue rate alone
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/red_probe tools/red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kWarps = 16;
constexpr int kRing = 4;  // staging buffers per warp

template <int ROW>  // bytes of one contiguous run (256, 512, 1024)
__global__ void __launch_bounds__(kWarps * 32, 1) probe(float* buf, uint32_t region_bytes_mask, int iters, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * kWarps + warp;
  uint8_t* stage = smem + (size_t)warp * kRing * 2 * ROW;  // [ring][2 runs][ROW]
  if (mode == 3) {
    for (int i = lane * 4; i < kRing * 2 * ROW; i += 128) *reinterpret_cast<float*>(stage + i) = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  for (int it = 0; it < iters; ++it) {
    const uint32_t a0 = (hash32(gw * 7919u + 2 * it) & region_bytes_mask) & ~(uint32_t)(ROW - 1);
    const uint32_t a1 = (hash32(gw * 7919u + 2 * it + 1) & region_bytes_mask) & ~(uint32_t)(ROW - 1);
    char* p0 = reinterpret_cast<char*>(buf) + a0;
    char* p1 = reinterpret_cast<char*>(buf) + a1;
    const float v = (float)(it & 3);
    if (mode == 0) {
#pragma unroll
      for (int k = 0; k < ROW / 256; ++k) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p0 + k * 256 + lane * 8), "f"(v), "f"(v) : "memory");
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p1 + k * 256 + lane * 8), "f"(v), "f"(v) : "memory");
      }
    } else if (mode == 1) {
#pragma unroll
      for (int k = 0; k < (ROW + 511) / 512; ++k) {
        if (ROW >= 512 || lane < 16) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p0 + k * 512 + lane * 16), "f"(v), "f"(v),
                       "f"(v), "f"(v) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p1 + k * 512 + lane * 16), "f"(v), "f"(v),
                       "f"(v), "f"(v) : "memory");
        }
      }
    } else {
      uint8_t* sb = stage + (size_t)(it % kRing) * 2 * ROW;
      // the bulk reduce that last read this buffer must have finished reading it
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kRing - 1) : "memory");
      __syncwarp();
      if (mode == 2) {
#pragma unroll
        for (int k = 0; k < 2 * ROW / 256; ++k)
          *reinterpret_cast<float2*>(sb + k * 256 + lane * 8) = make_float2(v, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
      }
      if (lane == 0) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(p0),
                     "r"(smem_u32(sb)), "n"(ROW) : "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(p1),
                     "r"(smem_u32(sb + ROW)), "n"(ROW) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (mode >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int ROW>
static int run(float* buf, size_t region_bytes, int mode, const char* name) {
  const int blocks = 148, iters = 4000;
  const size_t smem = (size_t)kWarps * kRing * 2 * ROW;
  CK(cudaFuncSetAttribute(probe<ROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  probe<ROW><<<blocks, kWarps * 32, smem>>>(buf, (uint32_t)(region_bytes - 1), iters, mode);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  probe<ROW><<<blocks, kWarps * 32, smem>>>(buf, (uint32_t)(region_bytes - 1), iters, mode);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)blocks * kWarps * iters * 2 * ROW;
  printf("ROW=%4d region=%4zu MB mode %d %-34s: %7.3f ms  %7.1f G adds/s  %6.1f G sectors/s  %5.2f sectors/clk/SM\n", ROW,
         region_bytes >> 20, mode, name, ms, bytes / 4 / ms * 1e-6, bytes / 32 / ms * 1e-6,
         bytes / 32 / 148 / (ms * 1e-3 * 1.9e9));
  return 0;
}

int main() {
  float* buf;
  const size_t cap = 1ull << 28;
  CK(cudaMalloc(&buf, cap));
  CK(cudaMemset(buf, 0, cap));
  const char* names[4] = {"red.v2 (today)", "red.v4", "STS + cp.reduce.async.bulk", "cp.reduce.async.bulk only"};
  for (size_t region : {(size_t)4 << 20, (size_t)64 << 20, (size_t)256 << 20})
    for (int mode = 0; mode < 4; ++mode) {
      if (run<256>(buf, region, mode, names[mode])) return 1;
      if (run<512>(buf, region, mode, names[mode])) return 1;
      if (run<1024>(buf, region, mode, names[mode])) return 1;
    }
  // verify the bulk reduce really added: one known spot
  CK(cudaMemset(buf, 0, 4096));
  probe<256><<<1, kWarps * 32, kWarps * kRing * 2 * 256>>>(buf, 255u, 8, 2);  // every run lands on bytes [0, 256)
  CK(cudaDeviceSynchronize());
  float h[64];
  CK(cudaMemcpy(h, buf, sizeof h, cudaMemcpyDeviceToHost));
  // 16 warps x 8 iterations x 2 runs x value (it & 3): sum over it of (it & 3) = 12 -> 16 * 2 * 12 = 384
  printf("bulk-reduce check: buf[0] = %g buf[63] = %g (expect 384)\n", h[0], h[63]);
  return (h[0] == 384.f && h[63] == 384.f) ? 0 : 2;
}
