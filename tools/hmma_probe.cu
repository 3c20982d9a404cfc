// hmma_probe.cu — issue rate of the warp-level tensor-core instructions on B200 (legacy mma.sync path), per SM
// sub-partition: HMMA.16816.F32.BF16 (m16n8k16) and HMMA.1688.F32.BF16 (m16n8k8), 1..8 warps per sub-partition,
// 8 independent accumulators per warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K16>
__global__ void probe(float* out, long long* cycles, int iters) {
  float d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (K16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int k16 = 1; k16 >= 0; --k16)
    for (int warps = 4; warps <= 32; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (k16) probe<1><<<148, warps * 32>>>(out, cyc, iters);
        else probe<0><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double per = (double)h / ((double)iters * 8 * (warps / 4));  // cycles per MMA per sub-partition
      const double flop = k16 ? 4096.0 : 2048.0;
      printf("%s  %2d warps/SM: %.2f cycles per MMA per sub-partition -> %.0f TFLOP/s chip at 1.9 GHz\n",
             k16 ? "m16n8k16" : "m16n8k8 ", warps, per, flop / per * 4 * 148 * 1.9e9 / 1e12);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
