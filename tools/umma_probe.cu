// tools/umma_probe.cu — standalone hardware probe (not part of the library):
//   1. tcgen05.mma descriptor/layout unit tests (K-major and MN-major, 128B swizzle) against
//      exact integer-valued GEMMs computed on the host;
//   2. micro-benchmarks that size the sampling / scatter kernels: red.global.add (scalar,
//      coalesced, v4), shared-memory float atomics, LDG.128 gathers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo
//        -I jittor_dcn_b200/csrc tools/umma_probe.cu -o tools/_build/umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dcn_umma.cuh"

using namespace dcn::ptx;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

// ------------------------------------------------------------------ UMMA unit test
// D[128, N] = sum_kb A[128, 64*KB] * B[N, 64*KB]^T
template <int N, int KB, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128) umma_test_kernel(const float* __restrict__ A,
                                                        const float* __restrict__ Bm,
                                                        float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int KT = 64 * KB;
  constexpr int A_TILE = 128 * 64 * 2, B_TILE = N * 64 * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + KB * A_TILE;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // MN-major geometry: atom (m/64, k/8) at (k/8)*SBO + (m/64)*LBO
  constexpr uint32_t A_LBO = 1024, A_SBO = 2048;                 // 128 rows = 2 MN atoms
  constexpr uint32_t B_LBO = 1024, B_SBO = 1024 * (N / 64);      // N/64 MN atoms
  for (int i = tid; i < 128 * KT; i += 128) {
    const int m = i / KT, k = i % KT, kb = k / 64, kk = k % 64;
    const uint32_t off = A_MN ? mnmajor_sw128_off(m, kk, A_LBO, A_SBO) : kmajor_sw128_off(m, kk);
    *reinterpret_cast<__nv_bfloat16*>(sA + kb * A_TILE + off) = __float2bfloat16_rn(A[i]);
  }
  for (int i = tid; i < N * KT; i += 128) {
    const int n = i / KT, k = i % KT, kb = k / 64, kk = k % 64;
    const uint32_t off = B_MN ? mnmajor_sw128_off(n, kk, B_LBO, B_SBO) : kmajor_sw128_off(n, kk);
    *reinterpret_cast<__nv_bfloat16*>(sB + kb * B_TILE + off) = __float2bfloat16_rn(Bm[i]);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<(N < 32 ? 32 : N)>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, A_MN, B_MN);
    for (int kb = 0; kb < KB; ++kb)
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint32_t a_addr = smem_u32(sA + kb * A_TILE) + (A_MN ? k4 * 2 * A_SBO : k4 * 32);
        const uint32_t b_addr = smem_u32(sB + kb * B_TILE) + (B_MN ? k4 * 2 * B_SBO : k4 * 32);
        const uint64_t da = A_MN ? make_sdesc_sw128(a_addr, A_LBO, A_SBO) : make_sdesc_sw128(a_addr, 16, 1024);
        const uint64_t db = B_MN ? make_sdesc_sw128(b_addr, B_LBO, B_SBO) : make_sdesc_sw128(b_addr, 16, 1024);
        umma_bf16(tmem_base, da, db, idesc, (kb | k4) ? 1u : 0u);
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<(N < 32 ? 32 : N)>(tmem_base);
}

template <int N, int KB, bool A_MN, bool B_MN>
static bool run_umma_test(const char* name) {
  constexpr int KT = 64 * KB;
  std::vector<float> A(128 * KT), B(N * KT), D(128 * N), R(128 * N, 0.f);
  srand(1234);
  for (auto& v : A) v = (float)(rand() % 9 - 4);
  for (auto& v : B) v = (float)(rand() % 9 - 4);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < KT; ++k) s += A[m * KT + k] * B[n * KT + k];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4));
  const int smem = KB * (128 * 64 * 2 + N * 64 * 2) + 1024;
  auto kern = umma_test_kernel<N, KB, A_MN, B_MN>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("UMMA %-28s : CUDA ERROR %s\n", name, cudaGetErrorString(e));
    exit(3);
  }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0, first = -1;
  for (int i = 0; i < 128 * N; ++i)
    if (D[i] != R[i]) {
      if (first < 0) first = i;
      ++bad;
    }
  printf("UMMA %-28s : %s (%d / %d mismatches", name, bad ? "FAIL" : "PASS", bad, 128 * N);
  if (bad) printf("; first at m=%d n=%d got %g want %g", first / N, first % N, D[first], R[first]);
  printf(")\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad == 0;
}

// ------------------------------------------------------------------ micro-benchmarks
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// mode 0: warp adds 32 consecutive floats at a random 128B-aligned spot (coalesced RED.32)
// mode 1: every lane adds to its own random float (scattered RED.32)
// mode 2: red.global.add.v4.f32, 8 lanes cover 128 B, 4 random spots per warp instruction
__global__ void __launch_bounds__(256) red_bench(float* buf, uint32_t nfloats_mask, int iters, int mode) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      uint32_t base = (hash32(gw * 7919u + it) & nfloats_mask) & ~31u;
      atomicAdd(buf + base + lane, 1.0f);
    } else if (mode == 1) {
      uint32_t a = hash32((gw * 32 + lane) * 7919u + it) & nfloats_mask;
      atomicAdd(buf + a, 1.0f);
    } else {
      uint32_t base = (hash32((gw * 4 + (lane >> 3)) * 7919u + it) & nfloats_mask) & ~31u;
      float* p = buf + base + (lane & 7) * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(1.f), "f"(1.f),
                   "f"(1.f)
                   : "memory");
    }
  }
}

__global__ void __launch_bounds__(256) smem_atomic_bench(float* out, int iters, int stride) {
  __shared__ float s[8192];
  for (int i = threadIdx.x; i < 8192; i += 256) s[i] = 0.f;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int it = 0; it < iters; ++it) {
    uint32_t base = (hash32(w * 977u + it) & 127u) * 64;
    atomicAdd(&s[(base + lane * stride) & 8191], 1.0f);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// LDG.128 gather: each quarter-warp reads 128 contiguous bytes at a random 128B-aligned spot
__global__ void __launch_bounds__(256) gather_bench(const float4* __restrict__ buf, uint32_t nvec_mask,
                                                    int iters, float* out) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int it = 0; it < iters; it += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t base = (hash32((gw * 4 + (lane >> 3)) * 7919u + it + u) & nvec_mask) & ~7u;
      float4 v = __ldg(buf + base + (lane & 7));
      acc += v.x + v.y + v.z + v.w;
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

static float time_ms(void (*launch)(void*), void* ctx) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch(ctx);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  launch(ctx);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main(int argc, char** argv) {
  bool ok = true;
  ok &= run_umma_test<64, 1, false, false>("K-major A/B N=64 KB=1");
  ok &= run_umma_test<64, 2, false, false>("K-major A/B N=64 KB=2");
  ok &= run_umma_test<256, 2, false, false>("K-major A/B N=256 KB=2");
  ok &= run_umma_test<32, 1, false, false>("K-major A/B N=32 KB=1");
  ok &= run_umma_test<64, 2, true, false>("MN-major A, K-major B N=64");
  ok &= run_umma_test<64, 2, false, true>("K-major A, MN-major B N=64");
  ok &= run_umma_test<128, 2, true, true>("MN-major A/B N=128");
  if (argc > 1) return ok ? 0 : 1;

  // ---- atomics
  const int blocks = 148 * 8, iters = 2000;
  float* buf;
  const size_t nfl = 1u << 26;  // 256 MB
  CK(cudaMalloc(&buf, nfl * 4));
  CK(cudaMemset(buf, 0, nfl * 4));
  struct Ctx { float* buf; uint32_t mask; int mode; int blocks; int iters; } c;
  auto launch = [](void* p) {
    Ctx* c = (Ctx*)p;
    red_bench<<<c->blocks, 256>>>(c->buf, c->mask, c->iters, c->mode);
  };
  for (uint32_t region_fl : {1u << 20, 1u << 26}) {  // 4 MB (L2 resident) and 256 MB
    for (int mode = 0; mode < 3; ++mode) {
      c = {buf, region_fl - 1, mode, blocks, iters};
      float ms = time_ms(launch, &c);
      double lane_ops = (double)blocks * 256 * iters;
      double floats = lane_ops * (mode == 2 ? 4 : 1);
      printf("RED mode=%d (%s) region=%4u MB: %.3f ms  %.1f G float-adds/s  %.1f G lane-ops/s\n", mode,
             mode == 0 ? "scalar coalesced 128B/warp" : mode == 1 ? "scalar scattered" : "v4, 4x128B/warp",
             region_fl >> 18, ms, floats / ms * 1e-6, lane_ops / ms * 1e-6);
    }
  }
  // ---- smem atomics
  float* o;
  CK(cudaMalloc(&o, 4096 * 4));
  for (int stride : {1, 2, 32}) {
    struct C2 { float* o; int stride; } c2{o, stride};
    auto l2 = [](void* p) { C2* c = (C2*)p; smem_atomic_bench<<<148 * 4, 256>>>(c->o, 4000, c->stride); };
    float ms = time_ms(l2, &c2);
    double ops = 148.0 * 4 * 256 * 4000;
    printf("SMEM atomicAdd.f32 lane-stride=%2d: %.3f ms  %.1f G lane-ops/s  (%.2f cyc/warp-instr/SM @1.9GHz)\n",
           stride, ms, ops / ms * 1e-6, ms * 1e-3 * 1.9e9 / (4 * 8 * 4000.0));
  }
  // ---- gathers
  for (uint32_t region_vec : {1u << 12, 1u << 18, 1u << 24}) {  // 64 KB, 4 MB, 256 MB
    struct C3 { const float4* b; uint32_t mask; float* o; } c3{(const float4*)buf, region_vec - 1, o};
    auto l3 = [](void* p) { C3* c = (C3*)p; gather_bench<<<148 * 8, 256>>>(c->b, c->mask, 4000, c->o); };
    float ms = time_ms(l3, &c3);
    double bytes = 148.0 * 8 * 256 * 4000 * 16;
    printf("LDG.128 gather (128B runs) region=%6u KB: %.3f ms  %.1f GB/s  (%.1f B/cyc/SM @1.9GHz)\n",
           region_vec >> 6, ms, bytes / ms * 1e-6, bytes / 148 / (ms * 1e-3 * 1.9e9));
  }
  return ok ? 0 : 1;
}
