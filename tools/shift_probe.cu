// tools/shift_probe.cu — hardware probe (not part of the library): can ONE converted shared-memory window serve
// several convolution taps as shifted operand views?
//
// Window W: R rows of 64 bf16 (128 B per row), stored with the 128-byte swizzle keyed on the ABSOLUTE row index
// (kmajor_sw128_off).  The same bytes are
//   (1) a K-major A operand  A_s[m, k] = W[m + s, k]          (rows = GEMM rows; forward conv: tap kj = row shift s)
//   (2) an MN-major B operand B_s[n, k] = W[k + s, n]          (rows = GEMM K;    weight gradient: tap = K shift s)
// when the descriptor's start address is moved by s * 128 bytes — IF the tensor core applies the swizzle on the final
// address bits (then nothing else is needed) or honours the descriptor's "matrix base offset" field (bits 49-51,
// (start_address >> 7) & 7) for starts that are not 1024-byte aligned.  This probe tries both encodings for
// s = 0..9 against exact integer GEMMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I jittor_dcn_b200/csrc tools/shift_probe.cu -o tools/_build/shift_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dcn_umma.cuh"

using namespace dcn::ptx;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

constexpr int R = 160;  // window rows

__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, int base_off_mode) {
  uint64_t d = make_sdesc_sw128(addr, lbo, sbo);
  if (base_off_mode == 1) d |= (uint64_t)((addr >> 7) & 7u) << 49;
  return d;
}

// MODE 0: D[128, 32] = A_s * Bw^T, A_s rows = window rows m + s (K-major, K = 64), Bw [32][64] K-major
// MODE 1: D[128, 64] = G * B_s^T over K = 64 window rows: G [128][64] K-major (rows = o), B_s[n, k] = W[k + s, n]
template <int MODE>
__global__ void __launch_bounds__(128) shift_kernel(const float* __restrict__ Wv, const float* __restrict__ Other,
                                                    float* __restrict__ D, int shift, int bo_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                      // R rows x 128 B
  uint8_t* sO = smem + R * 128;            // R*128 = 20480 = 20 * 1024: aligned
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int N = MODE == 0 ? 32 : 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < R * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sW + kmajor_sw128_off(r, k)) = __float2bfloat16_rn(Wv[i]);
  }
  const int orows = MODE == 0 ? 32 : 128;
  for (int i = tid; i < orows * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sO + kmajor_sw128_off(r, k)) = __float2bfloat16_rn(Other[i]);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    if (MODE == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 32, false, false);
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t da = sdesc(smem_u32(sW) + shift * 128 + k4 * 32, 16, 1024, bo_mode);
        const uint64_t db = make_sdesc_sw128(smem_u32(sO) + k4 * 32, 16, 1024);
        umma_bf16(tmem_base, da, db, idesc, k4 ? 1u : 0u);
      }
    } else {
      // A = G K-major [128 rows o][64 k]; B = window MN-major: N = 64 channels contiguous (one MN atom), K = rows,
      // 8-row K groups 1024 B apart (sbo); one MMA = 16 K rows = 2 groups
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, true);
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t da = make_sdesc_sw128(smem_u32(sO) + k4 * 32, 16, 1024);
        const uint64_t db = sdesc(smem_u32(sW) + shift * 128 + k4 * 2048, 1024, 1024, bo_mode);
        umma_bf16(tmem_base, da, db, idesc, k4 ? 1u : 0u);
      }
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
  if (warp == 0) tmem_dealloc<64>(tmem_base);
}

template <int MODE>
static void run(int shift, int bo_mode) {
  constexpr int N = MODE == 0 ? 32 : 64;
  const int orows = MODE == 0 ? 32 : 128;
  std::vector<float> W(R * 64), O(orows * 64), D(128 * N), Ref(128 * N, 0.f);
  srand(77 + shift);
  for (auto& v : W) v = (float)(rand() % 9 - 4);
  for (auto& v : O) v = (float)(rand() % 9 - 4);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < 64; ++k)
        s += MODE == 0 ? W[(m + shift) * 64 + k] * O[n * 64 + k]    // A_s[m,k] * Bw[n,k]
                       : O[m * 64 + k] * W[(k + shift) * 64 + n];   // G[m,k] * W[k+s, n]
      Ref[m * N + n] = s;
    }
  float *dW, *dO, *dD;
  CK(cudaMalloc(&dW, W.size() * 4));
  CK(cudaMalloc(&dO, O.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dO, O.data(), O.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4));
  const int smem = R * 128 + 128 * 128 + 2048;
  CK(cudaFuncSetAttribute(shift_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  shift_kernel<MODE><<<1, 128, smem>>>(dW, dO, dD, shift, bo_mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mode %d shift %d base_offset_mode %d : CUDA ERROR %s\n", MODE, shift, bo_mode, cudaGetErrorString(e));
    exit(3);
  }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 128 * N; ++i) bad += D[i] != Ref[i];
  printf("%s shift %d  base_offset field %s : %s (%d / %d mismatches)\n",
         MODE == 0 ? "K-major A, row shift    " : "MN-major B, K-row shift ", shift,
         bo_mode ? "(addr>>7)&7" : "0          ", bad ? "FAIL" : "PASS", bad, 128 * N);
  cudaFree(dW); cudaFree(dO); cudaFree(dD);
}

int main() {
  for (int bo = 0; bo < 2; ++bo)
    for (int s = 0; s <= 9; ++s) run<0>(s, bo);
  for (int bo = 0; bo < 2; ++bo)
    for (int s = 0; s <= 9; ++s) run<1>(s, bo);
  return 0;
}
