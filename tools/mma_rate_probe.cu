// tools/mma_rate_probe.cu — hardware probe: issue rate of small tcgen05.mma (kind::f16, M = 128, K = 16, operands in
// shared memory) as a function of N, operand alignment (row-shifted A view) and A major-ness.  One thread issues
// ITERS x 12 MMAs into one accumulator, commits, and the cycle count per MMA is printed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I jittor_dcn_b200/csrc tools/mma_rate_probe.cu -o tools/_build/mma_rate_probe
#include <cstdio>
#include <cstdlib>

#include "dcn_umma.cuh"
using namespace dcn::ptx;

template <int N>
__global__ void __launch_bounds__(1024) rate_kernel(long long* out, int iters, int shift, int a_mn, int spread, int alt, int bg, int commit_every, int elect) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, never, sink;
  __shared__ volatile int stop;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 16; i < 96 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&never, 1);
    mbar_init(&sink, 1);
    stop = 0;
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (elect && warp == 0) {
    // CUTLASS-style issue: the whole warp runs the (uniform) loop, elect.sync picks the lane for each instruction
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn != 0, false);
    const uint32_t a0 = smem_u32(smem) + shift * 128, b0 = smem_u32(smem) + 48 * 1024;
    const uint64_t da0 = make_sdesc_sw128(a0, 16, 1024), db0 = make_sdesc_sw128(b0, 16, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        uint32_t pe;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pe));
        if (pe) umma_bf16(tmem_base, da0 + 2u * (j & 3), db0 + 2u * (j & 3), idesc, 1u);
      }
    }
    const long long t1 = clock64();
    if (tid == 0) {
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      out[0] = t1 - t0;
      out[1] = clock64() - t0;
    }
  } else if (!elect && tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn != 0, false), idesc_h = make_idesc_bf16(128, N / 2, a_mn != 0, false);
    const uint32_t a0 = smem_u32(smem) + shift * 128, b0 = smem_u32(smem) + 48 * 1024;
    // descriptors are loop invariant: the timed loop holds nothing but the 12 tcgen05.mma (plus the optional commit)
    uint64_t da[4], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      da[j] = a_mn ? make_sdesc_sw128(a0 + j * 2048, 17408, 1024) : make_sdesc_sw128(a0 + j * 32, 16, 1024);
      db[j] = make_sdesc_sw128(b0 + j * 32, 16, 1024);
    }
    uint32_t dsts[12], ids[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      dsts[j] = tmem_base + (spread ? (uint32_t)((j % spread) * N) : 0u);
      ids[j] = (alt && (j & 1)) ? idesc_h : idesc;
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 12; ++j)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dsts[j]),
                     "l"(da[j & 3]), "l"(db[j & 3]), "r"(ids[j])
                     : "memory");
      // commit_every: a tcgen05.commit (arrive on an mbarrier nobody waits for) after every `commit_every` * 12 MMAs
      if (commit_every && (it % commit_every) == commit_every - 1) umma_commit(&sink);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    stop = 1;
  } else if (warp >= 4) {
    // background traffic of the other roles: bg 1 = all lanes spin on an mbarrier that never completes,
    // bg 2 = 8-byte shared-memory stores, bg 3 = lane 0 polls with nanosleep
    uint8_t* scratch = smem + 64 * 1024 + (tid & 1023) * 8;
    while (!stop) {
      if (bg == 1) { (void)mbar_try_wait(&never, 0); }
      else if (bg == 2) { asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(smem_u32(scratch)), "r"(tid) : "memory"); }
      else if (bg == 3) { if ((tid & 31) == 0) (void)mbar_try_wait(&never, 0); __nanosleep(64); }
      else break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int N>
static void run(int shift, int a_mn, int spread, int alt = 0, int bg = 0, int commit_every = 0, int elect = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 2000, smem = 97 * 1024;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N><<<1, bg ? 896 : 128, smem>>>(d, iters, shift, a_mn, spread, alt, bg, commit_every, elect);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("elect=%d alt_idesc=%d background=%d commit_every=%d ", elect, alt, bg, commit_every * 12);
  printf("N=%3d shift=%d A %s accumulators=%d : issue %.1f clk/MMA, complete %.1f clk/MMA %s\n", N, shift,
         a_mn ? "MN-major" : "K-major ", spread ? spread : 1, (double)h[0] / (iters * 12), (double)h[1] / (iters * 12),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<16>(0, 0, 0); run<32>(0, 0, 0); run<64>(0, 0, 0); run<128>(0, 0, 0); run<256>(0, 0, 0);
  run<32>(1, 0, 0); run<32>(2, 0, 0); run<64>(1, 0, 0);
  run<32>(0, 0, 3); run<32>(0, 0, 6); run<64>(0, 0, 3);
  run<32>(0, 1, 0); run<32>(1, 1, 0); run<32>(0, 1, 3);
  run<64>(0, 0, 0, 1, 0);
  run<64>(0, 0, 0, 0, 1); run<64>(0, 0, 0, 0, 2); run<64>(0, 0, 0, 0, 3); run<64>(1, 0, 0, 1, 1);
  run<64>(0, 0, 0, 0, 0, 1); run<64>(0, 0, 0, 0, 0, 2); run<64>(0, 0, 0, 0, 0, 6); run<32>(0, 0, 0, 0, 0, 2);
  run<32>(0, 0, 0, 0, 0, 0, 1); run<64>(0, 0, 0, 0, 0, 0, 1); run<128>(0, 0, 0, 0, 0, 0, 1); run<64>(1, 0, 0, 0, 0, 0, 1);
  return 0;
}
