// dcn_p2p.cu — the gradient all-reduce of the data-parallel step as ONE kernel over NVLink peer memory.
//
// the DCN path shards along the batch only; a training step ends with ONE sum over ranks of the flat
// gradient bucket (435,862 floats = 1.74 MB for the detector).  At that size an all-reduce is pure latency, so it is
// done "one shot": every rank publishes its bucket in a buffer its peers have mapped (CUDA IPC; NVSwitch gives every
// GPU full bandwidth to every peer), and every rank sums all world copies itself, in rank order, so that all ranks
// hold bit-identical results.  No NCCL call, no host synchronisation, nothing the CUDA-graph capture of the training
// step cannot record: the epoch counter that sequences the exchanges lives in device memory and is advanced by the
// kernel itself, so a replayed graph keeps working.
//
// Protocol, per CTA b (slice b of the bucket), epoch e (parity q = e & 1):
//   1. copy slice b of the caller's buffer into the own published buffer pub[q]
//   2. release: flag[q][my_rank][b] := e in EVERY peer's flag array (st.release.sys)
//   3. acquire: spin until own flag[q][p][b] >= e for every peer p   (ld.acquire.sys)
//   4. out[i] = scale * sum_p pub_p[q][i] over slice b, p = 0 .. world-1 in order, written to the caller's buffer
// pub[q] of epoch e is only overwritten in epoch e + 2: a rank reaches step 1 of e + 2 only after step 3 of e + 1, i.e.
// after every peer's CTA b released e + 1, which it does after finishing step 4 of e (same CTA index, stream order).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "dcn_common.cuh"

namespace dcn {

namespace p2p {

constexpr int kMaxWorld = 16;
constexpr int kBlocks = 64;     // CTAs = slices; all co-resident (<= 148 SMs), so the pairwise waits cannot deadlock
constexpr int kThreads = 512;

struct Comm {
  int rank, world;
  size_t cap_floats;            // capacity of one published buffer
  void* base;                   // own allocation: [pub[0] | pub[1] | flags[2][kMaxWorld][kBlocks] | epoch]
  void* peer_base[kMaxWorld];   // mapped allocations of all ranks (own entry = base)
  bool opened[kMaxWorld];
};

__host__ __device__ inline size_t pub_bytes(size_t cap_floats) { return (cap_floats * 4 + 1023) / 1024 * 1024; }
__host__ __device__ inline size_t flags_off(size_t cap_floats) { return 2 * pub_bytes(cap_floats); }
__host__ __device__ inline size_t epoch_off(size_t cap_floats) {
  return flags_off(cap_floats) + sizeof(uint32_t) * 2 * kMaxWorld * kBlocks;
}
inline size_t total_bytes(size_t cap_floats) { return epoch_off(cap_floats) + 1024; }

struct Ptrs {
  uint8_t* base[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads) allreduce_kernel(Ptrs P, int rank, int world, size_t cap_floats,
                                                             float* __restrict__ buf, size_t count, float scale) {
  uint8_t* mine = P.base[rank];
  uint32_t* epoch_p = reinterpret_cast<uint32_t*>(mine + epoch_off(cap_floats));
  // every CTA reads the epoch the PREVIOUS launch left; the last CTA to finish advances it (done counter below)
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_p) + 1;
  const int q = (int)(e & 1u);
  const int b = blockIdx.x;
  // slice b in units of float4 (count padded up by the caller to a multiple of 4 within capacity)
  const size_t n4 = (count + 3) / 4, per = (n4 + kBlocks - 1) / kBlocks;
  const size_t i0 = (size_t)b * per, i1 = i0 + per < n4 ? i0 + per : n4;
  float4* pub = reinterpret_cast<float4*>(mine + (size_t)q * pub_bytes(cap_floats));
  const float4* src = reinterpret_cast<const float4*>(buf);
  for (size_t i = i0 + threadIdx.x; i < i1; i += kThreads) pub[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int p = threadIdx.x;
    uint32_t* f = reinterpret_cast<uint32_t*>(P.base[p] + flags_off(cap_floats)) + ((size_t)q * kMaxWorld + rank) * kBlocks + b;
    st_release_sys(f, e);
  }
  if ((int)threadIdx.x < world) {
    const int p = threadIdx.x;
    const uint32_t* f = reinterpret_cast<const uint32_t*>(mine + flags_off(cap_floats)) + ((size_t)q * kMaxWorld + p) * kBlocks + b;
    // epochs only grow (wrap-around after 2^32 steps is not a concern)
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - e) < 0) {
      __nanosleep(200);
      if (++spins > (1u << 28)) __trap();  // a peer never arrived (~1 min): fail loudly instead of hanging the GPU
    }
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(buf);
  for (size_t i = i0 + threadIdx.x; i < i1; i += kThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; ++p) {
      const float4 v = ld_peer(reinterpret_cast<const float4*>(P.base[p] + (size_t)q * pub_bytes(cap_floats)) + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    dst[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  }
  // advance the epoch once per launch: the last CTA to get here does it (counter right behind the epoch word)
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* done = epoch_p + 1;
    __threadfence();
    if (atomicAdd(done, 1u) == (uint32_t)gridDim.x - 1) {
      *done = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(epoch_p) = e;
    }
  }
}

}  // namespace p2p

}  // namespace dcn

using namespace dcn;

extern "C" {

size_t dcn_p2p_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int dcn_p2p_create(int rank, int world, size_t max_floats, void** comm) {
  if (!comm) return DCN_ERR_NULL_POINTER;
  if (world <= 0 || world > p2p::kMaxWorld || rank < 0 || rank >= world || max_floats == 0) {
    set_error("dcn_p2p_create: bad rank/world/size %d/%d/%zu (world <= %d)", rank, world, max_floats, p2p::kMaxWorld);
    return DCN_ERR_BAD_SHAPE;
  }
  p2p::Comm* c = new p2p::Comm();
  c->rank = rank;
  c->world = world;
  c->cap_floats = (max_floats + 3) / 4 * 4;
  for (int p = 0; p < p2p::kMaxWorld; ++p) {
    c->peer_base[p] = nullptr;
    c->opened[p] = false;
  }
  const size_t bytes = p2p::total_bytes(c->cap_floats);
  cudaError_t e = cudaMalloc(&c->base, bytes);
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(e, "cudaMalloc(p2p buffer)");
  }
  e = cudaMemset(c->base, 0, bytes);
  if (e != cudaSuccess) {
    cudaFree(c->base);
    delete c;
    return cuda_fail(e, "cudaMemset(p2p buffer)");
  }
  cudaDeviceSynchronize();
  c->peer_base[rank] = c->base;
  *comm = c;
  return DCN_OK;
}

int dcn_p2p_local_handle(void* comm, void* out_handle_host) {
  if (!comm || !out_handle_host) return DCN_ERR_NULL_POINTER;
  p2p::Comm* c = (p2p::Comm*)comm;
  cudaIpcMemHandle_t h;
  DCN_CUDA_TRY(cudaIpcGetMemHandle(&h, c->base));
  memcpy(out_handle_host, &h, sizeof h);
  return DCN_OK;
}

int dcn_p2p_connect(void* comm, const void* all_handles_host) {
  if (!comm || !all_handles_host) return DCN_ERR_NULL_POINTER;
  p2p::Comm* c = (p2p::Comm*)comm;
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t*)all_handles_host + (size_t)p * sizeof h, sizeof h);
    DCN_CUDA_TRY(cudaIpcOpenMemHandle(&c->peer_base[p], h, cudaIpcMemLazyEnablePeerAccess));
    c->opened[p] = true;
  }
  return DCN_OK;
}

int dcn_p2p_allreduce_sum_f32(void* comm, void* buf, size_t count, float scale, void* stream) {
  if (!comm || !buf) return DCN_ERR_NULL_POINTER;
  p2p::Comm* c = (p2p::Comm*)comm;
  if ((uintptr_t)buf & 15u) {
    set_error("dcn_p2p_allreduce_sum_f32: buffer is not 16-byte aligned");
    return DCN_ERR_MISALIGNED;
  }
  if (count > c->cap_floats) {
    set_error("dcn_p2p_allreduce_sum_f32: %zu floats exceed the communicator's capacity %zu", count, c->cap_floats);
    return DCN_ERR_WORKSPACE;
  }
  for (int p = 0; p < c->world; ++p)
    if (!c->peer_base[p]) {
      set_error("dcn_p2p_allreduce_sum_f32: peer %d is not connected (dcn_p2p_connect)", p);
      return DCN_ERR_UNSUPPORTED;
    }
  p2p::Ptrs P;
  for (int p = 0; p < p2p::kMaxWorld; ++p) P.base[p] = (uint8_t*)c->peer_base[p];
  cudaStream_t st = (cudaStream_t)stream;
  KernelScope scope("p2p_allreduce_kernel", st);
  // the kernel walks float4s: the caller's buffer must be readable / writable up to the next multiple of 4 floats
  // (GradBucket pads its flat buffer)
  p2p::allreduce_kernel<<<p2p::kBlocks, p2p::kThreads, 0, st>>>(P, c->rank, c->world, c->cap_floats, (float*)buf, count,
                                                                scale);
  DCN_KERNEL_CHECK("p2p_allreduce_kernel");
  return DCN_OK;
}

int dcn_p2p_destroy(void* comm) {
  if (!comm) return DCN_OK;
  p2p::Comm* c = (p2p::Comm*)comm;
  cudaDeviceSynchronize();
  for (int p = 0; p < c->world; ++p)
    if (c->opened[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  cudaFree(c->base);
  delete c;
  return DCN_OK;
}

}  // extern "C"
