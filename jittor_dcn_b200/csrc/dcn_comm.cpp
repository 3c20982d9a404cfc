// dcn_comm.cpp — data-parallel helpers of the C ABI: one NCCL communicator per process
// (one process per GPU), one in-place float32 sum all-reduce over a flat gradient bucket.
// NCCL is dlopen'ed on first use so that libdcn_b200.so itself has no link-time dependency
// (inside a PyTorch process this binds to the libnccl.so.2 torch already loaded).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

#include "../../include/dcn_b200.h"

namespace dcn {
void set_error(const char* fmt, ...);
int scale_f32(float* buf, size_t count, float scale, cudaStream_t st);  // dcn_api helpers
}

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_once;

bool load_nccl() {
  std::call_once(g_once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
  });
  const bool ok = g_nccl.handle && g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce &&
                  g_nccl.CommDestroy;
  if (!ok) dcn::set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
  return ok;
}

int nccl_fail(int rc, const char* what) {
  dcn::set_error("NCCL error %d (%s) at %s", rc,
                 g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?", what);
  return DCN_ERR_NCCL;
}

}  // namespace

extern "C" {

int dcn_comm_unique_id(void* out128_host) {
  if (!out128_host) return DCN_ERR_NULL_POINTER;
  if (!load_nccl()) return DCN_ERR_NCCL;
  NcclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc) return nccl_fail(rc, "ncclGetUniqueId");
  memcpy(out128_host, &id, sizeof id);
  return DCN_OK;
}

int dcn_comm_init(int rank, int world, const void* unique_id128_host, void** comm) {
  if (!unique_id128_host || !comm) return DCN_ERR_NULL_POINTER;
  if (world <= 0 || rank < 0 || rank >= world) {
    dcn::set_error("bad rank/world %d/%d", rank, world);
    return DCN_ERR_BAD_SHAPE;
  }
  if (!load_nccl()) return DCN_ERR_NCCL;
  NcclUniqueId id;
  memcpy(&id, unique_id128_host, sizeof id);
  NcclComm c = nullptr;
  int rc = g_nccl.CommInitRank(&c, world, id, rank);
  if (rc) return nccl_fail(rc, "ncclCommInitRank");
  *comm = c;
  return DCN_OK;
}

int dcn_allreduce_sum_f32(void* comm, void* buf, size_t count, float scale, void* stream) {
  if (!comm || !buf) return DCN_ERR_NULL_POINTER;
  if (!load_nccl()) return DCN_ERR_NCCL;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = g_nccl.AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, (NcclComm)comm, st);
  if (rc) return nccl_fail(rc, "ncclAllReduce");
  if (scale != 1.0f) return dcn::scale_f32((float*)buf, count, scale, st);
  return DCN_OK;
}

int dcn_comm_destroy(void* comm) {
  if (!comm) return DCN_OK;
  if (!load_nccl()) return DCN_ERR_NCCL;
  int rc = g_nccl.CommDestroy((NcclComm)comm);
  return rc ? nccl_fail(rc, "ncclCommDestroy") : DCN_OK;
}

}  // extern "C"
