// The path's only collective through the C ABI: every rank owns a contiguous slice of the image
// batch, runs the denoising loop with zero per-step communication, and the finished latents are all-gathered ONCE over
// NCCL (NVLink 5 / NVSwitch). For contrast, the reference's only on-path collective site runs 4 scatters + 1 all_gather
// PER STEP (ppdiffusers/pipelines/stable_diffusion_3/pipeline_stable_diffusion_3.py:803-839, INFERENCE_OPTIMIZE_BP).
//
// NCCL is loaded at run time (dlopen of the libnccl.so.2 the caller names: the wheel torch ships, or the system one), so
// libb200mix.so has no link-time dependency on it and single-GPU users never touch it. The communicator is created from
// this side (ncclGetUniqueId on rank 0, the 128-byte id handed to the other ranks by the caller's own launcher /
// rendezvous, ncclCommInitRank on every rank), so the boundary stays plain C: no torch types, no ProcessGroup.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace b200 {

typedef struct {
  char internal[128];
} nccl_unique_id;
typedef void* nccl_comm_t;
enum { NCCL_UINT8 = 1 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;

static int nccl_fail(const char* what, int rc) {
  set_error("%s failed: %s (nccl status %d)", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?", rc);
  return B200MIX_ERR_CUDA;
}

}  // namespace b200

using namespace b200;

extern "C" int b200mix_nccl_load(const char* libnccl_path) {
  if (g_nccl.handle) return 0;
  const char* path = (libnccl_path && libnccl_path[0]) ? libnccl_path : "libnccl.so.2";
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("nccl_load: dlopen(%s) failed: %s", path, dlerror());
    return B200MIX_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
    set_error("nccl_load: %s does not export the NCCL 2 API", path);
    dlclose(h);
    return B200MIX_ERR_UNSUPPORTED;
  }
  g_nccl = a;
  return 0;
}

extern "C" int b200mix_nccl_version(void) {
  int v = 0;
  if (!g_nccl.handle || !g_nccl.GetVersion || g_nccl.GetVersion(&v) != 0) return -1;
  return v;
}

extern "C" int b200mix_nccl_unique_id(void* id128) {
  B200_CHECK_ARG(id128, "nccl_unique_id: null pointer");
  B200_CHECK_ARG(g_nccl.handle, "nccl_unique_id: call b200mix_nccl_load first");
  nccl_unique_id id;
  if (int rc = g_nccl.GetUniqueId(&id)) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int b200mix_comm_init(void** comm, int32_t world_size, int32_t rank, const void* id128) {
  if (int rc = ensure_device()) return rc;  // b200mix_init(device) selected this rank's GPU
  B200_CHECK_ARG(comm && id128 && world_size > 0 && rank >= 0 && rank < world_size, "comm_init: bad arguments");
  B200_CHECK_ARG(g_nccl.handle, "comm_init: call b200mix_nccl_load first");
  nccl_unique_id id;
  memcpy(&id, id128, sizeof(id));
  nccl_comm_t c = nullptr;
  if (int rc = g_nccl.CommInitRank(&c, world_size, id, rank)) return nccl_fail("ncclCommInitRank", rc);
  *comm = c;
  return 0;
}

extern "C" int b200mix_allgather_latents(void* comm, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
  if (int rc = ensure_device()) return rc;
  B200_CHECK_ARG(comm && send && recv && bytes_per_rank > 0, "allgather_latents: bad arguments");
  B200_CHECK_ARG(g_nccl.handle, "allgather_latents: call b200mix_nccl_load first");
  if (int rc = g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, NCCL_UINT8, comm, reinterpret_cast<cudaStream_t>(stream)))
    return nccl_fail("ncclAllGather", rc);
  return 0;
}

extern "C" int b200mix_comm_destroy(void* comm) {
  B200_CHECK_ARG(comm && g_nccl.handle, "comm_destroy: bad arguments");
  if (int rc = g_nccl.CommDestroy(comm)) return nccl_fail("ncclCommDestroy", rc);
  return 0;
}
