// mc_nccl_* / mc_allgather_kv: the COLLECTIVE formulation of the token shard's K|V exchange behind the C ABI — the
// literal counterpart of the reference's `dist.all_gather(tensor_list, input_)` + `torch.cat` (videosys/core/comm.py:272-292, used by
// eval/magcache/experiments/opensora.py:284-293, 356-361): one ncclAllGather of the rows every rank has just projected, on the
// caller's stream. The product path of a token-sharded run is the peer-to-peer exchange of p2p.cu (copy-engine pushes consumed
// inside the attention kernel); this one is the plain form next to it (MC_SHARD_P2P=0 MC_SHARD_NCCL=capi), and what a host without
// CUDA IPC between its processes would use.
// NCCL is not linked: the library is looked up at run time (the copy a PyTorch process has already loaded, else libnccl.so.2 on the
// loader path), through the handful of entry points of its stable C API declared below — the .so has no NCCL dependency unless these
// functions are called.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace mc {
namespace {

struct NcclUniqueId {
  char internal[128];  // NCCL_UNIQUE_ID_BYTES
};
using NcclComm = void*;
constexpr int kNcclBfloat16 = 9, kNcclUint8 = 1;  // ncclDataType_t

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process already uses (PyTorch's), if any
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GroupStart && a.GroupEnd && a.GetErrorString;
  });
  return a;
}

int32_t need_api(const char* who) {
  if (api().ok) return MC_OK;
  set_error("%s: libnccl.so.2 (ncclGetUniqueId, ncclCommInitRank, ncclAllGather, ...) could not be loaded", who);
  return MC_ERR_STATE;
}

int32_t nccl_fail(int rc, const char* what) {
  set_error("%s: %s", what, api().GetErrorString(rc));
  return MC_ERR_CUDA;
}

}  // namespace
}  // namespace mc

struct mc_nccl {
  mc::NcclComm comm = nullptr;
  int32_t rank = 0, world = 1;
};

extern "C" {

int32_t mc_nccl_unique_id(void* id_out_128) {
  MC_CHECK_ARG(id_out_128 != nullptr, "mc_nccl_unique_id: null pointer");
  int32_t rc = mc::need_api("mc_nccl_unique_id");
  if (rc) return rc;
  mc::NcclUniqueId id;
  const int e = mc::api().GetUniqueId(&id);
  if (e != 0) return mc::nccl_fail(e, "ncclGetUniqueId");
  std::memcpy(id_out_128, &id, sizeof(id));
  return MC_OK;
}

mc_nccl* mc_nccl_init(int32_t rank, int32_t world, const void* unique_id_128) {
  if (unique_id_128 == nullptr || world < 1 || rank < 0 || rank >= world) {
    mc::set_error("mc_nccl_init: rank %d of %d / null id", rank, world);
    return nullptr;
  }
  if (mc::need_api("mc_nccl_init")) return nullptr;
  mc::NcclUniqueId id;
  std::memcpy(&id, unique_id_128, sizeof(id));
  mc_nccl* h = new mc_nccl();
  h->rank = rank, h->world = world;
  const int e = mc::api().CommInitRank(&h->comm, world, id, rank);  // collective: every rank of the job calls it, on its own device
  if (e != 0) {
    mc::nccl_fail(e, "ncclCommInitRank");
    delete h;
    return nullptr;
  }
  return h;
}

int32_t mc_nccl_destroy(mc_nccl* h) {
  if (h == nullptr) return MC_OK;
  int e = 0;
  if (h->comm != nullptr && mc::api().ok) e = mc::api().CommDestroy(h->comm);
  delete h;
  return e != 0 ? mc::nccl_fail(e, "ncclCommDestroy") : MC_OK;
}

int32_t mc_allgather_kv(mc_nccl* h, const void* k_local, const void* v_local, void* k_full, void* v_full, int64_t elems_per_rank, void* stream) {
  MC_CHECK_ARG(h != nullptr && h->comm != nullptr, "mc_allgather_kv: null communicator");
  MC_CHECK_ARG(k_local != nullptr && k_full != nullptr && elems_per_rank >= 1, "mc_allgather_kv: null pointer / empty");
  MC_CHECK_ARG((v_local == nullptr) == (v_full == nullptr), "mc_allgather_kv: v_local and v_full go together");
  const mc::NcclApi& a = mc::api();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int e = a.GroupStart();
  if (e != 0) return mc::nccl_fail(e, "ncclGroupStart");
  e = a.AllGather(k_local, k_full, static_cast<size_t>(elems_per_rank), mc::kNcclBfloat16, h->comm, s);
  if (e == 0 && v_local != nullptr) e = a.AllGather(v_local, v_full, static_cast<size_t>(elems_per_rank), mc::kNcclBfloat16, h->comm, s);
  const int e2 = a.GroupEnd();
  if (e != 0) return mc::nccl_fail(e, "ncclAllGather");
  if (e2 != 0) return mc::nccl_fail(e2, "ncclGroupEnd");
  return MC_OK;
}

}  // extern "C"
