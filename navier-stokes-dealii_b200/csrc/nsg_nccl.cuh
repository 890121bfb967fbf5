// nsg_nccl.cuh — NCCL is bound at run time (dlopen), never at link time: a process that also
// imports torch must use torch's bundled libnccl.so.2, and loading the system copy first would
// shadow it.  Order: $NSG_NCCL_LIB, an already loaded libnccl.so.2, then the default search path.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <string>

namespace nsg {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  void *handle = nullptr;
  std::string error;
};

inline NcclApi &nccl_api() {
  static NcclApi api;
  if (api.handle || !api.error.empty()) return api;
  void *h = nullptr;
  if (const char *p = std::getenv("NSG_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
    return api;
  }
  auto sym = [&](const char *n) {
    void *s = dlsym(h, n);
    if (!s) api.error = std::string("libnccl.so.2 lacks ") + n;
    return s;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  if (api.error.empty()) api.handle = h;
  return api;
}

}  // namespace nsg
