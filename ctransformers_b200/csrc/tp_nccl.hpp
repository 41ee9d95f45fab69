// NCCL, bound at run time: the tensor-sharded mode (engine.cuh: TPShard) is the only user, so the library is opened lazily and a
// single-GPU process never needs it.  Inside a torch process `libnccl.so.2` resolves to the copy torch already loaded.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <stdexcept>
#include <string>

namespace ctb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  static NcclApi& get() {
    static NcclApi api = load();
    return api;
  }
  void check(ncclResult_t r, const char* what) const {
    if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL ") + what + ": " + (GetErrorString ? GetErrorString(r) : "error"));
  }

 private:
  static NcclApi load() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error(std::string("tensor parallel mode needs libnccl.so.2: ") + dlerror());
    NcclApi a;
    auto sym = [&](const char* name) {
      void* p = dlsym(h, name);
      if (!p) throw std::runtime_error(std::string("libnccl: missing symbol ") + name);
      return p;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    return a;
  }
};

}  // namespace ctb
