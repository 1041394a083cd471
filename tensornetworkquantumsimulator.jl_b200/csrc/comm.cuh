// comm.cuh — NCCL plumbing for the sharded engine.  libnccl.so.2 (the copy
// PyTorch already loaded into the process) is bound at run time with dlopen/dlsym, so the library
// has no link-time dependency on it and single-GPU users never touch it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <memory>
#include <string>

namespace tnqs {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclChar = 0, kNcclFloat64 = 8 };
enum { kNcclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  static NcclApi& get() {
    static NcclApi api;
    if (!api.lib) {
      const char* env = std::getenv("TNQS_NCCL_LIB");
      const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
      for (const char* n : names) {
        if (!n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
      }
      if (!api.lib) throw Error(TNQS_EINVAL, "cannot load libnccl.so.2 (import torch first or set TNQS_NCCL_LIB)");
      auto sym = [&](const char* s) {
        void* p = dlsym(api.lib, s);
        if (!p) throw Error(TNQS_EINVAL, std::string("libnccl is missing symbol ") + s);
        return p;
      };
      api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
      api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
      api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
      api.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclBroadcast");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
      api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
      api.GroupStart = (int (*)())sym("ncclGroupStart");
      api.GroupEnd = (int (*)())sym("ncclGroupEnd");
      api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    }
    return api;
  }
  void check(int r, const char* what) {
    if (r != 0) throw Error(TNQS_ECUDA, std::string(what) + ": " + (GetErrorString ? GetErrorString(r) : "nccl error"));
  }
};

// shared between a cache and its clones
struct CommHandle {
  NcclComm comm = nullptr;
  int rank = 0, nranks = 1;
  ~CommHandle() {
    if (comm) NcclApi::get().CommDestroy(comm);
  }
};

}  // namespace tnqs
