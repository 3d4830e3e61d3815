// comm.hpp -- NCCL plumbing of the multi-rank accept step (one process per GPU).
//
// The reference gathers every worker's ll array on the master and runs max / compare /
// where there (thejoker/multiproc_helpers.py:256-263, 373-381).  Here only the 8-byte
// max-key (integer MAX all-reduce), two counters and the accepted indices cross NVLink.
//
// NCCL is bound at run time (dlopen + dlsym), not at link time: the library must load on a
// machine without NCCL (single-GPU use, the CPU symbol tests), and inside a process that
// already carries an NCCL (PyTorch's bundled copy) it must use that same copy rather than a
// second one.  Types and enums come from <nccl.h>; only these entry points are used.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

namespace tjb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  void *dso = nullptr;
  std::string error;  // why loading failed, if it did
  bool ok = false;
};

inline NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // an NCCL that is already mapped into the process (torch's) wins; then the system's
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names)
      if (!api.dso) api.dso = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    for (const char *nm : names)
      if (!api.dso) api.dso = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!api.dso) {
      api.error = "NCCL not found (dlopen libnccl.so.2): multi-rank entry points are unavailable";
      return;
    }
    bool all = true;
    auto sym = [&](const char *name) {
      void *p = dlsym(api.dso, name);
      if (!p) {
        all = false;
        api.error = std::string("NCCL symbol missing: ") + name;
      }
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.ok = all;
  });
  return api;
}

}  // namespace tjb

struct TjbComm {
  ncclComm_t comm = nullptr;
  int n_ranks = 0, rank = 0, device = 0;
};
