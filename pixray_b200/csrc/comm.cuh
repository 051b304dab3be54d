// NCCL binding for the cutout-sharded multi-GPU mode.  libnccl.so.2 is resolved at run time with
// dlopen (torch has normally loaded the same soname already); only the handful of symbols the hot path needs are
// declared here, ABI-compatible with nccl.h 2.x (ncclUniqueId = 128 bytes, ncclSum = 0, ncclMin = 3, ncclFloat32 = 7, ncclFloat64 = 8).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstring>
#include <string>

namespace pxr {

struct NcclUniqueId {
  char internal[128];
};

class Comm {
 public:
  typedef int (*GetUniqueIdFn)(NcclUniqueId*);
  typedef int (*InitRankFn)(void**, int, NcclUniqueId, int);
  typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  typedef int (*DestroyFn)(void*);
  typedef const char* (*ErrStrFn)(int);
  typedef int (*GroupFn)(void);

  static Comm& api() {
    static Comm c;
    return c;
  }
  bool load(std::string* err) {
    if (lib_) return true;
    const char* env = getenv("PXR_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      lib_ = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib_) break;
    }
    if (!lib_) {
      if (err) *err = "cannot dlopen libnccl.so.2 (set PXR_NCCL_LIB to torch's nvidia/nccl/lib/libnccl.so.2)";
      return false;
    }
    get_unique_id = (GetUniqueIdFn)dlsym(lib_, "ncclGetUniqueId");
    init_rank = (InitRankFn)dlsym(lib_, "ncclCommInitRank");
    all_reduce = (AllReduceFn)dlsym(lib_, "ncclAllReduce");
    destroy = (DestroyFn)dlsym(lib_, "ncclCommDestroy");
    err_str = (ErrStrFn)dlsym(lib_, "ncclGetErrorString");
    group_start = (GroupFn)dlsym(lib_, "ncclGroupStart");
    group_end = (GroupFn)dlsym(lib_, "ncclGroupEnd");
    if (!get_unique_id || !init_rank || !all_reduce || !destroy || !group_start || !group_end) {
      if (err) *err = "libnccl is missing required symbols";
      return false;
    }
    return true;
  }
  GetUniqueIdFn get_unique_id = nullptr;
  InitRankFn init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  DestroyFn destroy = nullptr;
  ErrStrFn err_str = nullptr;
  GroupFn group_start = nullptr, group_end = nullptr;
  static constexpr int kSum = 0, kMin = 3, kFloat32 = 7, kFloat64 = 8;

 private:
  void* lib_ = nullptr;
};

}  // namespace pxr
