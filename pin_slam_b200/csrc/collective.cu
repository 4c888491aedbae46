// NCCL from inside the library: the all-reduce of data-parallel map training is enqueued on the kernel stream by
// pinb200_map_iterations itself, between the backward kernels and the Adam kernels of every iteration, so that the
// whole multi-iteration training loop stays ONE host call on every rank (round 1 issued two host calls and a Python
// all_reduce per iteration: the launch gaps cost more than the 3.8 MB exchange over NVLink).
//
// libnccl is resolved at run time (dlopen of the copy PyTorch already loaded): libpinb200.so has no link-time
// dependency on it and still loads on a box without NCCL (the ABI tests run on CPU-only machines).
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace pinb {

struct NcclUid {
  char internal[128];
};
typedef int (*fn_get_uid)(NcclUid*);
typedef int (*fn_init_rank)(void**, int, NcclUid, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.get_uid = (fn_get_uid)dlsym(api.lib, "ncclGetUniqueId");
      api.init_rank = (fn_init_rank)dlsym(api.lib, "ncclCommInitRank");
      api.all_reduce = (fn_all_reduce)dlsym(api.lib, "ncclAllReduce");
      api.destroy = (fn_destroy)dlsym(api.lib, "ncclCommDestroy");
      api.errstr = (fn_errstr)dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.get_uid || !api.init_rank || !api.all_reduce || !api.destroy) {
    set_error("NCCL is not available in this process (dlopen libnccl.so.2 failed: import torch first)");
    return nullptr;
  }
  return &api;
}

static int nccl_check(NcclApi* a, int rc, const char* what) {
  if (rc == 0) return PINB200_OK;
  set_error("%s: NCCL error %d (%s)", what, rc, a->errstr ? a->errstr(rc) : "?");
  return PINB200_ERR_CUDA;
}

// sum-all-reduce of `count` floats in place on `stream` (ncclFloat = 7, ncclSum = 0)
int nccl_allreduce_sum(void* comm, float* buf, int64_t count, cudaStream_t stream) {
  NcclApi* a = nccl_api();
  if (!a) return PINB200_ERR_UNSUPPORTED;
  return nccl_check(a, a->all_reduce(buf, buf, (size_t)count, 7, 0, comm, stream), "ncclAllReduce");
}

}  // namespace pinb

using namespace pinb;

extern "C" int pinb200_nccl_unique_id(uint8_t* out128) {
  NcclApi* a = nccl_api();
  if (!a) return PINB200_ERR_UNSUPPORTED;
  if (!out128) {
    set_error("nccl_unique_id: null output");
    return PINB200_ERR_BAD_ARG;
  }
  NcclUid id;
  const int rc = nccl_check(a, a->get_uid(&id), "ncclGetUniqueId");
  if (rc) return rc;
  memcpy(out128, id.internal, 128);
  return PINB200_OK;
}

extern "C" int pinb200_nccl_init(const uint8_t* uid128, int32_t world, int32_t rank, void** comm_out) {
  NcclApi* a = nccl_api();
  if (!a) return PINB200_ERR_UNSUPPORTED;
  if (!uid128 || !comm_out || world < 1 || rank < 0 || rank >= world) {
    set_error("nccl_init: bad argument");
    return PINB200_ERR_BAD_ARG;
  }
  NcclUid id;
  memcpy(id.internal, uid128, 128);
  void* comm = nullptr;
  const int rc = nccl_check(a, a->init_rank(&comm, world, id, rank), "ncclCommInitRank");
  if (rc) return rc;
  *comm_out = comm;
  return PINB200_OK;
}

extern "C" int pinb200_nccl_destroy(void* comm) {
  NcclApi* a = nccl_api();
  if (!a) return PINB200_ERR_UNSUPPORTED;
  if (!comm) return PINB200_OK;
  return nccl_check(a, a->destroy(comm), "ncclCommDestroy");
}
