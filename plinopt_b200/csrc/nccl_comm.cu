// nccl_comm.cu -- the one collective of the path, issued by the engine itself: an all-reduce (min / max) over a few int64 words
// of device memory through NCCL (NVLink / NVSwitch inside a box).  north_star (3): "a single tiny NCCL allreduce(min-with-index)
// picks the global best"; reference analogue: the `omp critical` of src/orbiter.cpp:298.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy a host process such as PyTorch has already loaded, else the
// system one), so the library has no link-time dependency on it and single-GPU users never touch it.  One communicator per process
// and device, one rank per GPU; the 128-byte unique id of rank 0 reaches the other ranks by whatever channel launched them
// (bench.py broadcasts it through torch.distributed; an MPI or file rendezvous works as well).
#include <dlfcn.h>

#include <cstring>

#include "plo_device.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclInt64 = 4, kNcclMax = 2, kNcclMin = 3 };  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.GetUniqueId = reinterpret_cast<int (*)(ncclUniqueId*)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(ncclComm_t*, int, ncclUniqueId, int)>(dlsym(h, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<int (*)(ncclComm_t)>(dlsym(h, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<int (*)(int*)>(dlsym(h, "ncclGetVersion"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) { dlclose(h); return nullptr; }
  api.handle = h;
  return &api;
}

int nccl_fail(NcclApi* a, const char* what, int rc) {
  plo::set_error("%s: NCCL error %d (%s)", what, rc, a && a->GetErrorString ? a->GetErrorString(rc) : "?");
  return PLO_E_CUDA;
}

}  // namespace

struct plo_comm {
  ncclComm_t comm;
  int rank, world;
};

extern "C" {

int plo_comm_nccl_version(void) {
  NcclApi* a = nccl();
  int v = 0;
  if (!a || !a->GetVersion || a->GetVersion(&v) != 0) return 0;
  return v;
}

int plo_comm_unique_id(uint8_t* id128) {
  if (!id128) { plo::set_error("plo_comm_unique_id: null buffer"); return PLO_E_ARG; }
  NcclApi* a = nccl();
  if (!a) { plo::set_error("plo_comm_unique_id: libnccl.so.2 could not be loaded (%s)", dlerror()); return PLO_E_NODEVICE; }
  ncclUniqueId id;
  const int rc = a->GetUniqueId(&id);
  if (rc) return nccl_fail(a, "ncclGetUniqueId", rc);
  memcpy(id128, id.internal, 128);
  return PLO_OK;
}

int plo_comm_create(plo_comm** comm, int rank, int world, const uint8_t* id128) {
  if (!comm || !id128 || world < 1 || rank < 0 || rank >= world) { plo::set_error("plo_comm_create: bad argument"); return PLO_E_ARG; }
  int rc = plo::check_device();
  if (rc) return rc;
  NcclApi* a = nccl();
  if (!a) { plo::set_error("plo_comm_create: libnccl.so.2 could not be loaded"); return PLO_E_NODEVICE; }
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  rc = a->CommInitRank(&c, world, id, rank);  // one rank per process, bound to the current device
  if (rc) return nccl_fail(a, "ncclCommInitRank", rc);
  *comm = new plo_comm{c, rank, world};
  return PLO_OK;
}

int plo_comm_rank(const plo_comm* c) { return c ? c->rank : -1; }
int plo_comm_size(const plo_comm* c) { return c ? c->world : 0; }

int plo_comm_allreduce_i64(plo_comm* c, int64_t* buf_device, uint64_t count, int op, void* stream) {
  if (!c || !buf_device || (op != PLO_REDUCE_MIN && op != PLO_REDUCE_MAX)) { plo::set_error("plo_comm_allreduce_i64: bad argument"); return PLO_E_ARG; }
  NcclApi* a = nccl();
  const int rc = a->AllReduce(buf_device, buf_device, (size_t)count, kNcclInt64, op == PLO_REDUCE_MIN ? kNcclMin : kNcclMax, c->comm, (cudaStream_t)stream);
  if (rc) return nccl_fail(a, "ncclAllReduce", rc);
  return PLO_OK;
}

void plo_comm_destroy(plo_comm* c) {
  if (!c) return;
  NcclApi* a = nccl();
  if (a && c->comm) a->CommDestroy(c->comm);
  delete c;
}

}  // extern "C"
