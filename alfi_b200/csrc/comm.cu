// Multi-GPU plumbing: one rank per GPU on one NVSwitch box.  Patches (the smoother's stars and
// the transfer's cell patches) and block rows of the level operators are sharded across ranks;
// level vectors are replicated.  The two exchange steps of the path are
//   * after a patch apply: ghost->owner sum + owner->ghost broadcast of y (PetscSF reduce+bcast
//     around PCApply_PATCH in the reference)  -> one ncclAllReduce(sum);
//   * after a row-sharded SpMV: owner->everyone broadcast of the owned rows (VecScatter of
//     MatMult_MPIBAIJ)                                             -> grouped ncclBroadcast.
// Both are enqueued on the ctx stream, so the cycle stays a fixed, sync-free launch sequence.
// NCCL is dlopen'ed (the torch-bundled libnccl.so.2 if already loaded, else the system one), so
// a single-GPU process never needs it.
#include <dlfcn.h>
#include <nccl.h>

#include "alfib_internal.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

NcclApi& nccl() {
  if (g_nccl.handle) return g_nccl;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("/usr/lib/x86_64-linux-gnu/libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw DeviceError{ALFIB_ECUDA, std::string("cannot load libnccl.so.2: ") + dlerror()};
  auto sym = [&](const char* name) {
    void* p = dlsym(h, name);
    if (!p) throw DeviceError{ALFIB_ECUDA, std::string("libnccl.so.2 lacks ") + name};
    return p;
  };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  g_nccl.handle = h;
  return g_nccl;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw DeviceError{ALFIB_ECUDA, std::string(what) + ": " + nccl().GetErrorString(r)};
}

}  // namespace

void comm_unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, &id, sizeof(id));
}

void comm_init(alfib_ctx* c, const void* id128, int rank, int nranks) {
  ALFIB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  ALFIB_REQUIRE(!c->comm, "communicator already initialised");
  c->rank = rank;
  c->nranks = nranks;
  if (nranks == 1) return;
  ALFIB_REQUIRE(id128 != nullptr, "nccl unique id required for nranks > 1");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  nccl_check(nccl().CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
  c->comm = comm;
}

void comm_destroy(alfib_ctx* c) {
  if (c->comm) nccl().CommDestroy((ncclComm_t)c->comm);
  c->comm = nullptr;
}

// y <- sum over ranks of y   (patch scatter: ghost->owner sum, owner->ghost broadcast)
void comm_allreduce_sum(alfib_ctx* c, double* y, size_t n) {
  if (c->nranks <= 1) return;
  nccl_check(nccl().AllReduce(y, y, n, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream), "ncclAllReduce");
}

// every rank owns dofs [start[r], start[r+1]) of y; afterwards every rank holds all of y
void comm_allgather_rows(alfib_ctx* c, double* y, const std::vector<int64_t>& start) {
  if (c->nranks <= 1) return;
  nccl_check(nccl().GroupStart(), "ncclGroupStart");
  for (int r = 0; r < c->nranks; ++r) {
    const size_t cnt = (size_t)(start[r + 1] - start[r]);
    if (cnt == 0) continue;
    nccl_check(nccl().Broadcast(y + start[r], y + start[r], cnt, ncclDouble, r, (ncclComm_t)c->comm, c->stream),
               "ncclBroadcast");
  }
  nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
}
