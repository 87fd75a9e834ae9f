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

#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "alfib_internal.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
  g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
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

// ---------------------------------------------------------------------------------------------
// Peer-memory exchanges over NVLink (replaces the NCCL calls above once the peers are mapped).
//
// Every rank owns a symmetric buffer [header | slot 0 | slot 1] that the other ranks of the box
// map through CUDA IPC.  A producer kernel (patch apply, row-sharded SpMV, coarse GEMV) writes
// this rank's part into slot (e & 1) of exchange number e; `peer_reduce_kernel` then publishes
// flag = e, waits until every peer's flag reaches e and *pulls*, for every index, the entries of
// the ranks whose range covers it — in rank order, so the result is bitwise identical on all
// ranks.  Only overlapping ranges cross NVLink (a patch apply moves its slab plus one macro
// layer, not the whole vector as an all-reduce does) and the whole exchange is one kernel.
// Two slots suffice without a second barrier: a rank overwrites slot (e & 1) only after it has
// seen every peer's flag e - 1, and a peer raises flag e - 1 only after it has finished reading
// exchange e - 2.  The exchange counter lives in device memory, so the sequence replays
// unchanged inside a CUDA graph.  Spin loops are bounded; a time-out raises an error flag
// instead of hanging the GPU.
namespace {

constexpr long long SPIN_LIMIT = 1ll << 22;      // ~ seconds of polling over NVLink

__global__ void peer_zero_kernel(PeerOut out, long long lo, long long hi) {
  double* y = resolve(out);
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = 0.0;
}

struct Ranges {
  long long lo[ALFIB_MAX_RANKS], hi[ALFIB_MAX_RANKS];
};

struct LocalGate {                     // in this rank's own memory: block 0 -> the other blocks
  unsigned long long go;
  long long lo[ALFIB_MAX_RANKS], hi[ALFIB_MAX_RANKS];
};

__global__ void __launch_bounds__(256) peer_reduce_kernel(long long n, int nranks, int rank,
                                                          double* const* __restrict__ peer_slot0, size_t stride,
                                                          const unsigned long long* __restrict__ epoch_ptr,
                                                          int hdr_slot, Ranges rg, double* __restrict__ y,
                                                          int* __restrict__ err, unsigned long long* epoch_rw,
                                                          unsigned int* __restrict__ done, LocalGate* local) {
  __shared__ long long s_lo[ALFIB_MAX_RANKS], s_hi[ALFIB_MAX_RANKS];
  const unsigned long long e = *epoch_ptr;
  const size_t hdr_doubles = ALFIB_SYM_HEADER_BYTES / sizeof(double);
  // Only block 0 talks to the peers' headers (thousands of blocks polling remote flags congest
  // NVLink with tiny requests): it publishes this rank's flag, waits for every peer, copies the
  // ranges into local memory and then releases the other blocks through a local flag.
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      __threadfence_system();                                // this rank's slot is complete
      SymHeader* mine = reinterpret_cast<SymHeader*>(peer_slot0[rank] - hdr_doubles);
      *reinterpret_cast<volatile unsigned long long*>(&mine->flag) = e;
    }
    if (threadIdx.x < nranks) {
      const int q = threadIdx.x;
      const SymHeader* h = reinterpret_cast<const SymHeader*>(peer_slot0[q] - hdr_doubles);
      if (q != rank) {
        long long spins = 0;
        while (*reinterpret_cast<const volatile unsigned long long*>(&h->flag) < e) {
          if (++spins > SPIN_LIMIT) {
            atomicExch(err, 1);
            break;
          }
        }
      }
      long long lo = rg.lo[q], hi = rg.hi[q];
      if (hdr_slot >= 0) {
        lo = *reinterpret_cast<const volatile long long*>(&h->lo[hdr_slot]);
        hi = *reinterpret_cast<const volatile long long*>(&h->hi[hdr_slot]);
      }
      local->lo[q] = lo;
      local->hi[q] = hi;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(&local->go) = e;
  }
  if (threadIdx.x == 0) {
    long long spins = 0;
    while (*reinterpret_cast<const volatile unsigned long long*>(&local->go) < e)
      if (++spins > SPIN_LIMIT * 16) {
        atomicExch(err, 1);
        break;
      }
  }
  __syncthreads();
  if (threadIdx.x < nranks) {
    s_lo[threadIdx.x] = *reinterpret_cast<const volatile long long*>(&local->lo[threadIdx.x]);
    s_hi[threadIdx.x] = *reinterpret_cast<const volatile long long*>(&local->hi[threadIdx.x]);
  }
  __threadfence_system();
  __syncthreads();
  const size_t off = (size_t)(e & 1ull) * stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int q = 0; q < nranks; ++q)
      if (i >= s_lo[q] && i < s_hi[q]) v += __ldcv(peer_slot0[q] + off + i);
    y[i] = v;
  }
  // the last block to finish advances the exchange counter (every block has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0;
      *epoch_rw = e + 1ull;
    }
  }
}

}  // namespace

// Channels of one halo in the mailbox arena (called by alfib_level_set_halo; host bookkeeping only, so it is
// harmless without peer memory): owner -> ghost needs room for this rank's ghosts, ghost -> owner for its send list.
static void mbox_reserve_channel(alfib_ctx* c, int ch, long long cap) {
  ALFIB_REQUIRE(ch >= 0 && ch < ALFIB_MBOX_CHANNELS, "mailbox channel out of range");
  MboxEntry& e = c->mbox[ch];
  if (e.data_off != 0 && e.cap >= cap) return;               // re-registration that still fits
  ALFIB_REQUIRE(!c->mbox_fixed, "alfib_level_set_halo after alfib_comm_peer_handle: the exchange buffer is already laid out");
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  e.cap = std::max<long long>(cap, 2);
  c->mbox_bytes = up(c->mbox_bytes);
  e.data_off = (long long)c->mbox_bytes + 256;               // + 256: 0 means "no channel"; made absolute at allocation
  c->mbox_bytes += up(2 * (size_t)e.cap * 2 * sizeof(double)) + 256;     // 2 slots x cap entries x 16 bytes (LL words)
  e.flags_off = (long long)c->mbox_bytes;
  c->mbox_bytes += up(ALFIB_MAX_RANKS * sizeof(unsigned long long));
}

void comm_mbox_reserve(alfib_ctx* c, Halo& H, int level, int which) {
  H.ch_update = (level * 2 + which) * 2;
  H.ch_reduce = H.ch_update + 1;
  mbox_reserve_channel(c, H.ch_update, H.recv_off.back());
  mbox_reserve_channel(c, H.ch_reduce, H.send_off.back());
}

// symmetric buffer sized for the largest level (and the coarse system); call after the levels exist
void comm_peer_alloc(alfib_ctx* c) {
  size_t maxn = 0;
  bool distributed = false;
  for (auto* L : c->levels)
    if (L) {
      maxn = std::max<size_t>(maxn, (size_t)L->n);
      distributed |= L->halo.on;
    }
  // distributed vectors: only the replicated level 0 goes through the slots, and the level sizes differ from rank
  // to rank — the slot stride must not (the peers address each other's slot 1 with their own stride)
  if (distributed && c->levels[0]) maxn = std::max<size_t>((size_t)c->levels[0]->n, 4096);
  maxn = (maxn + 31) & ~size_t(31);
  ALFIB_REQUIRE(maxn > 0, "create the levels before enabling peer memory");
  if (c->sym && c->sym_stride == maxn) return;
  ALFIB_REQUIRE(!c->peers_open, "peer memory already mapped");
  if (c->sym) cudaFree(c->sym);
  if (!c->mbox_fixed) {
    mbox_reserve_channel(c, ALFIB_MBOX_SMALL, (long long)ALFIB_MAX_RANKS * ALFIB_MBOX_NV);
    const size_t arena = ALFIB_SYM_HEADER_BYTES + 2 * maxn * sizeof(double);
    for (auto& e : c->mbox)
      if (e.data_off) {
        e.data_off += (long long)arena - 256;
        e.flags_off += (long long)arena;
      }
    c->mbox_fixed = true;
  }
  const size_t bytes = ALFIB_SYM_HEADER_BYTES + 2 * maxn * sizeof(double) + c->mbox_bytes + 512;
  CUDA_TRY(cudaMalloc(&c->sym, bytes));
  CUDA_TRY(cudaMemset(c->sym, 0, bytes));
  CUDA_TRY(cudaMemcpy(c->sym + ALFIB_MBOX_TABLE_OFF, c->mbox, sizeof(c->mbox), cudaMemcpyHostToDevice));
  c->sym_stride = maxn;
  c->d_epoch.alloc(1);
  const unsigned long long one = 1;
  CUDA_TRY(cudaMemcpy(c->d_epoch.p, &one, sizeof(one), cudaMemcpyHostToDevice));
  c->d_gate.alloc(sizeof(LocalGate) / sizeof(long long) + 1);
  CUDA_TRY(cudaMemset(c->d_gate.p, 0, sizeof(LocalGate)));
  c->d_comm_err.alloc(2);                                  // [0] time-out flag, [1] finished-block counter
  CUDA_TRY(cudaMemset(c->d_comm_err.p, 0, 2 * sizeof(int)));
  c->mbox_seq.alloc(ALFIB_MBOX_CHANNELS);
  CUDA_TRY(cudaMemset(c->mbox_seq.p, 0, ALFIB_MBOX_CHANNELS * sizeof(unsigned long long)));
  c->mbox_cnt.alloc(2 * ALFIB_MBOX_CHANNELS);
  CUDA_TRY(cudaMemset(c->mbox_cnt.p, 0, 2 * ALFIB_MBOX_CHANNELS * sizeof(unsigned int)));
}

void comm_peer_handle(alfib_ctx* c, void* out64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  comm_peer_alloc(c);
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c->sym));
  memcpy(out64, &h, sizeof(h));
}

void comm_peer_publish_ranges(alfib_ctx* c) {
  if (!c->sym) return;
  SymHeader hdr;
  memset(&hdr, 0, sizeof(hdr));
  for (int l = 0; l < ALFIB_MAX_LEVELS; ++l)
    if (c->levels[l])
      for (int w = 0; w < 2; ++w) {
        hdr.lo[2 * l + w] = c->levels[l]->ps[w].lo;
        hdr.hi[2 * l + w] = c->levels[l]->ps[w].hi;
      }
  // the flag (first 8 bytes) is owned by the kernels: copy only the ranges
  CUDA_TRY(cudaMemcpy(c->sym + offsetof(SymHeader, lo), &hdr.lo, sizeof(hdr) - offsetof(SymHeader, lo),
                      cudaMemcpyHostToDevice));
}

void comm_peer_open(alfib_ctx* c, const void* handles) {
  ALFIB_REQUIRE(c->nranks > 1 && c->nranks <= ALFIB_MAX_RANKS, "peer memory needs 2..8 ranks");
  ALFIB_REQUIRE(c->sym, "alfib_comm_peer_handle first");
  ALFIB_REQUIRE(!c->peers_open, "peer memory already mapped");
  std::vector<double*> slot0(c->nranks);
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int q = 0; q < c->nranks; ++q) {
    void* base = c->sym;
    if (q != c->rank) {
      CUDA_TRY(cudaIpcOpenMemHandle(&base, h[q], cudaIpcMemLazyEnablePeerAccess));
      c->peer_ptr[q] = base;
    }
    slot0[q] = reinterpret_cast<double*>(static_cast<unsigned char*>(base) + ALFIB_SYM_HEADER_BYTES);
  }
  c->d_peer_slot.alloc(c->nranks);
  CUDA_TRY(cudaMemcpy(c->d_peer_slot.p, slot0.data(), sizeof(double*) * c->nranks, cudaMemcpyHostToDevice));
  // the peers' channel tables (written before they produced their handles, i.e. before the host-side all-gather)
  for (int q = 0; q < c->nranks; ++q) {
    c->mbox_peer[q].assign(ALFIB_MBOX_CHANNELS, MboxEntry{0, 0, 0});
    const unsigned char* base = q == c->rank ? c->sym : static_cast<const unsigned char*>(c->peer_ptr[q]);
    CUDA_TRY(cudaMemcpy(c->mbox_peer[q].data(), base + ALFIB_MBOX_TABLE_OFF, sizeof(MboxEntry) * ALFIB_MBOX_CHANNELS,
                        cudaMemcpyDeviceToHost));
  }
  comm_peer_publish_ranges(c);
  CUDA_TRY(cudaDeviceSynchronize());
  c->peers_open = true;
}

void comm_peer_close(alfib_ctx* c) {
  for (int q = 0; q < ALFIB_MAX_RANKS; ++q)
    if (c->peer_ptr[q]) {
      cudaIpcCloseMemHandle(c->peer_ptr[q]);
      c->peer_ptr[q] = nullptr;
    }
  c->peers_open = false;
  if (c->sym) cudaFree(c->sym);
  c->sym = nullptr;
}

PeerOut comm_peer_out(alfib_ctx* c) {
  return PeerOut{reinterpret_cast<double*>(c->sym + ALFIB_SYM_HEADER_BYTES), c->sym_stride, c->d_epoch.p};
}

void comm_peer_zero(alfib_ctx* c, long long lo, long long hi) {
  if (hi <= lo) return;
  const int blocks = (int)std::min<long long>((hi - lo + 255) / 256, 4 * c->num_sms);
  peer_zero_kernel<<<blocks, 256, 0, c->stream>>>(comm_peer_out(c), lo, hi);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

void comm_peer_reduce(alfib_ctx* c, int64_t n, int hdr_slot, const long long* lo, const long long* hi, double* y) {
  Ranges rg;
  for (int q = 0; q < ALFIB_MAX_RANKS; ++q) {
    rg.lo[q] = (lo && q < c->nranks) ? lo[q] : 0;
    rg.hi[q] = (hi && q < c->nranks) ? hi[q] : 0;
  }
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 8 * c->num_sms);
  peer_reduce_kernel<<<std::max(blocks, 1), 256, 0, c->stream>>>(
      n, c->nranks, c->rank, c->d_peer_slot.p, c->sym_stride, c->d_epoch.p, hdr_slot, rg, y, c->d_comm_err.p,
      c->d_epoch.p, reinterpret_cast<unsigned int*>(c->d_comm_err.p + 1),
      reinterpret_cast<LocalGate*>(c->d_gate.p));
  c->launches += 1;
  CUDA_TRY(cudaGetLastError());
}

int comm_peer_error(alfib_ctx* c) {
  if (!c->d_comm_err.p) return 0;
  int e = 0;
  cudaMemcpy(&e, c->d_comm_err.p, sizeof(int), cudaMemcpyDeviceToHost);
  if (e) cudaMemset(c->d_comm_err.p, 0, sizeof(int));     // reported once; the next exchange starts clean
  return e;
}

// ---------------------------------------------------------------------------------------------
// Distributed level vectors (alfib_level_set_halo): the PetscSF bcast / reduce of the reference's
// parallel PCPATCH and MatMult, between neighbouring ranks only.  Bytes moved per exchange and rank:
// 8 x (ghost entries), e.g. <= 0.9 MB on cfg5's finest level at 8 ranks against the 11.7 MB
// all-reduce of the replicated design.  Two transports, the same packed layout:
//   * NCCL: pack -> one group of ncclSend/ncclRecv per neighbour -> unpack;
//   * NVLink peer memory (alfib_comm_peer_open): pack into this rank's symmetric slot, then ONE kernel
//     that runs the flag protocol of peer_reduce_kernel above (publish e, wait for every rank's flag,
//     advance the exchange counter) and pulls the neighbours' packed entries straight into place — the
//     transfer and the unpack (or the fixed-order sum) are the same loads.
// Everything is enqueued on the ctx stream and replays inside the cycle's CUDA graph.
namespace {

// buf[i] = x[idx[i]]; idx == nullptr: buf[i] = x[i]
__global__ void halo_pack_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ x, PeerOut out) {
  double* __restrict__ buf = resolve(out);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    buf[i] = idx ? x[idx[i]] : x[i];
}

// x[idx[i]] = buf[i]; the ghost positions of one layout are distinct
__global__ void halo_unpack_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ buf,
                                   double* __restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[idx[i]] = buf[i];
}

// y[red_dof[i]] += sum_k buf[red_src[k]], k ascending = peers in ascending rank order (reproducible)
__global__ void halo_sum_kernel(int n_red, const int32_t* __restrict__ red_ptr, const int32_t* __restrict__ red_dof,
                                const int32_t* __restrict__ red_src, const double* __restrict__ buf,
                                double* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_red; i += gridDim.x * blockDim.x) {
    double v = y[red_dof[i]];
    for (int k = red_ptr[i]; k < red_ptr[i + 1]; ++k) v += buf[red_src[k]];
    y[red_dof[i]] = v;
  }
}

inline int halo_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), 4 * 148)); }

// one NCCL group: this rank sends `out` (grouped by peer with offsets out_off) and receives `in`
void halo_sendrecv(alfib_ctx* c, const Halo& H, const double* out, const std::vector<int64_t>& out_off, double* in,
                   const std::vector<int64_t>& in_off) {
  nccl_check(nccl().GroupStart(), "ncclGroupStart");
  for (size_t p = 0; p < H.peers.size(); ++p) {
    const size_t ns = (size_t)(out_off[p + 1] - out_off[p]), nr = (size_t)(in_off[p + 1] - in_off[p]);
    if (ns) nccl_check(nccl().Send(out + out_off[p], ns, ncclDouble, H.peers[p], (ncclComm_t)c->comm, c->stream), "ncclSend");
    if (nr) nccl_check(nccl().Recv(in + in_off[p], nr, ncclDouble, H.peers[p], (ncclComm_t)c->comm, c->stream), "ncclRecv");
  }
  nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
}

// ---- the flag protocol of peer_reduce_kernel as two device functions ---------------------------
// begin: this rank's slot (e & 1) is complete (written by the preceding kernel on the stream); block 0
// publishes flag = e, waits until every rank has published e and releases the other blocks.
__device__ __forceinline__ void peer_exchange_begin(unsigned long long e, int nranks, int rank,
                                                    double* const* __restrict__ peer_slot0, int* __restrict__ err,
                                                    LocalGate* local) {
  const size_t hdr_doubles = ALFIB_SYM_HEADER_BYTES / sizeof(double);
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      __threadfence_system();
      SymHeader* mine = reinterpret_cast<SymHeader*>(peer_slot0[rank] - hdr_doubles);
      *reinterpret_cast<volatile unsigned long long*>(&mine->flag) = e;
    }
    if (threadIdx.x < nranks && threadIdx.x != rank) {
      const SymHeader* h = reinterpret_cast<const SymHeader*>(peer_slot0[threadIdx.x] - hdr_doubles);
      long long spins = 0;
      while (*reinterpret_cast<const volatile unsigned long long*>(&h->flag) < e) {
        if (++spins > SPIN_LIMIT) {
          atomicExch(err, 1);
          break;
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(&local->go) = e;
  }
  if (threadIdx.x == 0) {
    long long spins = 0;
    while (*reinterpret_cast<const volatile unsigned long long*>(&local->go) < e)
      if (++spins > SPIN_LIMIT * 16) {
        atomicExch(err, 1);
        break;
      }
  }
  __threadfence_system();
  __syncthreads();
}

// end: the last block to finish advances the exchange counter (every block has read it by then)
__device__ __forceinline__ void peer_exchange_end(unsigned long long e, unsigned long long* epoch_rw, unsigned int* done) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0;
      *epoch_rw = e + 1ull;
    }
  }
}

__device__ __forceinline__ int segment_of(const HaloPeers& hp, long long pos) {
  int p = 0;
  while (p + 1 < hp.npeers && pos >= hp.mine_off[p + 1]) ++p;
  return p;
}

// owner -> ghost over peer memory: x[recv_idx[i]] = (packed send buffer of the owner)[...]
__global__ void __launch_bounds__(256) peer_halo_update_kernel(long long nr, HaloPeers hp, int nranks, int rank,
                                                               double* const* __restrict__ peer_slot0, size_t stride,
                                                               const int32_t* __restrict__ recv_idx, double* __restrict__ x,
                                                               int* __restrict__ err, unsigned long long* epoch_rw,
                                                               unsigned int* __restrict__ done, LocalGate* local) {
  const unsigned long long e = *epoch_rw;
  peer_exchange_begin(e, nranks, rank, peer_slot0, err, local);
  const size_t off = (size_t)(e & 1ull) * stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (long long)gridDim.x * blockDim.x) {
    const int p = segment_of(hp, i);
    x[recv_idx[i]] = __ldcv(peer_slot0[hp.peers[p]] + off + hp.theirs_off[p] + (i - hp.mine_off[p]));
  }
  peer_exchange_end(e, epoch_rw, done);
}

// ghost -> owner over peer memory: y[d] += sum over the peers holding d as a ghost, ascending rank order
__global__ void __launch_bounds__(256) peer_halo_sum_kernel(int n_red, HaloPeers hp, int nranks, int rank,
                                                            double* const* __restrict__ peer_slot0, size_t stride,
                                                            const int32_t* __restrict__ red_ptr,
                                                            const int32_t* __restrict__ red_dof,
                                                            const int32_t* __restrict__ red_src, double* __restrict__ y,
                                                            int* __restrict__ err, unsigned long long* epoch_rw,
                                                            unsigned int* __restrict__ done, LocalGate* local) {
  const unsigned long long e = *epoch_rw;
  peer_exchange_begin(e, nranks, rank, peer_slot0, err, local);
  const size_t off = (size_t)(e & 1ull) * stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_red; i += gridDim.x * blockDim.x) {
    double v = y[red_dof[i]];
    for (int k = red_ptr[i]; k < red_ptr[i + 1]; ++k) {
      const long long pos = red_src[k];
      const int p = segment_of(hp, pos);
      v += __ldcv(peer_slot0[hp.peers[p]] + off + hp.theirs_off[p] + (pos - hp.mine_off[p]));
    }
    y[red_dof[i]] = v;
  }
  peer_exchange_end(e, epoch_rw, done);
}

// v[j] = sum over ranks (rank order) of their packed v[j]; optional sqrt / reciprocal of v[0]
__global__ void peer_small_sum_kernel(int nv, int nranks, int rank, double* const* __restrict__ peer_slot0, size_t stride,
                                      double* __restrict__ v, int sqrt_mode, double* __restrict__ inv,
                                      int* __restrict__ err, unsigned long long* epoch_rw, unsigned int* __restrict__ done,
                                      LocalGate* local) {
  const unsigned long long e = *epoch_rw;
  peer_exchange_begin(e, nranks, rank, peer_slot0, err, local);
  const size_t off = (size_t)(e & 1ull) * stride;
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < nranks; ++q) s += __ldcv(peer_slot0[q] + off + j);
    if (sqrt_mode) {
      s = sqrt(s);
      if (inv) inv[j] = s > 0.0 ? 1.0 / s : 0.0;
    }
    v[j] = s;
  }
  peer_exchange_end(e, epoch_rw, done);
}

__global__ void sqrt_inv_kernel(double* __restrict__ out, double* __restrict__ inv) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double v = sqrt(out[0]);
  out[0] = v;
  if (inv) inv[0] = v > 0.0 ? 1.0 / v : 0.0;
}

HaloPeers make_peers(const Halo& H, const std::vector<int64_t>& mine, const std::vector<int64_t>& theirs) {
  HaloPeers hp;
  memset(&hp, 0, sizeof(hp));
  hp.npeers = (int)H.peers.size();
  for (int p = 0; p < hp.npeers; ++p) {
    hp.peers[p] = H.peers[p];
    hp.mine_off[p] = mine[p];
    hp.theirs_off[p] = theirs[p];
  }
  hp.mine_off[hp.npeers] = mine[hp.npeers];
  return hp;
}

inline bool use_peers(const alfib_ctx* c, const Halo& H) { return c->nranks > 1 && c->peers_open && H.has_peer_off; }

// ---- mailbox ("push") transport ---------------------------------------------------------------------------------
// One kernel per exchange and no fences: every double travels as one 16-byte word {lo32, k, hi32, k} carrying the
// exchange number k of its channel in both 8-byte halves (the LL protocol of NCCL's low-latency path, here with FP64
// payload).  Exchange k (k = completed exchanges + 1, kept in device memory so the sequence replays inside a CUDA
// graph) uses slot k & 1 of the RECEIVER's channel:
//   push    every entry of this rank's outgoing list is written straight into its place in the neighbour's slot with
//           one 128-bit NVLink peer store — pack, transfer and "ready" signal are the same instruction;
//   pull    every incoming entry is polled in LOCAL memory until both halves carry k, then the ghosts are written /
//           the owned interface dofs gather-sum their contributions in ascending rank order (reproducible).
// A word of exchange k - 2 in the same slot carries k - 2 and cannot be mistaken for k.  Two slots suffice against
// overwriting: a neighbour's push k + 2 follows its pull k + 1, i.e. this rank's push k + 1, which this rank's stream
// issues only after its kernel k has finished reading slot k & 1 (the neighbour relation of a channel is symmetric:
// every listed peer is written to and polled, also with an empty segment... an empty segment needs no word).  All
// blocks of a kernel are resident (grid <= number of SMs), so blocks polling cannot starve blocks that still have
// to push.  A time-out (20 s on %globaltimer) raises the error flag alfib_synchronize / alfib_cycle_apply report.
struct MboxPeers {
  int npeers;
  int peers[ALFIB_MAX_RANKS];
  long long mine_off[ALFIB_MAX_RANKS + 1];        // segments of this rank's outgoing list
  double* data[ALFIB_MAX_RANKS];                  // slot 0 of the channel in the neighbour's memory
  long long cap[ALFIB_MAX_RANKS];                 // its slot size (entries; an entry is 16 bytes)
  long long theirs_off[ALFIB_MAX_RANKS];          // where this rank's segment starts in the neighbour's incoming list
  unsigned long long* flag[ALFIB_MAX_RANKS];      // (unused by the LL kernels)
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ int mbox_segment(const MboxPeers& mp, long long pos) {
  int p = 0;
  while (p + 1 < mp.npeers && pos >= mp.mine_off[p + 1]) ++p;
  return p;
}

__device__ __forceinline__ void ll_store(double* slot_entry /* 16-byte entry */, double v, unsigned int k) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot_entry), "r"((unsigned int)bits), "r"(k),
               "r"((unsigned int)(bits >> 32)), "r"(k)
               : "memory");
}

__device__ __forceinline__ double ll_load(const double* slot_entry, unsigned int k, int* __restrict__ err) {
  unsigned int lo, f0, hi, f1;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(slot_entry) : "memory");
    if (f0 == k && f1 == k) break;
    if (t0 == 0) t0 = global_ns();
    else if (global_ns() - t0 > 20000000000ull) {   // 20 s: the ranks enter the first exchange after rank-local setup phases
      atomicExch(err, 1);
      break;
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

__device__ __forceinline__ void mbox_finish(unsigned long long k, unsigned long long* seq, unsigned int* finished) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(finished, 1u) == gridDim.x - 1) {
      *finished = 0;
      *seq = k;
    }
  }
}

// owner -> ghost
__global__ void __launch_bounds__(256) mbox_update_kernel(long long ns, long long nr, MboxPeers mp,
                                                          const int32_t* __restrict__ send_idx,
                                                          const int32_t* __restrict__ recv_idx, double* __restrict__ x,
                                                          const double* my_data, long long my_cap,
                                                          const unsigned long long* my_flags, unsigned long long* seq,
                                                          unsigned int* cnt, int* __restrict__ err) {
  const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(seq) + 1ull;
  const long long slot = (long long)(k & 1ull);
  const long long stride = (long long)gridDim.x * blockDim.x, t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t; i < ns; i += stride) {
    const int p = mbox_segment(mp, i);
    ll_store(mp.data[p] + 2 * (slot * mp.cap[p] + mp.theirs_off[p] + (i - mp.mine_off[p])), x[send_idx[i]], (unsigned int)k);
  }
  const double* in = my_data + 2 * slot * my_cap;
  for (long long i = t; i < nr; i += stride) x[recv_idx[i]] = ll_load(in + 2 * i, (unsigned int)k, err);
  mbox_finish(k, seq, cnt + 1);
}

// ghost -> owner: the ghost entries are pushed (and cleared), the owned interface dofs sum what arrives
__global__ void __launch_bounds__(256) mbox_reduce_kernel(long long nr, int n_red, MboxPeers mp,
                                                          const int32_t* __restrict__ recv_idx,
                                                          const int32_t* __restrict__ red_ptr,
                                                          const int32_t* __restrict__ red_dof,
                                                          const int32_t* __restrict__ red_src, double* __restrict__ y,
                                                          const double* my_data, long long my_cap,
                                                          const unsigned long long* my_flags, unsigned long long* seq,
                                                          unsigned int* cnt, int* __restrict__ err) {
  const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(seq) + 1ull;
  const long long slot = (long long)(k & 1ull);
  const long long stride = (long long)gridDim.x * blockDim.x, t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t; i < nr; i += stride) {
    const int p = mbox_segment(mp, i);
    const int g = recv_idx[i];
    ll_store(mp.data[p] + 2 * (slot * mp.cap[p] + mp.theirs_off[p] + (i - mp.mine_off[p])), y[g], (unsigned int)k);
    y[g] = 0.0;
  }
  const double* in = my_data + 2 * slot * my_cap;
  for (long long i = t; i < n_red; i += stride) {
    double v = y[red_dof[i]];
    for (int j = red_ptr[i]; j < red_ptr[i + 1]; ++j) v += ll_load(in + 2 * (long long)red_src[j], (unsigned int)k, err);
    y[red_dof[i]] = v;
  }
  mbox_finish(k, seq, cnt + 1);
}

// v[j] summed over all ranks in rank order (one block); optional sqrt / reciprocal of the result
__global__ void __launch_bounds__(64) mbox_small_sum_kernel(int nv, int nranks, int rank, MboxPeers mp, double* __restrict__ v,
                                                            int sqrt_mode, double* __restrict__ inv, const double* my_data,
                                                            long long my_cap, const unsigned long long* my_flags,
                                                            unsigned long long* seq, unsigned int* cnt, int* __restrict__ err) {
  const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(seq) + 1ull;
  const long long slot = (long long)(k & 1ull);
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const double mine = v[j];
    for (int p = 0; p < mp.npeers; ++p) ll_store(mp.data[p] + 2 * (slot * mp.cap[p] + (long long)rank * ALFIB_MBOX_NV + j), mine, (unsigned int)k);
  }
  const double* in = my_data + 2 * slot * my_cap;
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < nranks; ++q) s += (q == rank) ? v[j] : ll_load(in + 2 * ((long long)q * ALFIB_MBOX_NV + j), (unsigned int)k, err);
    if (sqrt_mode) {
      s = sqrt(s);
      if (inv) inv[j] = s > 0.0 ? 1.0 / s : 0.0;
    }
    v[j] = s;
  }
  mbox_finish(k, seq, cnt + 1);
}

// The same with the second pass of a two-pass dot folded in: v[j] is first formed as the fixed-order sum of the
// nparts partial sums partial[j * nparts + b] (exactly finalize_kernel of krylov.cu: lanes stride over b, shuffle
// tree), one warp per j — one launch less per dot / norm of the distributed FGMRES.
__global__ void __launch_bounds__(256) mbox_small_sum_partials_kernel(int nv, int nparts, const double* __restrict__ partial,
                                                                      int nranks, int rank, MboxPeers mp, double* __restrict__ v,
                                                                      int sqrt_mode, double* __restrict__ inv, const double* my_data,
                                                                      long long my_cap, unsigned long long* seq, unsigned int* cnt,
                                                                      int* __restrict__ err) {
  const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(seq) + 1ull;
  const long long slot = (long long)(k & 1ull);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const double* in = my_data + 2 * slot * my_cap;
  for (int j = warp; j < nv; j += nwarp) {
    // eight loads in flight per lane (the plain loop is a chain of 19 dependent-latency loads for 592 partials); added in
    // the order of that loop, and x + 0.0 == x here, so the bits are finalize_kernel's
    double mine = 0.0;
    for (int b0 = lane; b0 < nparts; b0 += 32 * 8) {
      double a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = (b0 + 32 * u < nparts) ? partial[(long long)j * nparts + b0 + 32 * u] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) mine += a[u];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if (lane == 0) {
      for (int p = 0; p < mp.npeers; ++p)
        ll_store(mp.data[p] + 2 * (slot * mp.cap[p] + (long long)rank * ALFIB_MBOX_NV + j), mine, (unsigned int)k);
      double s = 0.0;
      for (int q = 0; q < nranks; ++q) s += (q == rank) ? mine : ll_load(in + 2 * ((long long)q * ALFIB_MBOX_NV + j), (unsigned int)k, err);
      if (sqrt_mode) {
        s = sqrt(s);
        if (inv) inv[j] = s > 0.0 ? 1.0 / s : 0.0;
      }
      v[j] = s;
    }
  }
  mbox_finish(k, seq, cnt + 1);
}

// the neighbours of one channel as the kernels take them; `mine` = offsets of the outgoing list, `theirs` = where
// each neighbour expects this rank's segment
MboxPeers mbox_peers(alfib_ctx* c, int ch, const std::vector<int>& peers, const std::vector<int64_t>& mine,
                     const std::vector<int64_t>& theirs) {
  MboxPeers mp;
  memset(&mp, 0, sizeof(mp));
  mp.npeers = (int)peers.size();
  for (int p = 0; p < mp.npeers; ++p) {
    const int q = peers[p];
    const MboxEntry& e = c->mbox_peer[q][ch];
    if (e.data_off == 0) throw DeviceError{ALFIB_EINVAL, "rank " + std::to_string(q) + " has no mailbox channel " + std::to_string(ch)};
    unsigned char* base = static_cast<unsigned char*>(c->peer_ptr[q]);
    mp.peers[p] = q;
    mp.mine_off[p] = mine[p];
    mp.theirs_off[p] = theirs[p];
    mp.data[p] = reinterpret_cast<double*>(base + e.data_off);
    mp.cap[p] = e.cap;
    mp.flag[p] = reinterpret_cast<unsigned long long*>(base + e.flags_off) + c->rank;
    if (theirs[p] + (mine[p + 1] - mine[p]) > e.cap)
      throw DeviceError{ALFIB_EINVAL, "mailbox segment does not fit the neighbour's channel"};
  }
  mp.mine_off[mp.npeers] = mine[mp.npeers];
  return mp;
}

inline int mbox_grid(alfib_ctx* c, int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 512), c->num_sms)); }
inline bool use_mbox(const alfib_ctx* c, const Halo& H) {
  return c->nranks > 1 && c->peers_open && H.has_peer_off && H.ch_update >= 0 && c->mbox_fixed &&
         !std::getenv("ALFIB_MBOX_OFF");      // ALFIB_MBOX_OFF=1: the pull transport (pack + flag barrier + peer loads)
}

}  // namespace

// owner -> ghost: every ghost entry of x takes its owner's value
static bool debug_skip_exchange() {
  static const bool skip = std::getenv("ALFIB_DEBUG_SKIP_EXCHANGE") != nullptr;   // timing experiments: WRONG results
  return skip;
}

void halo_update(alfib_ctx* c, Halo& H, double* x, int level) {
  if (!H.on || c->nranks <= 1) return;
  if (debug_skip_exchange()) return;
  ScopedEvent ev(c, ALFIB_EV_HALO, level);
  const int64_t ns = H.send_off.back(), nr = H.recv_off.back();
  if (use_mbox(c, H)) {
    if (H.peers.empty()) return;
    const int ch = H.ch_update;
    const MboxEntry& me = c->mbox[ch];
    mbox_update_kernel<<<mbox_grid(c, std::max(ns, nr)), 256, 0, c->stream>>>(
        ns, nr, mbox_peers(c, ch, H.peers, H.send_off, H.peer_recv_off), H.send_idx.p, H.recv_idx.p, x,
        reinterpret_cast<const double*>(c->sym + me.data_off), me.cap,
        reinterpret_cast<const unsigned long long*>(c->sym + me.flags_off), c->mbox_seq.p + ch, c->mbox_cnt.p + 2 * ch,
        c->d_comm_err.p);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return;
  }
  if (use_peers(c, H)) {
    // every rank takes part in every exchange (the flags count exchanges), also one without neighbours here
    ALFIB_REQUIRE((size_t)ns <= c->sym_stride, "packed halo does not fit the symmetric buffer");
    halo_pack_kernel<<<halo_grid(ns), 256, 0, c->stream>>>(ns, H.send_idx.p, x, comm_peer_out(c));
    peer_halo_update_kernel<<<halo_grid(nr), 256, 0, c->stream>>>(
        nr, make_peers(H, H.recv_off, H.peer_send_off), c->nranks, c->rank, c->d_peer_slot.p, c->sym_stride, H.recv_idx.p, x,
        c->d_comm_err.p, c->d_epoch.p, reinterpret_cast<unsigned int*>(c->d_comm_err.p + 1),
        reinterpret_cast<LocalGate*>(c->d_gate.p));
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return;
  }
  if (H.peers.empty()) return;
  if (ns) {
    halo_pack_kernel<<<halo_grid(ns), 256, 0, c->stream>>>(ns, H.send_idx.p, x, plain_out(H.sbuf.p));
    c->launches++;
  }
  halo_sendrecv(c, H, H.sbuf.p, H.send_off, H.rbuf.p, H.recv_off);
  if (nr) {
    halo_unpack_kernel<<<halo_grid(nr), 256, 0, c->stream>>>(nr, H.recv_idx.p, H.rbuf.p, x);
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
}

// ghost -> owner: the ghost entries of y are added to their owners' entries (peers in ascending rank order: a
// fixed summation order) and cleared
void halo_reduce(alfib_ctx* c, Halo& H, double* y, int level) {
  if (!H.on) return;
  ScopedEvent ev(c, ALFIB_EV_HALO, level);
  if (c->nranks > 1 && !debug_skip_exchange()) {
    const int64_t ns = H.send_off.back(), nr = H.recv_off.back();
    if (use_mbox(c, H)) {
      if (!H.peers.empty()) {
        const int ch = H.ch_reduce;
        const MboxEntry& me = c->mbox[ch];
        mbox_reduce_kernel<<<mbox_grid(c, std::max<int64_t>(nr, H.n_red)), 256, 0, c->stream>>>(
            nr, H.n_red, mbox_peers(c, ch, H.peers, H.recv_off, H.peer_send_off), H.recv_idx.p, H.red_ptr.p, H.red_dof.p,
            H.red_src.p, y, reinterpret_cast<const double*>(c->sym + me.data_off), me.cap,
            reinterpret_cast<const unsigned long long*>(c->sym + me.flags_off), c->mbox_seq.p + ch, c->mbox_cnt.p + 2 * ch,
            c->d_comm_err.p);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        return;                                  // the kernel has cleared the ghosts it pushed (all of them)
      }
    } else if (use_peers(c, H)) {
      ALFIB_REQUIRE((size_t)nr <= c->sym_stride, "packed halo does not fit the symmetric buffer");
      halo_pack_kernel<<<halo_grid(nr), 256, 0, c->stream>>>(nr, H.recv_idx.p, y, comm_peer_out(c));
      peer_halo_sum_kernel<<<halo_grid(H.n_red), 256, 0, c->stream>>>(
          H.n_red, make_peers(H, H.send_off, H.peer_recv_off), c->nranks, c->rank, c->d_peer_slot.p, c->sym_stride,
          H.red_ptr.p, H.red_dof.p, H.red_src.p, y, c->d_comm_err.p, c->d_epoch.p,
          reinterpret_cast<unsigned int*>(c->d_comm_err.p + 1), reinterpret_cast<LocalGate*>(c->d_gate.p));
      c->launches += 2;
    } else if (!H.peers.empty()) {
      if (nr) {
        halo_pack_kernel<<<halo_grid(nr), 256, 0, c->stream>>>(nr, H.recv_idx.p, y, plain_out(H.rbuf.p));
        c->launches++;
      }
      halo_sendrecv(c, H, H.rbuf.p, H.recv_off, H.sbuf.p, H.send_off);
      if (ns && H.n_red) {
        halo_sum_kernel<<<halo_grid(H.n_red), 256, 0, c->stream>>>(H.n_red, H.red_ptr.p, H.red_dof.p, H.red_src.p,
                                                                   H.sbuf.p, y);
        c->launches++;
      }
    }
    CUDA_TRY(cudaGetLastError());
  }
  if (H.n_local > H.n_owned)
    CUDA_TRY(cudaMemsetAsync(y + H.n_owned, 0, sizeof(double) * (size_t)(H.n_local - H.n_owned), c->stream));
}

// out[j] = (sqrt of) the sum over all ranks of the fixed-order sum of partial[j * nparts .. + nparts); returns false if
// the mailbox transport is not available (the caller then finalises and all-reduces separately)
bool comm_small_allreduce_partials(alfib_ctx* c, const double* partial, int nparts, double* v, int nv, int sqrt_mode, double* inv) {
  if (!(c->nranks > 1 && c->peers_open && c->mbox_fixed && nv <= ALFIB_MBOX_NV) || std::getenv("ALFIB_MBOX_OFF") ||
      debug_skip_exchange())
    return false;
  const int ch = ALFIB_MBOX_SMALL;
  const MboxEntry& me = c->mbox[ch];
  std::vector<int> peers;
  for (int q = 0; q < c->nranks; ++q)
    if (q != c->rank) peers.push_back(q);
  std::vector<int64_t> zeros(peers.size() + 1, 0);
  const int threads = 32 * std::min(nv, 8);
  mbox_small_sum_partials_kernel<<<1, threads, 0, c->stream>>>(nv, nparts, partial, c->nranks, c->rank,
                                                               mbox_peers(c, ch, peers, zeros, zeros), v, sqrt_mode, inv,
                                                               reinterpret_cast<const double*>(c->sym + me.data_off), me.cap,
                                                               c->mbox_seq.p + ch, c->mbox_cnt.p + 2 * ch, c->d_comm_err.p);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return true;
}

void comm_small_allreduce(alfib_ctx* c, double* v, int nv, int sqrt_mode, double* inv) {
  if (debug_skip_exchange()) {
    if (sqrt_mode) {
      sqrt_inv_kernel<<<1, 32, 0, c->stream>>>(v, inv);
      c->launches++;
    }
    return;
  }
  if (c->nranks > 1 && c->peers_open && c->mbox_fixed && nv <= ALFIB_MBOX_NV && !std::getenv("ALFIB_MBOX_OFF")) {
    const int ch = ALFIB_MBOX_SMALL;
    const MboxEntry& me = c->mbox[ch];
    std::vector<int> peers;
    std::vector<int64_t> zeros;
    for (int q = 0; q < c->nranks; ++q)
      if (q != c->rank) peers.push_back(q);
    zeros.assign(peers.size() + 1, 0);
    mbox_small_sum_kernel<<<1, 64, 0, c->stream>>>(nv, c->nranks, c->rank, mbox_peers(c, ch, peers, zeros, zeros), v, sqrt_mode, inv,
                                                   reinterpret_cast<const double*>(c->sym + me.data_off), me.cap,
                                                   reinterpret_cast<const unsigned long long*>(c->sym + me.flags_off),
                                                   c->mbox_seq.p + ch, c->mbox_cnt.p + 2 * ch, c->d_comm_err.p);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return;
  }
  if (c->nranks > 1 && c->peers_open) {
    ALFIB_REQUIRE((size_t)nv <= c->sym_stride, "too many values for the symmetric buffer");
    halo_pack_kernel<<<1, 64, 0, c->stream>>>(nv, nullptr, v, comm_peer_out(c));
    peer_small_sum_kernel<<<1, 64, 0, c->stream>>>(nv, c->nranks, c->rank, c->d_peer_slot.p, c->sym_stride, v, sqrt_mode, inv,
                                                   c->d_comm_err.p, c->d_epoch.p,
                                                   reinterpret_cast<unsigned int*>(c->d_comm_err.p + 1),
                                                   reinterpret_cast<LocalGate*>(c->d_gate.p));
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return;
  }
  comm_allreduce_sum(c, v, (size_t)nv);
  if (sqrt_mode) {
    sqrt_inv_kernel<<<1, 32, 0, c->stream>>>(v, inv);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
}
