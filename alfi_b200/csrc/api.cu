// C-ABI entry points of libalfib (see include/alfib.h for the contract of each function and the
// reference interface it replaces).  Host-side bookkeeping only; kernels live in the other files.
#include <algorithm>
#include <cstring>
#include <exception>
#include <numeric>

#include "alfib_internal.h"

namespace {

template <class F>
int guarded(alfib_ctx* c, F&& f) {
  if (!c) return ALFIB_EINVAL;
  try {
    CUDA_TRY(cudaSetDevice(c->device));
    f();
    if (c->sync_always) CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ALFIB_OK;
  } catch (const DeviceError& e) {
    c->err = e.msg;
    return e.code;
  } catch (const std::bad_alloc&) {
    c->err = "host allocation failed";
    return ALFIB_ENOMEM;
  } catch (const std::exception& e) {
    c->err = e.what();
    return ALFIB_EINVAL;
  }
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

void release_patchset(PatchSet& ps) {
  ps.off.release();
  ps.soff.release();
  ps.dofs.release();
  ps.sorted.release();
  ps.sperm.release();
  ps.forder.release();
  ps.work.release();
  ps.store_buf.release();
  ps.cond.release();
  ps.stage_work.release();
  ps.stage_rows.release();
  ps.corr_off.release();
  ps.corr_rows.release();
  ps.corr_cols.release();
  ps.corr_vals.release();
}

// equal split of the block rows across ranks (contiguous ranges)
void set_row_partition(alfib_ctx* c, Level& L) {
  L.row_start.assign(c->nranks + 1, 0);
  L.dof_start.assign(c->nranks + 1, 0);
  for (int r = 0; r <= c->nranks; ++r) {
    L.row_start[r] = (int64_t)L.n_nodes * r / c->nranks;
    L.dof_start[r] = L.row_start[r] * L.bs;
  }
}

Level& get_level(alfib_ctx* c, int level) {
  ALFIB_REQUIRE(level >= 0 && level < ALFIB_MAX_LEVELS && c->levels[level], "no such level");
  return *c->levels[level];
}

// Input vector: device pointers pass through, host pointers are staged.
const double* in_vec(alfib_ctx* c, const double* p, size_t n, DBuf<double>& stage) {
  ALFIB_REQUIRE(p != nullptr, "null vector");
  if (is_device_ptr(p)) return p;
  if (stage.n < n) stage.alloc(n);
  CUDA_TRY(cudaMemcpyAsync(stage.p, p, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return stage.p;
}

struct OutVec {                      // output vector: compute into dev, copy back if host
  alfib_ctx* c;
  double* user;
  double* dev;
  size_t n;
  bool host;
  OutVec(alfib_ctx* ctx, double* p, size_t count) : c(ctx), user(p), n(count) {
    ALFIB_REQUIRE(p != nullptr, "null vector");
    host = !is_device_ptr(p);
    if (host) {
      if (c->stage_out.n < n) c->stage_out.alloc(n);
      dev = c->stage_out.p;
    } else {
      dev = p;
    }
  }
  void load() {                      // for in/out vectors
    if (host) CUDA_TRY(cudaMemcpyAsync(dev, user, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  void finish() {
    if (host) {
      CUDA_TRY(cudaMemcpyAsync(user, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      // a result handed to the host must not come out of an exchange that gave up waiting for a peer
      if (c->nranks > 1 && c->peers_open && comm_peer_error(c))
        throw DeviceError{ALFIB_ECUDA, "peer-memory exchange timed out waiting for another rank"};
    }
  }
};

void upload_values(alfib_ctx* c, Level& L, DBuf<double>& dst, const double* vals, int block_col_major) {
  ALFIB_REQUIRE(L.nnzb > 0, "set the BSR pattern first");
  const size_t count = (size_t)L.nnzb * L.bs * L.bs;
  dst.alloc(count);
  const cudaMemcpyKind kind = is_device_ptr(vals) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CUDA_TRY(cudaMemcpyAsync(dst.p, vals, count * sizeof(double), kind, c->stream));
  if (block_col_major) launch_transpose_blocks(c, dst.p, L.nnzb, L.bs);
}

// Greedy colouring in iteration order — same definition as
// alfi_b200.patches.greedy_colouring, so host and library agree bit for bit.
void greedy_colour(PatchSet& ps, int ndofs) {
  std::vector<uint64_t> used(ndofs, 0);
  ps.h_colour.assign(ps.npatch, -1);
  for (int32_t p : ps.h_order) {
    if (ps.h_colour[p] >= 0) continue;
    uint64_t m = 0;
    for (int64_t k = ps.h_off[p]; k < ps.h_off[p + 1]; ++k) m |= used[ps.h_dofs[k]];
    int col = 0;
    while (col < 64 && ((m >> col) & 1)) ++col;
    if (col >= 64) throw DeviceError{ALFIB_EINVAL, "more than 64 colours needed"};
    if (ps.h_off[p + 1] == ps.h_off[p]) col = 0;
    ps.h_colour[p] = col;
    for (int64_t k = ps.h_off[p]; k < ps.h_off[p + 1]; ++k) used[ps.h_dofs[k]] |= (uint64_t(1) << col);
  }
  for (auto& col : ps.h_colour)
    if (col < 0) col = 0;               // patches not in the iteration set are never applied
}

}  // namespace

void profile_flush(alfib_ctx* c) {
  if (c->ev_used == 0) return;
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < c->ev_used; ++i) {
    float ms = 0;
    const EventRec& r = c->ev_pool[i];
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      c->ev_ms[r.level][r.id] += ms;
      c->ev_calls[r.level][r.id] += 1;
    }
  }
  c->ev_used = 0;
}

extern "C" {

int alfib_create(int device, alfib_ctx** out) {
  if (!out) return ALFIB_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ALFIB_ECUDA;
  alfib_ctx* c = new (std::nothrow) alfib_ctx();
  if (!c) return ALFIB_ENOMEM;
  c->device = device;
  int rc = guarded(c, [&] {
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  });
  if (rc != ALFIB_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return ALFIB_OK;
}

int alfib_destroy(alfib_ctx* c) {
  if (!c) return ALFIB_EINVAL;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto& L : c->levels) {
    if (!L) continue;
    for (auto& ps : L->ps) release_patchset(ps);
    for (auto* b : {&L->rowptr, &L->colidx, &L->bc, &L->cb, &L->p_rowptr, &L->p_colidx, &L->pt_rowptr, &L->pt_colidx})
      b->release();
    for (auto* b : {&L->vals, &L->dvals, &L->a0vals, &L->p_vals, &L->pt_vals, &L->b, &L->x, &L->r, &L->w, &L->t1,
                    &L->t2, &L->t3, &L->t4, &L->V, &L->Z, &L->tc})
      b->release();
    L->halo.release();
    L->thalo.release();
    delete L;
    L = nullptr;
  }
  for (auto* b : {&c->fwork, &c->partial, &c->scal, &c->stage_in, &c->stage_in2, &c->stage_out, &c->coarse_lu,
                  &c->coarse_work, &c->coarse_inv, &c->coarse_partial, &c->coarse_r, &c->coarse_dx})
    b->release();
  {
    Schur& S = c->schur;
    for (auto* b : {&S.b_rowptr, &S.b_colidx, &S.bt_rowptr, &S.bt_colidx, &S.mi_rowptr, &S.mi_colidx}) b->release();
    for (auto* b : {&S.b_vals, &S.bt_vals, &S.mi_vals, &S.y1, &S.tu, &S.tp, &S.tp2, &S.V, &S.Z, &S.w, &S.r, &S.xs, &S.hd})
      b->release();
  }
  c->d_peer_slot.release();
  c->d_epoch.release();
  c->d_comm_err.release();
  c->d_gate.release();
  c->finfo.release();
  c->coarse_piv.release();
  c->coarse_seppos.release();
  c->coarse_info.release();
  for (void* p : c->host_registered) cudaHostUnregister(p);
  c->host_registered.clear();
  cycle_graph_invalidate(c);
  comm_peer_close(c);
  comm_destroy(c);
  if (c->cusolver) cusolverDnDestroy(c->cusolver);
  for (auto& r : c->ev_pool) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return ALFIB_OK;
}

const char* alfib_last_error(const alfib_ctx* c) { return c ? c->err.c_str() : "null context"; }

int alfib_set_option(alfib_ctx* c, int key, int value) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    switch (key) {
      case ALFIB_OPT_DETERMINISTIC: c->deterministic = value != 0; break;
      case ALFIB_OPT_SYNC_ALWAYS: c->sync_always = value != 0; break;
      case ALFIB_OPT_ROBUST_RESTRICT: c->robust_restrict = value != 0; break;
      case ALFIB_OPT_TRANSFER_REFINE: c->transfer_refine = value != 0; break;
      case ALFIB_OPT_CUDA_GRAPH: c->use_graph = value != 0; break;
      default: throw DeviceError{ALFIB_EINVAL, "unknown option"};
    }
  });
}

int alfib_set_deterministic(alfib_ctx* c, int flag) { return alfib_set_option(c, ALFIB_OPT_DETERMINISTIC, flag); }

int alfib_synchronize(alfib_ctx* c) {
  return guarded(c, [&] {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (comm_peer_error(c)) throw DeviceError{ALFIB_ECUDA, "peer-memory exchange timed out waiting for another rank"};
  });
}

int64_t alfib_launch_count(const alfib_ctx* c) { return c ? c->launches : -1; }

void* alfib_stream(alfib_ctx* c) { return c ? (void*)c->stream : nullptr; }

int alfib_host_register(alfib_ctx* c, void* ptr, int64_t bytes) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(ptr && bytes > 0, "bad host buffer");
    if (std::find(c->host_registered.begin(), c->host_registered.end(), ptr) != c->host_registered.end()) return;
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {       // e.g. a torch pinned tensor: nothing to do, nothing to undo
      cudaGetLastError();
      return;
    }
    CUDA_TRY(e);
    c->host_registered.push_back(ptr);
  });
}

int alfib_host_unregister(alfib_ctx* c, void* ptr) {
  return guarded(c, [&] {
    auto it = std::find(c->host_registered.begin(), c->host_registered.end(), ptr);
    if (it == c->host_registered.end()) return;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->host_registered.erase(it);
    CUDA_TRY(cudaHostUnregister(ptr));
  });
}

int alfib_comm_unique_id(void* out128) {
  if (!out128) return ALFIB_EINVAL;
  try {
    comm_unique_id(out128);
    return ALFIB_OK;
  } catch (const DeviceError& e) {
    return e.code;
  }
}

int alfib_comm_peer_handle(alfib_ctx* c, void* out64) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(out64 != nullptr, "null handle buffer");
    cycle_graph_invalidate(c);
    comm_peer_handle(c, out64);
  });
}

int alfib_comm_peer_open(alfib_ctx* c, const void* handles) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(handles != nullptr, "null handles");
    cycle_graph_invalidate(c);
    comm_peer_open(c, handles);
  });
}

int alfib_comm_init(alfib_ctx* c, const void* nccl_unique_id, int rank, int nranks) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    comm_init(c, nccl_unique_id, rank, nranks);
    for (auto* L : c->levels)
      if (L) set_row_partition(c, *L);
  });
}

// ---- distributed level vectors ----------------------------------------------------------------
int alfib_level_set_halo(alfib_ctx* c, int level, int which, int32_t n_owned, int32_t n_local, int32_t npeers,
                         const int32_t* peers, const int64_t* send_off, const int32_t* send_idx,
                         const int64_t* recv_off, const int32_t* recv_idx, const int64_t* peer_send_off,
                         const int64_t* peer_recv_off) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(which == 0 || which == 1, "which must be 0 (level vectors) or 1 (transfer halo)");
    ALFIB_REQUIRE(n_owned >= 0 && n_local >= n_owned, "0 <= n_owned <= n_local required");
    ALFIB_REQUIRE(npeers >= 0 && npeers < ALFIB_MAX_RANKS, "bad peer count");
    ALFIB_REQUIRE(npeers == 0 || (peers && send_off && recv_off), "null exchange lists");
    ALFIB_REQUIRE(npeers == 0 || c->comm, "alfib_comm_init first");
    if (which == 0) {
      ALFIB_REQUIRE(n_local == L.n && n_owned % L.bs == 0, "level halo: n_local must be the level's size, ownership node-wise");
      ALFIB_REQUIRE(!L.has_transfer, "alfib_level_set_halo must precede alfib_transfer_set");
    } else {
      ALFIB_REQUIRE(level >= 1, "a transfer halo belongs to a level >= 1");
      ALFIB_REQUIRE(!L.has_transfer, "alfib_level_set_halo must precede alfib_transfer_set");
      ALFIB_REQUIRE(n_owned == get_level(c, level - 1).n_owned, "transfer halo: owned part must be the coarser level's");
    }
    Halo& H = which == 0 ? L.halo : L.thalo;
    H.release();
    H.n_owned = n_owned;
    H.n_local = n_local;
    H.peers.assign(peers, peers + npeers);
    H.send_off.assign(1, 0);
    H.recv_off.assign(1, 0);
    for (int p = 0; p < npeers; ++p) {
      ALFIB_REQUIRE(peers[p] >= 0 && peers[p] < c->nranks && peers[p] != c->rank, "bad peer rank");
      ALFIB_REQUIRE(p == 0 || peers[p] > peers[p - 1], "peers must be ascending");
      ALFIB_REQUIRE(send_off[p + 1] >= send_off[p] && recv_off[p + 1] >= recv_off[p] && send_off[0] == 0 && recv_off[0] == 0,
                    "bad exchange offsets");
      H.send_off.push_back(send_off[p + 1]);
      H.recv_off.push_back(recv_off[p + 1]);
    }
    const int64_t ns = H.send_off.back(), nr = H.recv_off.back();
    ALFIB_REQUIRE((ns == 0 || send_idx) && (nr == 0 || recv_idx), "null exchange indices");
    for (int64_t k = 0; k < ns; ++k) ALFIB_REQUIRE(send_idx[k] >= 0 && send_idx[k] < n_owned, "send entry is not an owned dof");
    {
      std::vector<char> seen((size_t)(n_local - n_owned), 0);
      for (int64_t k = 0; k < nr; ++k) {
        ALFIB_REQUIRE(recv_idx[k] >= n_owned && recv_idx[k] < n_local, "receive entry is not a ghost dof");
        ALFIB_REQUIRE(!seen[recv_idx[k] - n_owned], "a ghost dof is received twice");
        seen[recv_idx[k] - n_owned] = 1;
      }
      ALFIB_REQUIRE(c->nranks == 1 || nr == (int64_t)(n_local - n_owned), "every ghost dof needs an owner to receive from");
    }
    for (int p = 0; p < npeers; ++p) {
      std::vector<int32_t> part(send_idx + H.send_off[p], send_idx + H.send_off[p + 1]);
      std::sort(part.begin(), part.end());
      ALFIB_REQUIRE(std::adjacent_find(part.begin(), part.end()) == part.end(), "an owned dof is sent twice to one peer");
    }
    {
      // ghost -> owner sum as a gather per distinct owned dof; positions ascending = peers in ascending rank order
      std::vector<int32_t> pos((size_t)ns);
      std::iota(pos.begin(), pos.end(), 0);
      std::stable_sort(pos.begin(), pos.end(), [&](int32_t a, int32_t b) { return send_idx[a] < send_idx[b]; });
      std::vector<int32_t> red_ptr(1, 0), red_dof, red_src((size_t)ns);
      for (int64_t k = 0; k < ns; ++k) {
        if (k == 0 || send_idx[pos[k]] != send_idx[pos[k - 1]]) {
          if (k) red_ptr.push_back((int32_t)k);
          red_dof.push_back(send_idx[pos[k]]);
        }
        red_src[k] = pos[k];
      }
      if (ns) red_ptr.push_back((int32_t)ns);
      H.n_red = (int)red_dof.size();
      H.red_ptr.upload(red_ptr.data(), red_ptr.size(), c->stream);
      H.red_dof.upload(red_dof.data(), red_dof.size(), c->stream);
      H.red_src.upload(red_src.data(), red_src.size(), c->stream);
    }
    if (peer_send_off && peer_recv_off) {
      H.peer_send_off.assign(peer_send_off, peer_send_off + npeers);
      H.peer_recv_off.assign(peer_recv_off, peer_recv_off + npeers);
      for (int p = 0; p < npeers; ++p)
        ALFIB_REQUIRE(H.peer_send_off[p] >= 0 && H.peer_recv_off[p] >= 0, "negative peer offset");
      H.has_peer_off = true;
    }
    H.send_idx.upload(send_idx, ns, c->stream);
    H.recv_idx.upload(recv_idx, nr, c->stream);
    H.sbuf.alloc(std::max<int64_t>(ns, 1));
    H.rbuf.alloc(std::max<int64_t>(nr, 1));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    H.on = true;
    if (which == 0) L.n_owned = n_owned;
    comm_mbox_reserve(c, H, level, which);
  });
}

// ---- level operator ---------------------------------------------------------------------------
int alfib_level_create(alfib_ctx* c, int level, int n_nodes, int bs) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    ALFIB_REQUIRE(level >= 0 && level < ALFIB_MAX_LEVELS, "level out of range");
    ALFIB_REQUIRE(n_nodes > 0 && (bs == 2 || bs == 3), "n_nodes > 0 and bs in {2,3} required");
    ALFIB_REQUIRE(!c->levels[level], "level already exists");
    Level* L = new Level();
    L->n_nodes = n_nodes;
    L->bs = bs;
    L->n = n_nodes * bs;
    L->n_owned = L->n;
    L->index = level;
    set_row_partition(c, *L);
    c->levels[level] = L;
  });
}

int alfib_level_set_bsr_pattern(alfib_ctx* c, int level, int64_t nnzb, const int32_t* rowptr, const int32_t* colidx) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(rowptr && colidx && nnzb > 0, "bad pattern");
    ALFIB_REQUIRE(rowptr[0] == 0 && rowptr[L.n_nodes] == nnzb, "rowptr does not match nnzb");
    for (int64_t k = 0; k < nnzb; ++k) ALFIB_REQUIRE(colidx[k] >= 0 && colidx[k] < L.n_nodes, "column index out of range");
    L.nnzb = nnzb;
    L.h_rowptr.assign(rowptr, rowptr + L.n_nodes + 1);
    L.h_colidx.assign(colidx, colidx + nnzb);
    L.rowptr.upload(rowptr, L.n_nodes + 1, c->stream);
    L.colidx.upload(colidx, nnzb, c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  });
}

int alfib_level_set_bsr_values(alfib_ctx* c, int level, const double* vals, int block_col_major) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(vals, "null values");
    const double* before = L.vals.p;
    upload_values(c, L, L.vals, vals, block_col_major);
    if (L.vals.p != before) cycle_graph_invalidate(c);        // the captured cycle holds the old pointer
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    L.has_values = true;
    L.ps[ALFIB_PATCHES_SMOOTHER].factored = false;
    L.ps[ALFIB_PATCHES_SMOOTHER].corr_fresh = false;
    if (level == 0) c->coarse_factored = false;
  });
}

int alfib_level_set_bc(alfib_ctx* c, int level, int32_t nbc, const int32_t* bc_dofs) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(nbc >= 0 && (nbc == 0 || bc_dofs), "bad bc list");
    for (int i = 0; i < nbc; ++i) ALFIB_REQUIRE(bc_dofs[i] >= 0 && bc_dofs[i] < L.n, "bc dof out of range");
    L.nbc = nbc;
    L.bc.upload(bc_dofs, nbc, c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  });
}

int alfib_spmv(alfib_ctx* c, int level, const double* x, double* y) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(L.has_values, "level has no values");
    const double* dx = in_vec(c, x, L.n, c->stage_in);
    OutVec out(c, y, L.n);
    ScopedEvent ev(c, ALFIB_EV_MATMULT, level);
    launch_bsr_spmv(c, L, L.vals.p, dx, out.dev, nullptr);
    out.finish();
  });
}

int alfib_residual(alfib_ctx* c, int level, const double* b, const double* x, double* r) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(L.has_values, "level has no values");
    const double* db = in_vec(c, b, L.n, c->stage_in);
    const double* dx = in_vec(c, x, L.n, c->stage_in2);
    OutVec out(c, r, L.n);
    ScopedEvent ev(c, ALFIB_EV_MATMULT, level);
    launch_bsr_spmv(c, L, L.vals.p, dx, out.dev, db);
    out.finish();
  });
}

// ---- patches ----------------------------------------------------------------------------------
int alfib_level_set_patches(alfib_ctx* c, int level, int which, int32_t npatch, const int64_t* offsets,
                            const int32_t* dofs, int32_t norder, const int32_t* order, const int32_t* colours) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(which == 0 || which == 1, "which must be 0 or 1");
    ALFIB_REQUIRE(npatch >= 0 && offsets && offsets[0] == 0, "bad offsets");
    PatchSet& ps = L.ps[which];
    release_patchset(ps);
    ps = PatchSet();
    ps.npatch = npatch;
    ps.h_off.assign(offsets, offsets + npatch + 1);
    const int64_t total = offsets[npatch];
    ALFIB_REQUIRE(total == 0 || dofs, "null dof list");
    ps.h_dofs.assign(dofs, dofs + total);
    for (int p = 0; p < npatch; ++p) {
      ALFIB_REQUIRE(offsets[p + 1] >= offsets[p], "offsets must be non-decreasing");
      ps.maxn = std::max<int>(ps.maxn, (int)(offsets[p + 1] - offsets[p]));
    }
    for (int64_t k = 0; k < total; ++k) ALFIB_REQUIRE(dofs[k] >= 0 && dofs[k] < L.n, "patch dof out of range");
    if (order) {
      ps.h_order.assign(order, order + norder);
      for (int32_t p : ps.h_order) ALFIB_REQUIRE(p >= 0 && p < npatch, "iteration set entry out of range");
    } else {
      ps.h_order.resize(npatch);
      std::iota(ps.h_order.begin(), ps.h_order.end(), 0);
    }
    {
      std::vector<char> seen(npatch, 0);
      for (int32_t p : ps.h_order) {
        if (seen[p]) ps.repeated = true;
        seen[p] = 1;
      }
    }
    if (colours) {
      ps.h_colour.assign(colours, colours + npatch);
      for (int32_t col : ps.h_colour) ALFIB_REQUIRE(col >= 0 && col < 64, "colour out of range");
      // two patches of one colour must not share a dof: the deterministic scatter is a plain read-modify-write per colour
      std::vector<uint64_t> used(L.n, 0);
      std::vector<char> done(npatch, 0);
      for (int32_t p : ps.h_order) {
        if (done[p]) continue;
        done[p] = 1;
        const uint64_t bit = uint64_t(1) << ps.h_colour[p];
        for (int64_t k = offsets[p]; k < offsets[p + 1]; ++k) {
          ALFIB_REQUIRE(!(used[dofs[k]] & bit), "two patches of the same colour share a dof");
          used[dofs[k]] |= bit;
        }
      }
    } else {
      greedy_colour(ps, L.n);
    }
    ps.ncolour = npatch ? 1 + *std::max_element(ps.h_colour.begin(), ps.h_colour.end()) : 0;

    // sorted dof lists for the gather, factor storage offsets
    std::vector<int32_t> sorted(total), sperm(total);
    ps.h_soff.assign(npatch + 1, 0);
    for (int p = 0; p < npatch; ++p) {
      const int64_t o = offsets[p];
      const int n = (int)(offsets[p + 1] - o);
      std::vector<int32_t> idx(n);
      std::iota(idx.begin(), idx.end(), 0);
      std::sort(idx.begin(), idx.end(), [&](int a, int b) { return dofs[o + a] < dofs[o + b]; });
      for (int i = 0; i < n; ++i) {
        sorted[o + i] = dofs[o + idx[i]];
        sperm[o + i] = idx[i];
        if (i) ALFIB_REQUIRE(sorted[o + i] != sorted[o + i - 1], "duplicate dof inside a patch");
      }
      ps.h_soff[p + 1] = ps.h_soff[p] + (int64_t)n * roundup2(n);
    }
    ps.store_elems = ps.h_soff[npatch];
    if (total > 0) {
      ps.lo = *std::min_element(ps.h_dofs.begin(), ps.h_dofs.end());
      ps.hi = 1 + (long long)*std::max_element(ps.h_dofs.begin(), ps.h_dofs.end());
    }
    // factor order: largest patches first
    std::vector<int32_t> forder(npatch);
    std::iota(forder.begin(), forder.end(), 0);
    std::stable_sort(forder.begin(), forder.end(), [&](int a, int b) {
      return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b];
    });
    // apply work list: colour-major, iteration order inside a colour, tiles ascending
    std::vector<int2> work;
    ps.colour_work_start.assign(ps.ncolour + 1, 0);
    for (int col = 0; col < ps.ncolour; ++col) {
      ps.colour_work_start[col] = (int)work.size();
      for (int32_t p : ps.h_order) {
        if (ps.h_colour[p] != col) continue;
        const int n = (int)(offsets[p + 1] - offsets[p]);
        for (int t = 0; t * ALFIB_TILE_ROWS < n; ++t) work.push_back(make_int2(p, t));
      }
    }
    if (ps.ncolour) ps.colour_work_start[ps.ncolour] = (int)work.size();
    ps.nwork = (int)work.size();

    ps.off.upload(ps.h_off.data(), npatch + 1, c->stream);
    ps.soff.upload(ps.h_soff.data(), npatch + 1, c->stream);
    ps.dofs.upload(ps.h_dofs.data(), total, c->stream);
    ps.sorted.upload(sorted.data(), total, c->stream);
    ps.sperm.upload(sperm.data(), total, c->stream);
    ps.forder.upload(forder.data(), npatch, c->stream);
    ps.work.upload(work.data(), work.size(), c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    comm_peer_publish_ranges(c);
  });
}

int alfib_level_set_patch_blocks(alfib_ctx* c, int level, int which, const int32_t* block_of_dof) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(which == 0 || which == 1, "which must be 0 or 1");
    PatchSet& ps = L.ps[which];
    ALFIB_REQUIRE(ps.npatch > 0 || ps.h_off.size() == 1, "alfib_level_set_patches first");
    if (!block_of_dof) {                       // back to dense inverses
      if (ps.cond.on) {
        ps.cond.release();
        ps.cond = Condensed();
        ps.store_elems = ps.h_soff.empty() ? 0 : ps.h_soff.back();
        ps.store = nullptr;
        ps.store_buf.release();
        ps.store_owned = false;
        ps.factored = false;
      }
      return;
    }
    condense_setup(c, L, ps, block_of_dof);
  });
}

int alfib_level_set_sweep_stages(alfib_ctx* c, int level, int which, int32_t nvisit, const int32_t* stage_of_visit,
                                 int32_t nstage, int symmetric) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(which == 0 || which == 1, "which must be 0 or 1");
    PatchSet& ps = L.ps[which];
    ps.nstage = 0;
    ps.symmetric_sweep = false;
    ps.stage_work.release();
    ps.stage_rows.release();
    if (nstage == 0) return;                                   // additive composition
    ALFIB_REQUIRE(!ps.cond.on, "multiplicative composition needs dense patch inverses (no patch blocks)");
    ALFIB_REQUIRE(c->nranks == 1 && !L.halo.on, "multiplicative composition is single-GPU");
    ALFIB_REQUIRE(nvisit == (int32_t)ps.h_order.size() && stage_of_visit, "one stage per entry of the iteration set");
    ALFIB_REQUIRE(!L.h_rowptr.empty(), "set the BSR pattern before the sweep stages");
    const int bs = L.bs;
    // check: within a stage no two visits are coupled (node adjacency of the pattern, which includes shared nodes),
    // and a coupled pair of visits keeps its iteration order across stages
    std::vector<int> last_stage_of_node(L.n_nodes, -1), last_visit_of_node(L.n_nodes, -1);   // writer of each node so far
    std::vector<std::vector<int>> visits_of_stage(nstage);
    for (int k = 0; k < nvisit; ++k) {
      ALFIB_REQUIRE(stage_of_visit[k] >= 0 && stage_of_visit[k] < nstage, "stage out of range");
      visits_of_stage[stage_of_visit[k]].push_back(k);
    }
    for (int k = 0; k < nvisit; ++k) {                         // iteration order
      const int p = ps.h_order[k], s = stage_of_visit[k];
      const int64_t o = ps.h_off[p], e = ps.h_off[p + 1];
      for (int64_t i = o; i < e; i += bs) {                    // node-wise patches: bs consecutive dofs per node
        const int node = ps.h_dofs[i] / bs;
        for (int q = L.h_rowptr[node]; q < L.h_rowptr[node + 1]; ++q) {
          const int nb = L.h_colidx[q];                        // this visit reads y at nb
          if (last_visit_of_node[nb] >= 0 && last_visit_of_node[nb] != k)
            ALFIB_REQUIRE(last_stage_of_node[nb] < s, "sweep stages: coupled visits must lie in increasing stages");
        }
      }
      for (int64_t i = o; i < e; ++i) {
        const int node = ps.h_dofs[i] / bs;
        last_stage_of_node[node] = s;
        last_visit_of_node[node] = k;
      }
    }
    std::vector<int2> work;
    std::vector<int32_t> rows;
    ps.stage_work_start.assign(nstage + 1, 0);
    ps.stage_row_start.assign(nstage + 1, 0);
    std::vector<int> mark(L.n_nodes, -1);
    for (int s = 0; s < nstage; ++s) {
      ps.stage_work_start[s] = (int)work.size();
      ps.stage_row_start[s] = (int)rows.size();
      for (int k : visits_of_stage[s]) {
        const int p = ps.h_order[k];
        const int n = (int)(ps.h_off[p + 1] - ps.h_off[p]);
        for (int t = 0; t * ALFIB_TILE_ROWS < n; ++t) work.push_back(make_int2(p, t));
        for (int64_t i = ps.h_off[p]; i < ps.h_off[p + 1]; ++i) {
          const int node = ps.h_dofs[i] / bs;
          if (mark[node] != s) {
            mark[node] = s;
            rows.push_back(node);
          }
        }
      }
    }
    ps.stage_work_start[nstage] = (int)work.size();
    ps.stage_row_start[nstage] = (int)rows.size();
    ps.stage_work.upload(work.data(), work.size(), c->stream);
    ps.stage_rows.upload(rows.data(), rows.size(), c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    ps.nstage = nstage;
    ps.symmetric_sweep = symmetric != 0;
  });
}

int alfib_level_set_patch_corrections(alfib_ctx* c, int level, int which, const int64_t* corr_off, const int32_t* rows,
                                      const int32_t* cols) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(which == 0 || which == 1, "which must be 0 or 1");
    PatchSet& ps = L.ps[which];
    ps.factored = false;
    ps.has_corr = ps.corr_fresh = false;
    if (!corr_off) {
      ps.corr_off.release(); ps.corr_rows.release(); ps.corr_cols.release(); ps.corr_vals.release();
      ps.corr_nnz = 0;
      return;
    }
    ALFIB_REQUIRE(ps.npatch > 0, "alfib_level_set_patches first");
    ALFIB_REQUIRE(!ps.cond.on, "patch corrections need dense patch inverses (no patch blocks)");
    ALFIB_REQUIRE(corr_off[0] == 0, "bad correction offsets");
    const int64_t nnz = corr_off[ps.npatch];
    ALFIB_REQUIRE(nnz == 0 || (rows && cols), "null correction pattern");
    for (int p = 0; p < ps.npatch; ++p) {
      ALFIB_REQUIRE(corr_off[p + 1] >= corr_off[p], "correction offsets must be non-decreasing");
      const int n = (int)(ps.h_off[p + 1] - ps.h_off[p]);
      for (int64_t e = corr_off[p]; e < corr_off[p + 1]; ++e) {
        ALFIB_REQUIRE(rows[e] >= 0 && rows[e] < n && cols[e] >= 0 && cols[e] < n, "correction entry outside its patch");
        // distinct entries: the kernel adds them without atomics
        ALFIB_REQUIRE(e == corr_off[p] || rows[e] > rows[e - 1] || (rows[e] == rows[e - 1] && cols[e] > cols[e - 1]),
                      "correction entries must be sorted by (row, col) and distinct");
      }
    }
    ps.corr_off.upload(corr_off, ps.npatch + 1, c->stream);
    ps.corr_rows.upload(rows, nnz, c->stream);
    ps.corr_cols.upload(cols, nnz, c->stream);
    ps.corr_vals.alloc(nnz);
    ps.corr_nnz = nnz;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    ps.has_corr = true;
  });
}

int alfib_level_set_patch_correction_values(alfib_ctx* c, int level, int which, const double* vals) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(which == 0 || which == 1, "which must be 0 or 1");
    PatchSet& ps = L.ps[which];
    ALFIB_REQUIRE(ps.has_corr, "alfib_level_set_patch_corrections first");
    ALFIB_REQUIRE(vals || ps.corr_nnz == 0, "null correction values");
    if (ps.corr_nnz) {
      const cudaMemcpyKind kind = is_device_ptr(vals) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
      CUDA_TRY(cudaMemcpyAsync(ps.corr_vals.p, vals, sizeof(double) * (size_t)ps.corr_nnz, kind, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    ps.factored = false;
    ps.corr_fresh = true;
  });
}

int64_t alfib_patch_apply_bytes(alfib_ctx* c, int level, int which) {
  if (!c || level < 0 || level >= ALFIB_MAX_LEVELS || !c->levels[level] || which < 0 || which > 1) return -1;
  const Level& L = *c->levels[level];
  const PatchSet& ps = L.ps[which];
  const int64_t index = ps.cond.on ? ps.cond.h.index_bytes : (int64_t)sizeof(int32_t) * (int64_t)ps.h_dofs.size();
  return ps.store_elems * (int64_t)sizeof(double) + index + 16 * (int64_t)L.n;
}

int64_t alfib_patch_storage_bytes(alfib_ctx* c, int level, int which) {
  if (!c || level < 0 || level >= ALFIB_MAX_LEVELS || !c->levels[level] || which < 0 || which > 1) return -1;
  return c->levels[level]->ps[which].store_elems * (int64_t)sizeof(double);
}

int alfib_patch_storage_form(alfib_ctx* c, int level, int which) {
  if (!c || level < 0 || level >= ALFIB_MAX_LEVELS || !c->levels[level] || which < 0 || which > 1) return -1;
  const PatchSet& ps = c->levels[level]->ps[which];
  return !ps.cond.on ? 0 : (ps.cond.h.shared ? 2 : 1);
}

int alfib_patch_bind_storage(alfib_ctx* c, int level, int which, void* dev_ptr, int64_t bytes) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(which == 0 || which == 1, "which must be 0 or 1");
    PatchSet& ps = L.ps[which];
    ALFIB_REQUIRE(dev_ptr && is_device_ptr(dev_ptr), "storage must be a device pointer");
    ALFIB_REQUIRE(bytes >= ps.store_elems * (int64_t)sizeof(double), "storage too small");
    ALFIB_REQUIRE(((uintptr_t)dev_ptr & 15) == 0, "storage must be 16-byte aligned");
    ps.store_buf.release();
    ps.store = static_cast<double*>(dev_ptr);
    ps.store_owned = false;
    ps.factored = false;
  });
}

static void ensure_storage(PatchSet& ps) {
  if (!ps.store) {
    ps.store_buf.alloc((size_t)std::max<int64_t>(ps.store_elems, 2));
    ps.store = ps.store_buf.p;
    ps.store_owned = true;
  }
}

int alfib_level_factor(alfib_ctx* c, int level) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(L.has_values, "level has no values");
    PatchSet& ps = L.ps[ALFIB_PATCHES_SMOOTHER];
    if (ps.npatch == 0) {                 // a rank may own no patch of a small level
      ps.factored = true;
      return;
    }
    ensure_storage(ps);
    ScopedEvent ev(c, ALFIB_EV_PCSETUP_PATCH, level);
    launch_patch_factor(c, L, ps, L.vals.p);
  });
}

int alfib_smoother_apply(alfib_ctx* c, int level, const double* x, double* y) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    const double* dx = in_vec(c, x, L.n, c->stage_in);
    OutVec out(c, y, L.n);
    ALFIB_REQUIRE(dx != out.dev, "x and y must not alias");
    smoother_apply_device(c, L, level, dx, out.dev);
    out.finish();
  });
}

int alfib_get_colours(alfib_ctx* c, int level, int which, int32_t* colours) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE((which == 0 || which == 1) && colours, "bad arguments");
    const PatchSet& ps = L.ps[which];
    std::copy(ps.h_colour.begin(), ps.h_colour.end(), colours);
  });
}

int alfib_get_patch_inverse(alfib_ctx* c, int level, int which, int32_t patch, double* out) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE((which == 0 || which == 1) && out, "bad arguments");
    patch_extract_inverse(c, L.ps[which], patch, out);
  });
}

// ---- transfer ---------------------------------------------------------------------------------
int alfib_transfer_set(alfib_ctx* c, int level, int32_t n_fine_nodes, int32_t n_coarse_nodes,
                       const int32_t* P_rowptr, const int32_t* P_colidx, const double* P_vals, int32_t ncb,
                       const int32_t* cb_dofs, int dof_level) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(level >= 1, "transfers live on levels >= 1");
    Level& Lc = get_level(c, level - 1);
    if (L.halo.on) {
      // distributed vectors: P holds the owned fine rows; columns = the coarser level in the transfer-halo layout,
      // or the whole (replicated) coarser level
      ALFIB_REQUIRE(dof_level, "a level with a halo takes a dof-level P");
      ALFIB_REQUIRE(L.thalo.on || !Lc.halo.on, "the coarser level is distributed: set the transfer halo (which = 1) first");
      ALFIB_REQUIRE(n_fine_nodes == L.n_owned && n_coarse_nodes == (L.thalo.on ? L.thalo.n_local : Lc.n),
                    "local P shape does not match the halo layouts");
    } else if (dof_level)
      ALFIB_REQUIRE(n_fine_nodes == L.n && n_coarse_nodes == Lc.n, "dof-level P shape does not match the levels");
    else
      ALFIB_REQUIRE(n_fine_nodes == L.n_nodes && n_coarse_nodes == Lc.n_nodes, "P shape does not match the levels");
    L.p_bs = dof_level ? 1 : L.bs;
    ALFIB_REQUIRE(P_rowptr && P_colidx && P_vals && P_rowptr[0] == 0, "bad prolongation matrix");
    const int64_t nnz = P_rowptr[n_fine_nodes];
    for (int64_t k = 0; k < nnz; ++k) ALFIB_REQUIRE(P_colidx[k] >= 0 && P_colidx[k] < n_coarse_nodes, "P column out of range");
    L.p_rows = n_fine_nodes;
    L.p_cols = n_coarse_nodes;
    L.p_rowptr.upload(P_rowptr, n_fine_nodes + 1, c->stream);
    L.p_colidx.upload(P_colidx, nnz, c->stream);
    L.p_vals.upload(P_vals, nnz, c->stream);
    // explicit transpose (CSR of P^T), rows in ascending fine-node order => fixed summation order
    std::vector<int32_t> trow(n_coarse_nodes + 1, 0), tcol(nnz);
    std::vector<double> tval(nnz);
    for (int64_t k = 0; k < nnz; ++k) trow[P_colidx[k] + 1]++;
    for (int i = 0; i < n_coarse_nodes; ++i) trow[i + 1] += trow[i];
    std::vector<int32_t> cursor(trow.begin(), trow.end() - 1);
    for (int i = 0; i < n_fine_nodes; ++i)
      for (int k = P_rowptr[i]; k < P_rowptr[i + 1]; ++k) {
        const int dst = cursor[P_colidx[k]]++;
        tcol[dst] = i;
        tval[dst] = P_vals[k];
      }
    L.pt_rowptr.upload(trow.data(), n_coarse_nodes + 1, c->stream);
    L.pt_colidx.upload(tcol.data(), nnz, c->stream);
    L.pt_vals.upload(tval.data(), nnz, c->stream);
    ALFIB_REQUIRE(ncb >= 0 && (ncb == 0 || cb_dofs), "bad coarse-boundary list");
    for (int i = 0; i < ncb; ++i) ALFIB_REQUIRE(cb_dofs[i] >= 0 && cb_dofs[i] < L.n, "coarse-boundary dof out of range");
    L.ncb = ncb;
    L.cb.upload(cb_dofs, ncb, c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    L.has_transfer = true;
  });
}

int alfib_transfer_update(alfib_ctx* c, int level, const double* A0_vals, const double* D_vals, int block_col_major) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(L.has_transfer, "alfib_transfer_set first");
    ALFIB_REQUIRE(A0_vals || D_vals, "nothing to update");
    // the captured cycle follows the transfer sequence has_d selects and holds the value pointers
    const double *d_before = L.dvals.p, *a0_before = L.a0vals.p;
    const bool had_d = L.has_d;
    if (D_vals) {
      upload_values(c, L, L.dvals, D_vals, block_col_major);
      L.has_d = true;
    }
    if (L.has_d != had_d || L.dvals.p != d_before) cycle_graph_invalidate(c);
    if (A0_vals) {
      PatchSet& ps = L.ps[ALFIB_PATCHES_TRANSFER];
      upload_values(c, L, L.a0vals, A0_vals, block_col_major);
      if (L.a0vals.p != a0_before) cycle_graph_invalidate(c);
      if (ps.npatch > 0) {
        ensure_storage(ps);
        ScopedEvent ev(c, ALFIB_EV_PCSETUP_PATCH, level);
        launch_patch_factor(c, L, ps, L.a0vals.p);
      } else {
        ps.factored = true;
      }
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  });
}

int alfib_prolong(alfib_ctx* c, int level, const double* coarse, double* fine) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    Level& Lc = get_level(c, level - 1);
    const double* dc = in_vec(c, coarse, Lc.n, c->stage_in);
    OutVec out(c, fine, L.n);
    prolong_device(c, L, level, dc, out.dev);
    out.finish();
  });
}

int alfib_restrict(alfib_ctx* c, int level, const double* fine, double* coarse) {
  return guarded(c, [&] {
    Level& L = get_level(c, level);
    Level& Lc = get_level(c, level - 1);
    const double* df = in_vec(c, fine, L.n, c->stage_in);
    OutVec out(c, coarse, Lc.n);
    restrict_device(c, L, Lc, level, df, out.dev);
    out.finish();
  });
}

// ---- smoother / cycle -------------------------------------------------------------------------
int alfib_smooth(alfib_ctx* c, int level, int m, const double* b, double* x) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);             // may reallocate the Krylov bases
    Level& L = get_level(c, level);
    ALFIB_REQUIRE(L.has_values, "level has no values");
    const double* db = in_vec(c, b, L.n, c->stage_in);
    OutVec out(c, x, L.n);
    out.load();
    fgmres_device(c, L, level, m, db, out.dev);
    out.finish();
  });
}

int alfib_coarse_factor(alfib_ctx* c) {
  return guarded(c, [&] { coarse_factor_device(c); });
}

int alfib_coarse_solve(alfib_ctx* c, const double* b, double* x) {
  return guarded(c, [&] {
    Level& L = get_level(c, 0);
    const double* db = in_vec(c, b, L.n, c->stage_in);
    OutVec out(c, x, L.n);
    coarse_solve_device(c, db, out.dev);
    out.finish();
  });
}

int alfib_cycle_setup(alfib_ctx* c, int nlevels, int smoothing) {
  return guarded(c, [&] {
    cycle_graph_invalidate(c);
    ALFIB_REQUIRE(nlevels >= 1 && nlevels <= ALFIB_MAX_LEVELS, "bad level count");
    ALFIB_REQUIRE(smoothing >= 1 && smoothing <= ALFIB_MAX_KRYLOV, "bad smoothing count");
    for (int l = 0; l < nlevels; ++l) {
      Level& L = get_level(c, l);
      for (auto* v : {&L.b, &L.x, &L.w, &L.r, &L.t1, &L.t2}) {
        v->alloc(L.n);
        if (L.halo.on) CUDA_TRY(cudaMemsetAsync(v->p, 0, sizeof(double) * L.n, c->stream));   // ghost parts defined
      }
      if (l > 0) ALFIB_REQUIRE(L.has_transfer, "level without transfer");
    }
    c->nlevels = nlevels;
    c->smoothing = smoothing;
  });
}

int alfib_cycle_apply(alfib_ctx* c, const double* b, double* x) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(c->nlevels >= 1, "alfib_cycle_setup first");
    Level& L = get_level(c, c->nlevels - 1);
    const double* db = in_vec(c, b, L.n, c->stage_in);
    OutVec out(c, x, L.n);
    cycle_apply_device(c, db, out.dev);
    out.finish();
  });
}

// ---- outer Schur-complement fieldsplit (outer.cu) -------------------------------------------------
int alfib_schur_set(alfib_ctx* c, int32_t n_p, const int32_t* B_rowptr, const int32_t* B_colidx, const double* B_vals,
                    const int32_t* Mi_rowptr, const int32_t* Mi_colidx, const double* Mi_vals, int remove_constant) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(c->nlevels >= 1, "alfib_cycle_setup first");
    Level& L = get_level(c, c->nlevels - 1);
    ALFIB_REQUIRE(!L.halo.on && c->nranks == 1, "the outer pieces run on one GPU (finest level not distributed)");
    ALFIB_REQUIRE(n_p >= 1 && B_rowptr && B_colidx && B_vals && B_rowptr[0] == 0, "bad divergence matrix");
    ALFIB_REQUIRE(Mi_rowptr && Mi_colidx && Mi_vals && Mi_rowptr[0] == 0, "bad pressure mass inverse");
    const int nu = L.n;
    const int64_t nnz = B_rowptr[n_p], mnnz = Mi_rowptr[n_p];
    for (int i = 0; i < n_p; ++i)
      ALFIB_REQUIRE(B_rowptr[i + 1] >= B_rowptr[i] && Mi_rowptr[i + 1] >= Mi_rowptr[i], "row pointers must not decrease");
    for (int64_t k = 0; k < nnz; ++k) ALFIB_REQUIRE(B_colidx[k] >= 0 && B_colidx[k] < nu, "B column out of range");
    for (int64_t k = 0; k < mnnz; ++k) ALFIB_REQUIRE(Mi_colidx[k] >= 0 && Mi_colidx[k] < n_p, "M_p^-1 column out of range");
    Schur& S = c->schur;
    S.on = false;
    S.nu = nu;
    S.np = n_p;
    S.remove_mean = remove_constant != 0;
    S.b_nnz = nnz;
    S.mi_nnz = mnnz;
    S.b_rowptr.upload(B_rowptr, n_p + 1, c->stream);
    S.b_colidx.upload(B_colidx, nnz, c->stream);
    S.b_vals.upload(B_vals, nnz, c->stream);
    S.mi_rowptr.upload(Mi_rowptr, n_p + 1, c->stream);
    S.mi_colidx.upload(Mi_colidx, mnnz, c->stream);
    S.mi_vals.upload(Mi_vals, mnnz, c->stream);
    // explicit transpose, rows in ascending pressure-dof order => fixed summation order
    std::vector<int32_t> trow(nu + 1, 0), tcol(nnz);
    std::vector<double> tval(nnz);
    for (int64_t k = 0; k < nnz; ++k) trow[B_colidx[k] + 1]++;
    for (int i = 0; i < nu; ++i) trow[i + 1] += trow[i];
    std::vector<int32_t> cursor(trow.begin(), trow.end() - 1);
    for (int i = 0; i < n_p; ++i)
      for (int k = B_rowptr[i]; k < B_rowptr[i + 1]; ++k) {
        const int dst = cursor[B_colidx[k]]++;
        tcol[dst] = i;
        tval[dst] = B_vals[k];
      }
    S.bt_rowptr.upload(trow.data(), nu + 1, c->stream);
    S.bt_colidx.upload(tcol.data(), nnz, c->stream);
    S.bt_vals.upload(tval.data(), nnz, c->stream);
    S.y1.alloc(nu);
    S.tu.alloc(nu);
    S.tp.alloc(n_p);
    S.tp2.alloc(n_p);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    S.on = true;
  });
}

int alfib_schur_apply(alfib_ctx* c, double nu, double gamma, const double* r, double* y) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(c->schur.on, "alfib_schur_set first");
    const size_t n = (size_t)c->schur.nu + c->schur.np;
    const double* dr = in_vec(c, r, n, c->stage_in2);
    OutVec out(c, y, n);
    schur_apply_device(c, nu, gamma, dr, out.dev);
    out.finish();
  });
}

int alfib_jacobian_apply(alfib_ctx* c, const double* z, double* Jz) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(c->schur.on, "alfib_schur_set first");
    const size_t n = (size_t)c->schur.nu + c->schur.np;
    const double* dz = in_vec(c, z, n, c->stage_in2);
    OutVec out(c, Jz, n);
    jacobian_apply_device(c, dz, out.dev);
    out.finish();
  });
}

int alfib_outer_solve(alfib_ctx* c, double nu, double gamma, const double* rhs, double* x, double rtol, double atol,
                      int32_t maxit, int32_t restart, int32_t* iterations, double* history, int32_t nhistory) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(c->schur.on, "alfib_schur_set first");
    ALFIB_REQUIRE(nhistory >= 0 && (nhistory == 0 || history), "bad history buffer");
    const size_t n = (size_t)c->schur.nu + c->schur.np;
    const double* db = in_vec(c, rhs, n, c->stage_in2);
    OutVec out(c, x, n);
    int its = 0;
    outer_solve_device(c, nu, gamma, db, out.dev, rtol, atol, maxit, restart, &its, history, nhistory);
    if (iterations) *iterations = its;
    out.finish();
  });
}

// ---- instrumentation --------------------------------------------------------------------------
int alfib_profile(alfib_ctx* c, int enable) {
  return guarded(c, [&] {
    if (!enable) profile_flush(c);
    c->profile = enable != 0;
  });
}

int alfib_profile_get(alfib_ctx* c, int level, double* ms, int64_t* calls) {
  return guarded(c, [&] {
    ALFIB_REQUIRE(level >= -1 && level < ALFIB_MAX_LEVELS, "level out of range");
    profile_flush(c);
    for (int i = 0; i < ALFIB_EV_COUNT; ++i) {
      double t = 0;
      int64_t n = 0;
      for (int l = 0; l < ALFIB_MAX_LEVELS; ++l)
        if (level < 0 || l == level) {
          t += c->ev_ms[l][i];
          n += c->ev_calls[l][i];
        }
      if (ms) ms[i] = t;
      if (calls) calls[i] = n;
    }
  });
}

int alfib_profile_reset(alfib_ctx* c) {
  return guarded(c, [&] {
    profile_flush(c);
    for (int l = 0; l < ALFIB_MAX_LEVELS; ++l)
      for (int i = 0; i < ALFIB_EV_COUNT; ++i) {
        c->ev_ms[l][i] = 0;
        c->ev_calls[l][i] = 0;
      }
  });
}

}  // extern "C"
