// Internal declarations of libalfib (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "alfib.h"
#include "condense_host.h"

#define ALFIB_MAX_LEVELS 16
#define ALFIB_TILE_ROWS 64          // rows per apply tile (one warp, double2 per lane)
#define ALFIB_MAX_KRYLOV 32         // upper bound on FGMRES(m) per level
#define ALFIB_MAX_RANKS 8           // one NVSwitch box

// Destination of a kernel whose result takes part in a peer-memory exchange: either a plain
// pointer (epoch == nullptr) or one of the two symmetric slots, chosen on the device from the
// exchange counter so that the choice survives CUDA-graph replay.
struct PeerOut {
  double* base;
  size_t stride;
  const unsigned long long* epoch;
};
#ifdef __CUDACC__
__device__ __forceinline__ double* resolve(const PeerOut& o) {
  return o.epoch ? o.base + ((*o.epoch) & 1ull) * o.stride : o.base;
}
#endif
inline PeerOut plain_out(double* p) { return PeerOut{p, 0, nullptr}; }

// Header at the start of every rank's symmetric buffer, read by the peers over NVLink.
struct SymHeader {
  unsigned long long flag;                 // number of the last exchange this rank's data is ready for
  long long lo[2 * ALFIB_MAX_LEVELS];      // support [lo, hi) of this rank's patches, per (level, which)
  long long hi[2 * ALFIB_MAX_LEVELS];
};
#define ALFIB_SYM_HEADER_BYTES 4096

// Mailbox ("push") transport of the distributed-vector exchanges (comm.cu): behind the two slots every rank's
// symmetric buffer holds an arena of channels, one per (level, halo, direction) plus one for the small
// all-reduces of the dots.  A channel is [2 slots x cap doubles | one flag per sender rank]; senders write the
// receiver's slot with NVLink peer stores and then raise their flag there, the receiver polls LOCAL memory.
// Where a rank keeps each channel is published in a table inside its header (read by the peers at
// alfib_comm_peer_open), so the arenas need not be laid out identically on all ranks.
#define ALFIB_MBOX_TABLE_OFF 2048
#define ALFIB_MBOX_NV 64                                   // values per rank in a small all-reduce
#define ALFIB_MBOX_CHANNELS (ALFIB_MAX_LEVELS * 4 + 1)
#define ALFIB_MBOX_SMALL (ALFIB_MAX_LEVELS * 4)
struct MboxEntry {
  long long data_off;     // bytes from the start of the symmetric buffer to slot 0 of the channel (0 = none)
  long long cap;          // doubles per slot
  long long flags_off;    // bytes to the ALFIB_MAX_RANKS flags (unsigned long long) of the channel
};
static_assert(ALFIB_MBOX_TABLE_OFF + ALFIB_MBOX_CHANNELS * sizeof(MboxEntry) <= ALFIB_SYM_HEADER_BYTES, "channel table must fit the header");

struct DeviceError {
  int code;
  std::string msg;
};

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw DeviceError{ALFIB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)};      \
  } while (0)

#define ALFIB_REQUIRE(cond, what)                                                              \
  do {                                                                                         \
    if (!(cond)) throw DeviceError{ALFIB_EINVAL, std::string(what)};                           \
  } while (0)

template <class T>
struct DBuf {                       // device buffer owned by the library
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    if (count == n && p) return;
    release();
    if (count) {
      cudaError_t e = cudaMalloc(&p, count * sizeof(T));
      if (e != cudaSuccess)
        throw DeviceError{ALFIB_ENOMEM, "cudaMalloc of " + std::to_string(count * sizeof(T)) + " bytes: " +
                                            cudaGetErrorString(e)};
    }
    n = count;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void upload(const T* host, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) CUDA_TRY(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};

// ---- condensed ("statically condensed") patch sets: condense.cu / condense_host.h ----------------
struct Condensed {
  bool on = false;
  CondensedHost h;                      // layout, index lists and op lists (host copies)
  DBuf<int64_t> sepoff, ssoff;
  DBuf<int32_t> seplocal, sepdofs, cidx, bdofs, bkeys, bperm, cptr, cg1, zptr, zsrc;
  DBuf<BlockDesc> blocks;               // what the block setup kernel runs over: instances, or distinct blocks (shared form)
  DBuf<TileOp> opsV, opsS, opsDW;
  DBuf<double> g1, rs, us, z;           // V outputs, per-patch separator rhs / solution, shared form: summed us per block
  // Schur-complement setup (ALFIB_SCHUR_SETUP=1; condense_host.h build_schur_lists): X_SS = inverse of
  // A_SS - sum_k A_Sk A_kk^-1 A_kS instead of the S x S cut of the inverse of the whole patch
  bool schur = false;
  SchurHost sh;
  DBuf<int32_t> sepsorted, sepperm, nb_pos, sc_upos, sc_inst_ld, sforder;
  DBuf<int64_t> blk_start, nb_off, sc_coff, sc_inst_c;
  DBuf<double> cbuf;                    // C = A_Nk A_kk^-1 A_kN of every factor block (m x m)
  void release() {
    sepoff.release(); ssoff.release(); seplocal.release(); sepdofs.release(); cidx.release();
    bdofs.release(); bkeys.release(); bperm.release(); cptr.release(); cg1.release(); blocks.release();
    zptr.release(); zsrc.release(); z.release();
    sepsorted.release(); sepperm.release(); nb_pos.release(); sc_upos.release(); sc_inst_ld.release(); sforder.release();
    blk_start.release(); nb_off.release(); sc_coff.release(); sc_inst_c.release(); cbuf.release();
    schur = false;
    opsV.release(); opsS.release(); opsDW.release(); g1.release(); rs.release(); us.release();
  }
};

// One set of patches (the smoother's vertex/macro stars or the transfer's cell patches):
// index sets, colouring, inverse factors in the tiled apply layout, and the warp work list.
struct PatchSet {
  int npatch = 0;
  int maxn = 0;
  int ncolour = 0;
  bool repeated = false;            // a patch occurs twice in the iteration set
  long long lo = 0, hi = 0;         // dof range touched by these patches (peer-memory reduction)
  bool factored = false;
  std::vector<int64_t> h_off;       // npatch+1
  std::vector<int32_t> h_dofs, h_order, h_colour;
  std::vector<int64_t> h_soff;      // element offset of each patch's factor storage (npatch+1)
  std::vector<int> colour_work_start;   // ncolour+1, ranges of the work list
  DBuf<int64_t> off, soff;
  DBuf<int32_t> dofs, sorted, sperm, forder;
  DBuf<int2> work;                  // (patch, tile) per warp, colour-major
  int nwork = 0;
  double* store = nullptr;          // inverse factors, tiled layout
  int64_t store_elems = 0;
  bool store_owned = false;
  DBuf<double> store_buf;
  Condensed cond;                   // block/separator form of the inverses (alfib_level_set_patch_blocks)
  // multiplicative composition (alfib_level_set_sweep_stages): per stage the (patch, tile) work items and the block
  // rows whose residual the stage reads
  int nstage = 0;
  bool symmetric_sweep = false;
  std::vector<int> stage_work_start, stage_row_start;   // nstage + 1
  DBuf<int2> stage_work;
  DBuf<int32_t> stage_rows;
  // patch operators that are not sub-matrices (alfib_level_set_patch_corrections): A_i = A[I_i, I_i] + C_i
  bool has_corr = false, corr_fresh = false;      // fresh: values received since the operator values changed
  DBuf<int64_t> corr_off;
  DBuf<int32_t> corr_rows, corr_cols;
  DBuf<double> corr_vals;
  int64_t corr_nnz = 0;
};

// Exchange lists of one distributed-vector layout (alfib_level_set_halo): local vector = owned dofs
// [0, n_owned) then ghosts [n_owned, n_local).  comm.cu: halo_update / halo_reduce.
struct Halo {
  bool on = false;
  int n_owned = 0, n_local = 0;
  std::vector<int> peers;                    // ascending rank order
  std::vector<int64_t> send_off, recv_off;   // npeers + 1
  DBuf<int32_t> send_idx, recv_idx;          // owned / ghost local dofs, grouped by peer
  DBuf<double> sbuf, rbuf;                   // packed owned-side / ghost-side entries
  // ghost -> owner sum as a gather: distinct owned interface dof d = red_dof[i] receives the packed entries
  // red_src[red_ptr[i] .. red_ptr[i+1]) (positions in the send list, ascending = ascending peer rank)
  int n_red = 0;
  DBuf<int32_t> red_ptr, red_dof, red_src;
  // NVLink peer-memory exchanges: where, in peer p's packed send / ghost buffer, the part for this rank starts
  bool has_peer_off = false;
  std::vector<int64_t> peer_send_off, peer_recv_off;
  int ch_update = -1, ch_reduce = -1;        // mailbox channels of the two directions (comm_mbox_reserve)
  void release() {
    send_idx.release(); recv_idx.release(); sbuf.release(); rbuf.release();
    red_ptr.release(); red_dof.release(); red_src.release();
    on = false;
    has_peer_off = false;
    n_red = 0;
  }
};

// Peer lists of one exchange, passed to the pull kernels by value: segment p of this rank's list, [mine_off[p],
// mine_off[p+1]), is found in rank peers[p]'s packed buffer from theirs_off[p] on.
struct HaloPeers {
  int npeers;
  int peers[ALFIB_MAX_RANKS];
  long long mine_off[ALFIB_MAX_RANKS + 1];
  long long theirs_off[ALFIB_MAX_RANKS];
};

struct Level {
  int n_nodes = 0, bs = 0, n = 0;
  int index = 0;                             // level number (event accounting)
  int n_owned = 0;                           // == n unless the level has a halo
  Halo halo, thalo;                          // own layout; layout in which this level's P_H reads level-1
  DBuf<double> tc;                           // a level-1 vector in the thalo layout
  int64_t nnzb = 0;
  DBuf<int32_t> rowptr, colidx, bc, cb;
  std::vector<int32_t> h_rowptr, h_colidx;   // host copy of the pattern (block structure checks)
  DBuf<double> vals, dvals, a0vals;
  int nbc = 0, ncb = 0;
  bool has_values = false, has_transfer = false, has_d = false;
  PatchSet ps[2];
  // standard prolongation (scalar CSR, fine nodes x coarse nodes) and its transpose
  int p_rows = 0, p_cols = 0, p_bs = 0;     // p_bs = bs: scalar CSR per node; 1: dof-level CSR
  DBuf<int32_t> p_rowptr, p_colidx, pt_rowptr, pt_colidx;
  DBuf<double> p_vals, pt_vals;
  // block rows [row_start[r], row_start[r+1]) of every BSR operator on this level belong to rank r
  std::vector<int64_t> row_start;       // in block rows, nranks+1 entries
  std::vector<int64_t> dof_start;       // the same in scalar dofs
  // cycle work vectors (allocated by alfib_cycle_setup / alfib_smooth)
  DBuf<double> b, x, r, w, t1, t2, t3, t4, V, Z;
  int krylov_m = 0;
};

// Pressure-side operators of the outer Schur-complement fieldsplit on the finest level (outer.cu; alfib_schur_set)
struct Schur {
  bool on = false;
  int nu = 0, np = 0;                      // velocity dofs of the finest level, pressure dofs
  int remove_mean = 0;                     // constant-pressure nullspace removed after the Schur solve
  int64_t b_nnz = 0, mi_nnz = 0;
  DBuf<int32_t> b_rowptr, b_colidx, bt_rowptr, bt_colidx, mi_rowptr, mi_colidx;
  DBuf<double> b_vals, bt_vals, mi_vals;   // B (np x nu), its explicit transpose, M_p^-1 (np x np), scalar CSR
  DBuf<double> y1, tu, tp, tp2;            // work vectors of one preconditioner application
  DBuf<double> V, Z, w, r, xs, hd;         // outer FGMRES: bases, work vectors, device scalars
  int restart = 0;
};

struct EventRec;
struct alfib_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int deterministic = 0, sync_always = 0, robust_restrict = 1, transfer_refine = 1, use_graph = 1;
  void* graph_exec = nullptr;        // captured F-cycle (cudaGraphExec_t)
  int cycles_run = 0;
  int64_t graph_launches = 0;
  Level* levels[ALFIB_MAX_LEVELS] = {nullptr};
  int nlevels = 0, smoothing = 0;
  int num_sms = 148;
  int64_t launches = 0;
  // multi-GPU (comm.cu): NCCL communicator, rank layout
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  // peer-memory exchanges (comm.cu): symmetric buffer = header + 2 slots of sym_stride doubles
  unsigned char* sym = nullptr;
  size_t sym_stride = 0;
  bool peers_open = false;
  void* peer_ptr[ALFIB_MAX_RANKS] = {nullptr};
  DBuf<double*> d_peer_slot;             // device array: slot 0 of every rank
  DBuf<unsigned long long> d_epoch;      // exchange counter (device)
  DBuf<int> d_comm_err;
  DBuf<long long> d_gate;                // block 0 -> other blocks gate of the reduction kernel
  // mailbox transport (distributed vectors): this rank's channel table, the arena size reserved so far (relative
  // to the arena start until the buffer is allocated), the peers' tables, per-channel exchange numbers and counters
  MboxEntry mbox[ALFIB_MBOX_CHANNELS] = {};
  size_t mbox_bytes = 0;
  bool mbox_fixed = false;               // offsets made absolute (buffer allocated)
  std::vector<MboxEntry> mbox_peer[ALFIB_MAX_RANKS];
  DBuf<unsigned long long> mbox_seq;     // [channel]: number of completed exchanges
  DBuf<unsigned int> mbox_cnt;           // [2 * channel]: blocks that have pushed / finished
  // patch factor workspace (one slot per resident CTA) + status word + work counter
  DBuf<double> fwork;
  DBuf<int> finfo;
  DBuf<unsigned> tile_counter;          // work counter of the TMA tile-op kernel (resets itself)
  // reductions / Krylov scalars
  DBuf<double> partial, scal;
  // staging for host-pointer calls
  DBuf<double> stage_in, stage_in2, stage_out;
  // coarse solve
  cusolverDnHandle_t cusolver = nullptr;
  DBuf<double> coarse_lu, coarse_work, coarse_inv, coarse_partial, coarse_r, coarse_dx;
  int64_t coarse_ld = 0;
  DBuf<int> coarse_piv, coarse_info;
  DBuf<int32_t> coarse_seppos;
  int coarse_n = 0;
  bool coarse_factored = false;
  Schur schur;                          // outer fieldsplit pieces (alfib_schur_set)
  std::vector<void*> host_registered;   // caller buffers page-locked through alfib_host_register
  // profiling
  int profile = 0;
  double ev_ms[ALFIB_MAX_LEVELS][ALFIB_EV_COUNT] = {{0}};
  int64_t ev_calls[ALFIB_MAX_LEVELS][ALFIB_EV_COUNT] = {{0}};
  std::vector<struct EventRec> ev_pool;
  int ev_used = 0;
};

// CUDA-event timing of one kernel family on one level (opt-in).  Events are recorded on the
// ctx stream without synchronising and resolved lazily (profile_flush), so profiling does not
// perturb the timed region beyond the event records themselves.
struct EventRec {
  cudaEvent_t a, b;
  int id, level;
};
void profile_flush(alfib_ctx* c);

struct ScopedEvent {
  alfib_ctx* c;
  int slot = -1;
  ScopedEvent(alfib_ctx* ctx, int ev, int level = 0) : c(ctx) {
    if (!c->profile) return;
    if (c->ev_used == (int)c->ev_pool.size()) {
      if (c->ev_pool.size() >= 8192) {
        profile_flush(c);
      } else {
        EventRec r;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        c->ev_pool.push_back(r);
      }
    }
    slot = c->ev_used++;
    c->ev_pool[slot].id = ev;
    c->ev_pool[slot].level = level;
    cudaEventRecord(c->ev_pool[slot].a, c->stream);
  }
  ~ScopedEvent() {
    if (slot >= 0) cudaEventRecord(c->ev_pool[slot].b, c->stream);
  }
};

inline int roundup2(int n) { return (n + 1) & ~1; }
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- kernels (implemented in the .cu files; all enqueue on ctx->stream) ----------------------
// spmv.cu
void launch_bsr_spmv(alfib_ctx* c, const Level& L, const double* vals, const double* x, double* y,
                     const double* b /* nullptr: y = A x ; else y = b - A x */);
void launch_csr_apply(alfib_ctx* c, int nrows, int bs, const int32_t* rowptr, const int32_t* colidx,
                      const double* vals, const double* x, double* y, int64_t nnz);
void launch_transpose_blocks(alfib_ctx* c, double* vals, int64_t nnzb, int bs);
// comm.cu
void comm_unique_id(void* out128);
void comm_init(alfib_ctx* c, const void* id128, int rank, int nranks);
void comm_destroy(alfib_ctx* c);
void comm_allreduce_sum(alfib_ctx* c, double* y, size_t n);
void comm_allgather_rows(alfib_ctx* c, double* y, const std::vector<int64_t>& dof_start);
// distributed vectors: owner -> ghost copy / ghost -> owner sum (ghosts cleared) between neighbour ranks
void halo_update(alfib_ctx* c, Halo& H, double* x, int level);
void halo_reduce(alfib_ctx* c, Halo& H, double* y, int level);
// v[0..nv) summed over the ranks, result on every rank (FGMRES dots); sqrt_mode: v[0] = sqrt(sum), inv = 1 / v[0]
void comm_small_allreduce(alfib_ctx* c, double* v, int nv, int sqrt_mode, double* inv);
bool comm_small_allreduce_partials(alfib_ctx* c, const double* partial, int nparts, double* v, int nv, int sqrt_mode, double* inv);
void comm_mbox_reserve(alfib_ctx* c, Halo& H, int level, int which);    // channels of a halo (before the peer buffer exists)
void comm_peer_alloc(alfib_ctx* c);
void comm_peer_handle(alfib_ctx* c, void* out64);
void comm_peer_open(alfib_ctx* c, const void* handles);
void comm_peer_publish_ranges(alfib_ctx* c);
void comm_peer_close(alfib_ctx* c);
int comm_peer_error(alfib_ctx* c);
PeerOut comm_peer_out(alfib_ctx* c);                       // the current symmetric slot as a kernel destination
void comm_peer_zero(alfib_ctx* c, long long lo, long long hi);
// y[i] = sum over ranks q with lo_q <= i < hi_q of slot_q[i]; hdr_slot >= 0: ranges from the peers'
// headers, else the explicit ranges
void comm_peer_reduce(alfib_ctx* c, int64_t n, int hdr_slot, const long long* lo, const long long* hi, double* y);
// patch_apply.cu
void launch_patch_apply(alfib_ctx* c, const PatchSet& ps, const double* x, PeerOut y);
// multiplicative sweep(s) over the stages of alfib_level_set_sweep_stages: y = 0; per stage r = x - A y on the rows
// the stage reads, y[I_i] += A_i^-1 r[I_i]
void patch_apply_multiplicative(alfib_ctx* c, Level& L, int which, const double* x, double* y);
void launch_bsr_residual_rows(alfib_ctx* c, const Level& L, const double* vals, const int32_t* rows, int nrows,
                              const double* x, const double* y, double* r);
// y = sum over ranks of this rank's patch contributions (zeroing, exchange included)
// (a level with a halo: the ghosts of x are refreshed first — x is a local work vector there — and the ghost
// contributions of y are summed into their owners; y is valid on the owned entries only)
void patch_apply_sum(alfib_ctx* c, Level& L, int level, int which, const double* x, double* y);
// condense.cu
void condense_setup(alfib_ctx* c, Level& L, PatchSet& ps, const int32_t* block_of_dof);
void launch_condense_blocks(alfib_ctx* c, const Level& L, PatchSet& ps, const double* vals, bool schur = false);
void launch_condensed_apply(alfib_ctx* c, const PatchSet& ps, const double* x, PeerOut y);
void condensed_extract_inverse(alfib_ctx* c, const PatchSet& ps, int patch, double* host_out);
// patch_factor.cu
void launch_patch_factor(alfib_ctx* c, const Level& L, PatchSet& ps, const double* vals);
void patch_extract_inverse(alfib_ctx* c, const PatchSet& ps, int patch, double* host_out);
// vector.cu
void launch_set_rows(alfib_ctx* c, double* y, const double* x /* nullptr: zero */, const int32_t* idx, int nidx);
void launch_axpby(alfib_ctx* c, int n, double a, const double* x, double b, double* y);   // y = a x + b y
void launch_sub(alfib_ctx* c, int n, const double* a, const double* b, double* out);       // out = a - b
void launch_bsr_to_dense(alfib_ctx* c, const Level& L, double* dense /* col-major n x n */);
// krylov.cu
void fgmres_device(alfib_ctx* c, Level& L, int level, int m, const double* b, double* x);
void krylov_reserve(alfib_ctx* c);
void launch_multi_dot(alfib_ctx* c, int n, int nv, const double* V, int64_t ldv, const double* w, double* out);
void launch_maxpy_norm(alfib_ctx* c, int n, int nv, const double* coef, double sign, const double* V, int64_t ldv,
                       double* w, double* nrm, double* inv);
void launch_scale_by(alfib_ctx* c, int n, const double* scale, const double* in, double* out);
// outer.cu
void schur_apply_device(alfib_ctx* c, double nu, double gamma, const double* r, double* y);
void jacobian_apply_device(alfib_ctx* c, const double* z, double* out);
void outer_solve_device(alfib_ctx* c, double nu, double gamma, const double* b, double* x, double rtol, double atol,
                        int maxit, int restart, int* iterations, double* history, int nhistory);
// cycle.cu
void smoother_apply_device(alfib_ctx* c, Level& L, int level, const double* x, double* y);
void prolong_device(alfib_ctx* c, Level& Lf, int level, const double* coarse, double* fine);
void restrict_device(alfib_ctx* c, Level& Lf, Level& Lc, int level, const double* fine, double* coarse);
void coarse_factor_device(alfib_ctx* c);
void coarse_solve_device(alfib_ctx* c, const double* b, double* x);
void cycle_apply_device(alfib_ctx* c, const double* b, double* x);
void cycle_graph_invalidate(alfib_ctx* c);
