// fieldsplit_0 = richardson(1) + PCMG "full" with V inner cycles, FGMRES(m)/PatchPC smoothing on
// every level > 0, Schoeberl transfers, direct coarse solve (alfi/solver.py:359-379;
// alfi/transfer.py:186-275; PCMG semantics restated in SURVEY Appendix A.5/A.6).
#include <algorithm>

#include "alfib_internal.h"

#define CUSOLVER_TRY(expr)                                                                      \
  do {                                                                                          \
    cusolverStatus_t _s = (expr);                                                               \
    if (_s != CUSOLVER_STATUS_SUCCESS)                                                          \
      throw DeviceError{ALFIB_ECUDA, std::string(#expr) + ": cusolver status " + std::to_string((int)_s)}; \
  } while (0)

// PCApply_PATCH (additive): y = sum_i R_i^T A_i^-1 R_i x ; y[bc] = x[bc]
void smoother_apply_device(alfib_ctx* c, Level& L, int level, const double* x, double* y) {
  PatchSet& ps = L.ps[ALFIB_PATCHES_SMOOTHER];
  ALFIB_REQUIRE(ps.factored, "alfib_level_factor has not been called for this level");
  ScopedEvent ev(c, ALFIB_EV_PCPATCH_APPLY, level);
  patch_apply_sum(c, L, level, ALFIB_PATCHES_SMOOTHER, x, y);     // incl. the multi-GPU exchange
  launch_set_rows(c, y, x, L.bc.p, L.nbc);
}

// transfer.py:254-257 — PatchPC apply of the transfer solver: y = blockdiag(A0)^-1 b on the
// (disjoint) cell patches, y[cb] = b[cb]
static void cell_block_solve(alfib_ctx* c, Level& L, int level, const double* b, double* y) {
  PatchSet& ps = L.ps[ALFIB_PATCHES_TRANSFER];
  ALFIB_REQUIRE(ps.factored, "alfib_transfer_update has not been called for this level");
  patch_apply_sum(c, L, level, ALFIB_PATCHES_TRANSFER, b, y);
  launch_set_rows(c, y, b, L.cb.p, L.ncb);
}

// The same solve followed by one step of iterative refinement, t += S (b - A0 t).  The explicit
// inverse applied as a GEMV loses ~eps*|A0^-1|*|b| — and here |b| = |gamma D rhs| is ~1e4..1e6
// times |t| — while the reference's LU solve (transfer.py:112-113) does not; one refinement
// step restores LU-level accuracy for one extra SpMV + block apply (measured: 1.7e-6 -> 6e-11
// on the 3-D SV k=3 cell patches at gamma = 1e4, nu = 0.02).
static void cell_block_solve_refined(alfib_ctx* c, Level& L, int level, const double* b, double* y) {
  cell_block_solve(c, L, level, b, y);
  if (!c->transfer_refine || !L.a0vals.p) return;
  L.t3.alloc(L.n);
  L.t4.alloc(L.n);
  launch_bsr_spmv(c, L, L.a0vals.p, y, L.t3.p, b);                     // r = b - A0 y
  patch_apply_sum(c, L, level, ALFIB_PATCHES_TRANSFER, L.t3.p, L.t4.p);  // dy = S r (patch rows only)
  launch_axpby(c, L.n, 1.0, L.t4.p, 1.0, y);
}

// fine = (I - A0^-1 gamma D) P_H coarse, fine[bc] = 0   (transfer.py:246-259; Appendix A.6)
void prolong_device(alfib_ctx* c, Level& L, int level, const double* coarse, double* fine) {
  ALFIB_REQUIRE(L.has_transfer, "alfib_transfer_set has not been called for this level");
  ScopedEvent ev(c, ALFIB_EV_PROLONG, level);
  L.t1.alloc(L.n);
  L.t2.alloc(L.n);
  L.r.alloc(L.n);
  double* rhs = L.r.p;
  // Distributed vectors (alfib_level_set_halo; the sequence of oracle/distributed.py DistHierarchy.prolong): P_H
  // holds this rank's owned fine rows; its columns are the coarser level's vector in the transfer-halo layout
  // (owned part copied, ghosts fetched from their owners) or, with a replicated coarser level, that vector
  // itself.  Every later step produces owned entries; the gathers refresh the ghosts they read.
  const double* csrc = coarse;
  if (L.thalo.on) {
    L.tc.alloc(L.thalo.n_local);
    CUDA_TRY(cudaMemcpyAsync(L.tc.p, coarse, sizeof(double) * L.thalo.n_owned, cudaMemcpyDeviceToDevice, c->stream));
    halo_update(c, L.thalo, L.tc.p, level);
    csrc = L.tc.p;
  }
  launch_csr_apply(c, L.p_rows, L.p_bs, L.p_rowptr.p, L.p_colidx.p, L.p_vals.p, csrc, rhs, (int64_t)L.p_vals.n);
  if (L.has_d) {
    launch_bsr_spmv(c, L, L.dvals.p, rhs, L.t1.p, nullptr);          // b = gamma D rhs
    launch_set_rows(c, L.t1.p, nullptr, L.cb.p, L.ncb);              // coarse-boundary rows zeroed
    cell_block_solve_refined(c, L, level, L.t1.p, L.t2.p);           // t = A0^-1 b
    launch_sub(c, L.n, rhs, L.t2.p, fine);                           // fine = rhs - t
  } else {
    CUDA_TRY(cudaMemcpyAsync(fine, rhs, sizeof(double) * L.n, cudaMemcpyDeviceToDevice, c->stream));
  }
  launch_set_rows(c, fine, nullptr, L.bc.p, L.nbc);
}

// coarse = P_H^T (I - gamma D A0^-1) fine, coarse[bc_c] = 0   (transfer.py:261-275; A.6)
void restrict_device(alfib_ctx* c, Level& L, Level& Lc, int level, const double* fine, double* coarse) {
  ALFIB_REQUIRE(L.has_transfer, "alfib_transfer_set has not been called for this level");
  ScopedEvent ev(c, ALFIB_EV_RESTRICT, level);
  const double* src = fine;
  if (L.has_d && c->robust_restrict) {
    L.t1.alloc(L.n);
    L.t2.alloc(L.n);
    CUDA_TRY(cudaMemcpyAsync(L.t1.p, fine, sizeof(double) * L.n, cudaMemcpyDeviceToDevice, c->stream));
    launch_set_rows(c, L.t1.p, nullptr, L.cb.p, L.ncb);              // bcs.apply(tildeu)
    cell_block_solve_refined(c, L, level, L.t1.p, L.t2.p);           // r = A0^-1 t
    launch_bsr_spmv(c, L, L.dvals.p, L.t2.p, L.t1.p, nullptr);       // b = gamma D r (no bcs)
    launch_sub(c, L.n, fine, L.t1.p, L.t2.p);                        // r2 = fine - b
    src = L.t2.p;
  }
  if (L.thalo.on) {
    // P_H^T of the owned fine rows lands on the transfer-halo layout of the coarser level; the ghost parts are
    // summed into their owners (DistHierarchy.restrict)
    L.tc.alloc(L.thalo.n_local);
    launch_csr_apply(c, L.p_cols, L.p_bs, L.pt_rowptr.p, L.pt_colidx.p, L.pt_vals.p, src, L.tc.p, (int64_t)L.pt_vals.n);
    halo_reduce(c, L.thalo, L.tc.p, level);
    CUDA_TRY(cudaMemcpyAsync(coarse, L.tc.p, sizeof(double) * L.thalo.n_owned, cudaMemcpyDeviceToDevice, c->stream));
  } else {
    launch_csr_apply(c, L.p_cols, L.p_bs, L.pt_rowptr.p, L.pt_colidx.p, L.pt_vals.p, src, coarse, (int64_t)L.pt_vals.n);
    // replicated coarser level under a distributed fine level: every rank holds the part of its owned fine rows
    if (L.halo.on && c->nranks > 1) comm_allreduce_sum(c, coarse, (size_t)Lc.n);
  }
  launch_set_rows(c, coarse, nullptr, Lc.bc.p, Lc.nbc);
}

// Coarse level: explicit dense inverse, applied as a row-sharded streaming GEMV.
// Setup (per Newton step): BSR -> dense, cuSOLVER getrf + getrs against the identity (the north
// star allows a gathered dense LU for the coarsest level; replaces AssembledPC + telescope +
// superlu_dist, solver.py:369-378).  Apply: x = Ainv b, then one step of iterative refinement
// x += Ainv (b - A x) so the result is as accurate as an LU solve.  The GEMV reads the
// column-major inverse with 128-bit loads, 64 rows per warp, the column range split over the
// warps of a CTA and over KSPLIT CTAs; partial sums go through a fixed-order two-pass reduction
// (deterministic).  HBM-bound: 8 n^2 bytes per application, and it shards by rows over ranks.
namespace {

constexpr int KSPLIT = 8;          // column chunks per row tile (grid.y)
constexpr int GW = 8;              // warps per CTA, each takes an interleaved part of the chunk

__global__ void set_identity_kernel(double* B, int64_t n, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) B[i + i * ld] = 1.0;
}

__global__ void __launch_bounds__(GW * 32) dense_gemv_kernel(const double* __restrict__ A, int64_t ld, int n,
                                                             int tile0, const double* __restrict__ x,
                                                             double* __restrict__ partial) {
  __shared__ double red[GW][ALFIB_TILE_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = (tile0 + blockIdx.x) * ALFIB_TILE_ROWS;
  const int r = row0 + 2 * lane;
  const bool active = r < n;                      // ld is even and >= n+1 when n is odd: r+1 < ld
  const int per = (n + KSPLIT - 1) / KSPLIT;
  const int c0 = blockIdx.y * per, c1 = min(n, c0 + per);
  double acc0 = 0.0, acc1 = 0.0;
  const double2* __restrict__ Ap = reinterpret_cast<const double2*>(A + r);
  if (active) {
#pragma unroll 4
    for (int c = c0 + warp; c < c1; c += GW) {
      const double2 a = __ldcs(Ap + (int64_t)c * (ld >> 1));
      const double xc = __ldg(x + c);
      acc0 = fma(a.x, xc, acc0);
      acc1 = fma(a.y, xc, acc1);
    }
  }
  red[warp][2 * lane] = acc0;
  red[warp][2 * lane + 1] = acc1;
  __syncthreads();
  if (threadIdx.x < ALFIB_TILE_ROWS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < GW; ++w) v += red[w][threadIdx.x];
    const int rr = row0 + threadIdx.x;
    if (rr < n) partial[(int64_t)blockIdx.y * n + rr] = v;
  }
}

// y[i] (+)= sum_k partial[k*n + i] for rows [r0, r1)
__global__ void gemv_reduce_kernel(int n, int r0, int r1, const double* __restrict__ partial, PeerOut yout,
                                   const double* __restrict__ yold) {
  double* __restrict__ y = resolve(yout);
  const int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  double v = 0.0;
#pragma unroll
  for (int k = 0; k < KSPLIT; ++k) v += partial[(int64_t)k * n + i];
  y[i] = yold ? yold[i] + v : v;
}

void coarse_gemv(alfib_ctx* c, const double* b, double* y, int accumulate) {
  const int n = c->coarse_n;
  const int ntile = (n + ALFIB_TILE_ROWS - 1) / ALFIB_TILE_ROWS;
  const int t0 = (int)((int64_t)ntile * c->rank / c->nranks), t1 = (int)((int64_t)ntile * (c->rank + 1) / c->nranks);
  const int r0 = std::min(n, t0 * ALFIB_TILE_ROWS), r1 = std::min(n, t1 * ALFIB_TILE_ROWS);
  const bool peer = c->nranks > 1 && c->peers_open;
  if (t1 > t0) {
    dense_gemv_kernel<<<dim3(t1 - t0, KSPLIT), GW * 32, 0, c->stream>>>(c->coarse_inv.p, c->coarse_ld, n, t0, b,
                                                                        c->coarse_partial.p);
    gemv_reduce_kernel<<<cdiv(r1 - r0, 256), 256, 0, c->stream>>>(n, r0, r1, c->coarse_partial.p,
                                                                   peer ? comm_peer_out(c) : plain_out(y),
                                                                   accumulate ? y : nullptr);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
  }
  if (c->nranks > 1) {
    std::vector<int64_t> start(c->nranks + 1);
    for (int r = 0; r <= c->nranks; ++r)
      start[r] = std::min<int64_t>(n, (int64_t)ntile * r / c->nranks * ALFIB_TILE_ROWS);
    if (peer) {
      long long lo[ALFIB_MAX_RANKS], hi[ALFIB_MAX_RANKS];
      for (int r = 0; r < c->nranks; ++r) { lo[r] = start[r]; hi[r] = start[r + 1]; }
      comm_peer_reduce(c, n, -1, lo, hi, y);
    } else {
      comm_allgather_rows(c, y, start);
    }
  }
}

}  // namespace

namespace {

// Condensed coarse inverse: the coarse level registered as ONE patch with macro-cell blocks on
// (level 0, ALFIB_PATCHES_SMOOTHER) — X_SS = inv[S, S] of the dense inverse in the 64-row tile layout of
// condense.cu; D, V, W of the blocks come from condense_blocks_kernel.  n = number of separator dofs.
__global__ void extract_xss_kernel(const double* __restrict__ inv, int64_t ld, const int32_t* __restrict__ sepdofs,
                                   int ns, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)ns * ns) return;
  const int r = (int)(e % ns), col = (int)(e / ns);                 // r fastest: consecutive rows of one column
  const int t = r / ALFIB_TILE_ROWS, row0 = t * ALFIB_TILE_ROWS;
  const int rows = min(ns - row0, ALFIB_TILE_ROWS), rt = (rows + 1) & ~1;
  const int64_t sr = sepdofs ? sepdofs[r] : r, sc = sepdofs ? sepdofs[col] : col;
  out[(int64_t)row0 * ns + (int64_t)col * rt + (r - row0)] = inv[sr + ld * sc];
}

// Schur-complement setup of the condensed coarse inverse (ALFIB_SCHUR_SETUP, condense_host.h): S_c = A_SS - sum_k C_k
// assembled densely (ns x ns), inverted by cuSOLVER — 2.7 ns^3 flops instead of 2.7 n^3 (cfg5: 6 591 of 23 871 dofs,
// 47x fewer).  seppos[g] = position of dof g in the separator list or -1.
__global__ void seppos_kernel(int ns, const int32_t* __restrict__ sepdofs, int32_t* __restrict__ seppos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ns) seppos[sepdofs[i]] = i;
}

__global__ void schur_gather_kernel(int nbrows, int bs, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                    const double* __restrict__ vals, const int32_t* __restrict__ seppos, int64_t ld,
                                    double* __restrict__ Sc) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nbrows) return;
  const int b2 = bs * bs;
  for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
    const int cn = colidx[k];
    for (int r = 0; r < bs; ++r) {
      const int pr = seppos[row * bs + r];
      if (pr < 0) continue;
      for (int c2 = 0; c2 < bs; ++c2) {
        const int pc = seppos[cn * bs + c2];
        if (pc >= 0) Sc[pr + ld * pc] = vals[(int64_t)k * b2 + r * bs + c2];
      }
    }
  }
}

// Sc[nb_pos[i], nb_pos[j]] -= C_q[upos[i], upos[j]] for every block instance q (one CTA each); instances overlap on
// the separator, hence atomics (the sum order is not fixed: the deterministic mode keeps the dense setup)
__global__ void schur_subtract_kernel(int64_t ninst, const int64_t* __restrict__ nb_off, const int32_t* __restrict__ nb_pos,
                                      const int32_t* __restrict__ upos, const int64_t* __restrict__ inst_c,
                                      const int32_t* __restrict__ inst_ld, const double* __restrict__ cbuf, int64_t ld,
                                      double* __restrict__ Sc) {
  for (int64_t q = blockIdx.x; q < ninst; q += gridDim.x) {
    const int64_t o2 = nb_off[q];
    const int mq = (int)(nb_off[q + 1] - o2);
    const double* __restrict__ C = cbuf + inst_c[q];
    const int ldc = inst_ld[q];
    for (int i = threadIdx.x; i < mq * mq; i += blockDim.x) {
      const int jj = i / mq, ii = i - jj * mq;
      atomicAdd(Sc + nb_pos[o2 + ii] + ld * nb_pos[o2 + jj], -C[upos[o2 + ii] + (int64_t)upos[o2 + jj] * ldc]);
    }
  }
}

bool coarse_is_condensed(const Level& L0) {
  const PatchSet& ps = L0.ps[ALFIB_PATCHES_SMOOTHER];
  return ps.npatch == 1 && ps.cond.on;
}

}  // namespace

void coarse_factor_device(alfib_ctx* c) {
  Level* L0 = c->levels[0];
  ALFIB_REQUIRE(L0 && L0->has_values, "level 0 has no operator values");
  ScopedEvent ev(c, ALFIB_EV_COARSE, 0);
  if (!c->cusolver) {
    CUSOLVER_TRY(cusolverDnCreate(&c->cusolver));
    CUSOLVER_TRY(cusolverDnSetStream(c->cusolver, c->stream));
  }
  const int n = L0->n;
  const int64_t ld = roundup2(n);
  c->coarse_n = n;
  c->coarse_ld = ld;
  c->coarse_partial.alloc((size_t)KSPLIT * n);
  c->coarse_r.alloc(n);
  c->coarse_dx.alloc(n);
  if (coarse_is_condensed(*L0) && L0->ps[ALFIB_PATCHES_SMOOTHER].cond.schur && !c->deterministic &&
      L0->ps[ALFIB_PATCHES_SMOOTHER].cond.h.nsep_total > 0) {
    // ---- Schur-complement setup: only the separator system is factorised densely -------------------------
    PatchSet& ps = L0->ps[ALFIB_PATCHES_SMOOTHER];
    Condensed& cd = ps.cond;
    const CondensedHost& h = cd.h;
    if (!ps.store) {
      ps.store_buf.alloc((size_t)std::max<int64_t>(ps.store_elems, 2));
      ps.store = ps.store_buf.p;
      ps.store_owned = true;
    }
    launch_condense_blocks(c, *L0, ps, L0->vals.p, true);            // D, V, -Ws tiles and C = A_Nk Ws per block
    const int ns = (int)h.nsep_total;
    const int64_t lds = roundup2(ns);
    c->coarse_lu.alloc((size_t)lds * ns);                            // S_c, then its LU
    c->coarse_inv.alloc((size_t)lds * ns);                           // identity, then X_SS (column-major)
    c->coarse_piv.alloc(ns);
    c->coarse_info.alloc(1);
    DBuf<int32_t>& seppos = c->coarse_seppos;               // kept: a cudaMalloc / cudaFree pair per Newton step otherwise
    seppos.alloc(n);
    CUDA_TRY(cudaMemsetAsync(seppos.p, 0xff, sizeof(int32_t) * n, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->coarse_lu.p, 0, sizeof(double) * lds * ns, c->stream));
    seppos_kernel<<<cdiv(ns, 256), 256, 0, c->stream>>>(ns, cd.sepdofs.p, seppos.p);
    schur_gather_kernel<<<cdiv(L0->n_nodes, 8), 256, 0, c->stream>>>(L0->n_nodes, L0->bs, L0->rowptr.p, L0->colidx.p,
                                                                     L0->vals.p, seppos.p, lds, c->coarse_lu.p);
    schur_subtract_kernel<<<(int)std::min<int64_t>(std::max<int64_t>(h.nblocks, 1), 4 * c->num_sms), 256, 0, c->stream>>>(
        h.nblocks, cd.nb_off.p, cd.nb_pos.p, cd.sc_upos.p, cd.sc_inst_c.p, cd.sc_inst_ld.p, cd.cbuf.p, lds, c->coarse_lu.p);
    c->launches += 3;
    CUDA_TRY(cudaGetLastError());
    int lwork = 0;
    CUSOLVER_TRY(cusolverDnDgetrf_bufferSize(c->cusolver, ns, ns, c->coarse_lu.p, (int)lds, &lwork));
    c->coarse_work.alloc(lwork);
    CUSOLVER_TRY(cusolverDnDgetrf(c->cusolver, ns, ns, c->coarse_lu.p, (int)lds, c->coarse_work.p, c->coarse_piv.p,
                                  c->coarse_info.p));
    int info = 0;
    CUDA_TRY(cudaMemcpyAsync(&info, c->coarse_info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (info != 0) throw DeviceError{ALFIB_ESINGULAR, "coarse Schur complement LU failed, info = " + std::to_string(info)};
    CUDA_TRY(cudaMemsetAsync(c->coarse_inv.p, 0, sizeof(double) * lds * ns, c->stream));
    set_identity_kernel<<<cdiv(ns, 256), 256, 0, c->stream>>>(c->coarse_inv.p, ns, lds);
    CUSOLVER_TRY(cusolverDnDgetrs(c->cusolver, CUBLAS_OP_N, ns, ns, c->coarse_lu.p, (int)lds, c->coarse_piv.p,
                                  c->coarse_inv.p, (int)lds, c->coarse_info.p));
    CUDA_TRY(cudaMemsetAsync(ps.store + h.ssoff[0], 0, sizeof(double) * (size_t)(h.ssoff[1] - h.ssoff[0]), c->stream));
    extract_xss_kernel<<<cdiv((int64_t)ns * ns, 256), 256, 0, c->stream>>>(c->coarse_inv.p, lds, nullptr, ns,
                                                                           ps.store + h.ssoff[0]);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    ps.factored = true;
    // transient buffers: small ones are kept for the next Newton step (cudaFree / cudaMalloc per call cost more than
    // the factorisation of a small coarse level)
    if ((size_t)lds * ns * sizeof(double) > ((size_t)256 << 20)) {
      c->coarse_inv.release();
      c->coarse_work.release();
      c->coarse_lu.release();
    }
    c->coarse_factored = true;
    return;
  }
  c->coarse_lu.alloc((size_t)n * n);
  c->coarse_inv.alloc((size_t)ld * n);
  c->coarse_piv.alloc(n);
  c->coarse_info.alloc(1);
  c->coarse_partial.alloc((size_t)KSPLIT * n);
  c->coarse_r.alloc(n);
  c->coarse_dx.alloc(n);
  launch_bsr_to_dense(c, *L0, c->coarse_lu.p);
  int lwork = 0;
  CUSOLVER_TRY(cusolverDnDgetrf_bufferSize(c->cusolver, n, n, c->coarse_lu.p, n, &lwork));
  c->coarse_work.alloc(lwork);
  CUSOLVER_TRY(cusolverDnDgetrf(c->cusolver, n, n, c->coarse_lu.p, n, c->coarse_work.p, c->coarse_piv.p,
                                c->coarse_info.p));
  int info = 0;
  CUDA_TRY(cudaMemcpyAsync(&info, c->coarse_info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (info != 0) throw DeviceError{ALFIB_ESINGULAR, "coarse LU failed, info = " + std::to_string(info)};
  CUDA_TRY(cudaMemsetAsync(c->coarse_inv.p, 0, sizeof(double) * ld * n, c->stream));
  set_identity_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(c->coarse_inv.p, n, ld);
  c->launches++;
  CUSOLVER_TRY(cusolverDnDgetrs(c->cusolver, CUBLAS_OP_N, n, n, c->coarse_lu.p, n, c->coarse_piv.p, c->coarse_inv.p,
                                (int)ld, c->coarse_info.p));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (coarse_is_condensed(*L0)) {
    PatchSet& ps = L0->ps[ALFIB_PATCHES_SMOOTHER];
    const CondensedHost& h = ps.cond.h;
    if (!ps.store) {
      ps.store_buf.alloc((size_t)std::max<int64_t>(ps.store_elems, 2));
      ps.store = ps.store_buf.p;
      ps.store_owned = true;
    }
    const int ns = (int)h.nsep_total;
    if (ns > 0) {
      CUDA_TRY(cudaMemsetAsync(ps.store + h.ssoff[0], 0, sizeof(double) * (size_t)(h.ssoff[1] - h.ssoff[0]), c->stream));
      extract_xss_kernel<<<cdiv((int64_t)ns * ns, 256), 256, 0, c->stream>>>(c->coarse_inv.p, ld, ps.cond.sepdofs.p, ns,
                                                                             ps.store + h.ssoff[0]);
      c->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    launch_condense_blocks(c, *L0, ps, L0->vals.p);
    ps.factored = true;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if ((size_t)ld * n * sizeof(double) > ((size_t)256 << 20))
      c->coarse_inv.release();                    // 8 n^2 bytes: the condensed pieces replace it
  }
  // the LU and its workspace are transient; small ones are kept for the next Newton step (cudaFree /
  // cudaMalloc per call cost more than the factorisation of a small coarse level)
  if ((size_t)n * n * sizeof(double) > ((size_t)256 << 20)) {
    c->coarse_work.release();
    c->coarse_lu.release();
  }
  c->coarse_factored = true;
}

void coarse_solve_device(alfib_ctx* c, const double* b, double* x) {
  ALFIB_REQUIRE(c->coarse_factored, "alfib_coarse_factor has not been called");
  ALFIB_REQUIRE(x != b, "coarse solve: x and b must not alias");
  Level& L0 = *c->levels[0];
  ScopedEvent ev(c, ALFIB_EV_COARSE, 0);
  if (coarse_is_condensed(L0)) {
    // every rank applies the whole condensed inverse (0.4 GB at cfg5): no exchange
    PatchSet& ps = L0.ps[ALFIB_PATCHES_SMOOTHER];
    auto apply = [&](const double* src, double* dst) {
      CUDA_TRY(cudaMemsetAsync(dst, 0, sizeof(double) * L0.n, c->stream));
      launch_patch_apply(c, ps, src, plain_out(dst));
      launch_set_rows(c, dst, src, L0.bc.p, L0.nbc);                    // identity rows of the Dirichlet dofs
    };
    apply(b, x);                                                        // x = Ainv b
    launch_bsr_spmv(c, L0, L0.vals.p, x, c->coarse_r.p, b);             // r = b - A x
    apply(c->coarse_r.p, c->coarse_dx.p);
    launch_axpby(c, L0.n, 1.0, c->coarse_dx.p, 1.0, x);                 // x += Ainv r
    return;
  }
  coarse_gemv(c, b, x, 0);                                              // x = Ainv b
  launch_bsr_spmv(c, L0, L0.vals.p, x, c->coarse_r.p, b);               // r = b - A x
  coarse_gemv(c, c->coarse_r.p, x, 1);                                  // x += Ainv r
}

// One V visit on level l: b, x are the level's own vectors (Appendix A.5)
static void vcycle(alfib_ctx* c, int l) {
  Level& L = *c->levels[l];
  if (l == 0) {
    coarse_solve_device(c, L.b.p, L.x.p);
    return;
  }
  Level& Lc = *c->levels[l - 1];
  fgmres_device(c, L, l, c->smoothing, L.b.p, L.x.p);                         // pre-smooth
  {
    ScopedEvent ev(c, ALFIB_EV_MATMULT, l);
    launch_bsr_spmv(c, L, L.vals.p, L.x.p, L.w.p, L.b.p);                     // r = b - A x
  }
  restrict_device(c, L, Lc, l, L.w.p, Lc.b.p);
  CUDA_TRY(cudaMemsetAsync(Lc.x.p, 0, sizeof(double) * Lc.n, c->stream));
  vcycle(c, l - 1);
  prolong_device(c, L, l, Lc.x.p, L.w.p);
  launch_axpby(c, L.n, 1.0, L.w.p, 1.0, L.x.p);                               // x += P x_c
  fgmres_device(c, L, l, c->smoothing, L.b.p, L.x.p);                         // post-smooth
}

// PCMG full: restrict b to every level, then for l = 0..L-2: V(l), x_{l+1} = P x_l ; V(L-1)
static void cycle_body(alfib_ctx* c) {
  const int nl = c->nlevels;
  for (int l = nl - 1; l > 0; --l) restrict_device(c, *c->levels[l], *c->levels[l - 1], l, c->levels[l]->b.p,
                                                   c->levels[l - 1]->b.p);
  CUDA_TRY(cudaMemsetAsync(c->levels[0]->x.p, 0, sizeof(double) * c->levels[0]->n, c->stream));
  for (int l = 0; l < nl - 1; ++l) {
    vcycle(c, l);
    prolong_device(c, *c->levels[l + 1], l + 1, c->levels[l]->x.p, c->levels[l + 1]->x.p);
  }
  vcycle(c, nl - 1);
}

void cycle_graph_invalidate(alfib_ctx* c) {
  if (c->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)c->graph_exec);
  c->graph_exec = nullptr;
  c->cycles_run = 0;
}

// The F-cycle is a fixed sequence of a few hundred launches without host synchronisation: after
// one eager application (which performs all lazy allocations) it is captured into a CUDA graph
// and replayed, which removes the launch overhead that dominates the small 2-D configurations.
// Profiling (per-kernel events) and ALFIB_OPT_CUDA_GRAPH = 0 keep the eager path.
void cycle_apply_device(alfib_ctx* c, const double* b, double* x) {
  const int nl = c->nlevels;
  ALFIB_REQUIRE(nl >= 1, "alfib_cycle_setup has not been called");
  Level& Lt = *c->levels[nl - 1];
  CUDA_TRY(cudaMemcpyAsync(Lt.b.p, b, sizeof(double) * Lt.n, cudaMemcpyDeviceToDevice, c->stream));
  const bool want_graph = c->use_graph && !c->profile;
  if (!want_graph) {
    cycle_body(c);
  } else if (c->graph_exec) {
    // the eager path checks these inside its steps; a replay must not run on stale inverses either
    for (int l = 1; l < nl; ++l)
      ALFIB_REQUIRE(c->levels[l]->ps[ALFIB_PATCHES_SMOOTHER].factored, "alfib_level_factor has not been called since the values changed");
    ALFIB_REQUIRE(c->coarse_factored, "alfib_coarse_factor has not been called since the values changed");
    CUDA_TRY(cudaGraphLaunch((cudaGraphExec_t)c->graph_exec, c->stream));
    c->launches += c->graph_launches;
  } else if (c->cycles_run == 0) {
    cycle_body(c);                                   // eager warm-up: allocations happen here
    c->cycles_run = 1;
  } else {
    const int64_t before = c->launches;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    try {
      cycle_body(c);
    } catch (...) {
      cudaStreamEndCapture(c->stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CUDA_TRY(cudaStreamEndCapture(c->stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {                          // fall back to eager launches for good
      cudaGetLastError();
      c->use_graph = 0;
      cycle_body(c);
    } else {
      c->graph_exec = exec;
      c->graph_launches = c->launches - before;
      CUDA_TRY(cudaGraphLaunch(exec, c->stream));
    }
  }
  CUDA_TRY(cudaMemcpyAsync(x, Lt.x.p, sizeof(double) * Lt.n, cudaMemcpyDeviceToDevice, c->stream));
}
