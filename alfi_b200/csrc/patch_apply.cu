// Additive-Schwarz patch apply:  y += sum_i R_i^T A_i^{-1} R_i x   (PCApply_PATCH, additive, no
// partition of unity — alfi/solver.py:318-324).
//
// The explicit inverses (patch_pc_patch_dense_inverse, alfi/solver.py:602) are stored in a
// layout made for streaming: patch i with n rows is cut into tiles of 64 rows; a tile is
// column-major with `rows_t` (even) rows per column, so the double2 of lane l in column c holds
// rows (2l, 2l+1) and a warp reads one contiguous 16*rows_t/2-byte run per column and one
// contiguous 64*n*8-byte region overall.  One warp = one (patch, tile) work item; the gathered
// right-hand side x[I_i] is fetched 32 entries at a time by the lanes and broadcast with
// shuffles, so no shared memory or block barrier is needed and warps of one CTA may work on
// different patches.  Roofline: HBM; algorithmic bytes sum_i (8 n_i^2 + 4 n_i) + 16 N.
//
// Scatter-add is race-free by colouring in deterministic mode (one launch per colour, plain
// stores: patches of one colour share no dof, tiles of one patch own disjoint rows) or uses
// fp64 atomicAdd in a single launch otherwise.
#include "alfib_internal.h"

namespace {

template <bool ATOMIC>
__global__ void __launch_bounds__(128) patch_apply_kernel(const int2* __restrict__ work, int nwork,
                                                          const int64_t* __restrict__ poff,
                                                          const int32_t* __restrict__ pdofs,
                                                          const int64_t* __restrict__ soff,
                                                          const double* __restrict__ store,
                                                          const double* __restrict__ x, PeerOut yout) {
  double* __restrict__ y = resolve(yout);
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nwork) return;
  const int2 wk = work[w];
  const int64_t o = poff[wk.x];
  const int n = (int)(poff[wk.x + 1] - o);
  const int32_t* __restrict__ I = pdofs + o;
  const int row0 = wk.y * ALFIB_TILE_ROWS;
  int rows = n - row0;
  rows = rows > ALFIB_TILE_ROWS ? ALFIB_TILE_ROWS : ((rows + 1) & ~1);
  const int half = rows >> 1;                       // double2 per column of this tile
  const bool active = lane < half;
  const double2* __restrict__ T =
      reinterpret_cast<const double2*>(store + soff[wk.x] + (int64_t)row0 * n) + (active ? lane : 0);

  double acc0 = 0.0, acc1 = 0.0;
  // software-pipelined gather of x[I[c]] for the next 32 columns
  double xv = (lane < n) ? __ldg(x + I[lane]) : 0.0;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int cn = c0 + 32 + lane;
    const double xnext = (cn < n) ? __ldg(x + I[cn]) : 0.0;
    const int cnt = (n - c0) < 32 ? (n - c0) : 32;
    const double2* __restrict__ Tc = T + (int64_t)c0 * half;
    if (cnt == 32) {
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const double xc = __shfl_sync(0xffffffffu, xv, j);
        if (active) {
          const double2 a = __ldcs(Tc + (int64_t)j * half);
          acc0 = fma(a.x, xc, acc0);
          acc1 = fma(a.y, xc, acc1);
        }
      }
    } else {
      for (int j = 0; j < cnt; ++j) {
        const double xc = __shfl_sync(0xffffffffu, xv, j);
        if (active) {
          const double2 a = __ldcs(Tc + (int64_t)j * half);
          acc0 = fma(a.x, xc, acc0);
          acc1 = fma(a.y, xc, acc1);
        }
      }
    }
    xv = xnext;
  }
  if (active) {
    const int r = row0 + 2 * lane;
    if (ATOMIC) {
      atomicAdd(y + I[r], acc0);
      if (r + 1 < n) atomicAdd(y + I[r + 1], acc1);
    } else {
      y[I[r]] += acc0;
      if (r + 1 < n) y[I[r + 1]] += acc1;
    }
  }
}

}  // namespace

void launch_patch_apply(alfib_ctx* c, const PatchSet& ps, const double* x, PeerOut y) {
  if (ps.nwork == 0) return;
  if (ps.cond.on) {                   // block/separator form of the inverses (condense.cu)
    launch_condensed_apply(c, ps, x, y);
    return;
  }
  const int threads = 128, wpb = threads / 32;
  const bool coloured = c->deterministic && !ps.repeated;
  if (!coloured && ps.ncolour > 1) {
    patch_apply_kernel<true><<<cdiv(ps.nwork, wpb), threads, 0, c->stream>>>(
        ps.work.p, ps.nwork, ps.off.p, ps.dofs.p, ps.soff.p, ps.store, x, y);
    c->launches++;
  } else {
    // one launch per colour; a single colour (disjoint cell patches) never needs atomics
    for (int col = 0; col < ps.ncolour; ++col) {
      const int s = ps.colour_work_start[col], e = ps.colour_work_start[col + 1];
      if (e == s) continue;
      if (ps.repeated)
        patch_apply_kernel<true><<<cdiv(e - s, wpb), threads, 0, c->stream>>>(
            ps.work.p + s, e - s, ps.off.p, ps.dofs.p, ps.soff.p, ps.store, x, y);
      else
        patch_apply_kernel<false><<<cdiv(e - s, wpb), threads, 0, c->stream>>>(
            ps.work.p + s, e - s, ps.off.p, ps.dofs.p, ps.soff.p, ps.store, x, y);
      c->launches++;
    }
  }
  CUDA_TRY(cudaGetLastError());
}

// Multiplicative composition: the sequential sweep of PCApply_PATCH as a schedule of stages (include/alfib.h,
// alfib_level_set_sweep_stages).  Per stage: the residual x - A y on the block rows the stage's patches read (a
// component-wise sum: lanes over the row's blocks, fixed shuffle tree), then the dense-inverse apply of the stage's
// (patch, tile) items with plain stores — patches of a stage share no dof.  Backward sweep = stages in reverse.
void patch_apply_multiplicative(alfib_ctx* c, Level& L, int which, const double* x, double* y) {
  PatchSet& ps = L.ps[which];
  ALFIB_REQUIRE(ps.nstage > 0, "no sweep stages on this patch set");
  ALFIB_REQUIRE(!ps.cond.on, "multiplicative composition needs dense patch inverses");
  ALFIB_REQUIRE(c->nranks == 1 && !L.halo.on, "multiplicative composition is single-GPU");
  L.t3.alloc(L.n);
  double* r = L.t3.p;
  CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * L.n, c->stream));
  const int threads = 128, wpb = threads / 32;
  auto stage = [&](int s) {
    const int r0 = ps.stage_row_start[s], r1 = ps.stage_row_start[s + 1];
    launch_bsr_residual_rows(c, L, L.vals.p, ps.stage_rows.p + r0, r1 - r0, x, y, r);
    const int w0 = ps.stage_work_start[s], w1 = ps.stage_work_start[s + 1];
    if (w1 > w0) {
      patch_apply_kernel<false><<<cdiv(w1 - w0, wpb), threads, 0, c->stream>>>(ps.stage_work.p + w0, w1 - w0, ps.off.p, ps.dofs.p,
                                                                               ps.soff.p, ps.store, r, plain_out(y));
      c->launches++;
    }
  };
  for (int s = 0; s < ps.nstage; ++s) stage(s);
  if (ps.symmetric_sweep)
    for (int s = ps.nstage - 1; s >= 0; --s) stage(s);
  CUDA_TRY(cudaGetLastError());
}

// Sum of the patch contributions over all ranks.  Serial: zero + apply.  Multi-GPU with peer
// memory: every rank applies its patches into its symmetric slot (only the slab the patches
// touch is zeroed) and then *pulls* the overlapping ranges of the other ranks' slots over NVLink
// inside one reduction kernel (comm.cu) — the ghost->owner sum and owner->ghost broadcast of the
// reference's PetscSF in one pass.  Without peer memory: NCCL all-reduce of the whole vector.
void patch_apply_sum(alfib_ctx* c, Level& L, int level, int which, const double* x, double* y) {
  const PatchSet& ps = L.ps[which];
  if (ps.nstage > 0) {
    patch_apply_multiplicative(c, L, which, x, y);
    return;
  }
  if (L.halo.on) {
    // distributed vectors: the PetscSF bcast of x, this rank's patches, the PetscSF reduce of y (Appendix A.3)
    halo_update(c, L.halo, const_cast<double*>(x), level);
    CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * L.n, c->stream));
    launch_patch_apply(c, ps, x, plain_out(y));
    halo_reduce(c, L.halo, y, level);
    return;
  }
  if (c->nranks > 1 && c->peers_open) {
    comm_peer_zero(c, ps.lo, ps.hi);
    launch_patch_apply(c, ps, x, comm_peer_out(c));
    comm_peer_reduce(c, L.n, level * 2 + which, nullptr, nullptr, y);
    return;
  }
  CUDA_TRY(cudaMemsetAsync(y, 0, sizeof(double) * L.n, c->stream));
  launch_patch_apply(c, ps, x, plain_out(y));
  comm_allreduce_sum(c, y, L.n);
}
