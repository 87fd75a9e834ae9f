/* alfib — B200-native velocity-block multigrid for alfi's augmented-Lagrangian preconditioner.
 *
 * C ABI of libalfib.so.  This is the drop-in boundary: everything above it (alfi's solver
 * dictionaries, Firedrake forms, mesh hierarchies, continuation drivers) stays Python and is
 * unchanged; everything below is hand-written CUDA for sm_100a.  Each entry point names the
 * reference interface it replaces (paths relative to the alfi repository).  The arithmetic
 * the reference delegates to PETSc/Firedrake (PCPATCH, MatMult_SeqBAIJ, KSPFGMRES, PCMG,
 * firedrake.mg prolong/restrict) is restated in SURVEY.md Appendix A.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ALFIB_E* code on failure; the message is
 *     available from alfib_last_error(ctx).  Nothing throws across the ABI.
 *   - vector / value pointers may be HOST or DEVICE pointers; the library detects which
 *     (cudaPointerGetAttributes).  Host buffers are borrowed for the duration of the call and
 *     copied; the library never keeps or frees caller memory.  Index arrays are host pointers.
 *   - one ctx per process <-> one GPU; calls are serialised by the caller (PETSc's single
 *     Python thread); work runs on the ctx's own non-default stream and calls return after
 *     host-visible results are complete (device-pointer calls return after enqueue unless
 *     ALFIB_SYNC_ALWAYS is set with alfib_set_option).
 *   - FP64 everywhere, int32 indices (PETSc int32 build), int64 patch offsets.
 *   - levels are numbered 0 (coarsest) .. L-1 (finest) like PCMG.
 */
#ifndef ALFIB_H
#define ALFIB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct alfib_ctx alfib_ctx;

enum {
  ALFIB_OK = 0,
  ALFIB_EINVAL = -1,   /* bad argument / call order */
  ALFIB_ECUDA = -2,    /* CUDA runtime, cuSOLVER or NCCL failure */
  ALFIB_ESINGULAR = -3,/* a patch (or the coarse) matrix is numerically singular */
  ALFIB_ENOMEM = -4
};

/* which set of patches of a level an operation addresses */
enum { ALFIB_PATCHES_SMOOTHER = 0, ALFIB_PATCHES_TRANSFER = 1 };

/* alfib_set_option keys */
enum {
  ALFIB_OPT_DETERMINISTIC = 1, /* 1: colour-ordered scatter-add + two-pass reductions (bitwise
                                  reproducible); 0: atomicAdd scatter in one launch            */
  ALFIB_OPT_SYNC_ALWAYS = 2,   /* 1: synchronise the stream before every return                */
  ALFIB_OPT_ROBUST_RESTRICT = 3,/* 1: Schoeberl restriction (alfi --restriction, solver.py:595,
                                  646); 0: plain P_H^T (firedrake restrict)                    */
  ALFIB_OPT_TRANSFER_REFINE = 4,/* 1 (default): one step of iterative refinement after the
                                  explicit-inverse cell-patch solve of the transfer, which
                                  restores the accuracy of the reference's LU solve            */
  ALFIB_OPT_CUDA_GRAPH = 5      /* 1 (default): alfib_cycle_apply replays a captured CUDA graph
                                  from its third call on (profiling keeps the eager path)      */
};

/* ---- context ------------------------------------------------------------------------------ */
int alfib_create(int device, alfib_ctx** out);
int alfib_destroy(alfib_ctx* ctx);
const char* alfib_last_error(const alfib_ctx* ctx);
int alfib_set_option(alfib_ctx* ctx, int key, int value);
/* convenience wrapper named in SURVEY §8b */
int alfib_set_deterministic(alfib_ctx* ctx, int flag);
int alfib_synchronize(alfib_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches claim) */
int64_t alfib_launch_count(const alfib_ctx* ctx);
/* CUDA stream handle (cudaStream_t) the ctx enqueues on, for event timing by the caller */
void* alfib_stream(alfib_ctx* ctx);

/* Optional: page-lock a HOST buffer the caller hands over repeatedly (the value array of the level operator, which
 * PETSc keeps at a fixed address from one Newton step to the next; right-hand-side / solution vectors), so that the
 * copies of alfib_level_set_bsr_values / alfib_cycle_apply run at PCIe speed instead of through the driver's staging
 * of pageable memory (1.8 GB of values per Newton step on the headline configuration).  The buffer must stay allocated
 * until alfib_host_unregister or alfib_destroy; registering twice, or a buffer that is already page-locked, is a no-op. */
int alfib_host_register(alfib_ctx* ctx, void* ptr, int64_t bytes);
int alfib_host_unregister(alfib_ctx* ctx, void* ptr);

/* Multi-GPU: one rank per GPU on one NVSwitch box.  (This paragraph is the replicated-vector mode of round 1, still
 * selectable; the default for N > 1 is the distributed mode of alfib_level_set_halo below.)  Each rank passes only ITS patches to
 * alfib_level_set_patches (the owned vertices of the DMPlex vertex-overlap partition the
 * reference uses, solver.py:604-605, 661-662, relaxation.py:120-121); level vectors are
 * replicated, block rows of the operators are split evenly.  The library then inserts the two
 * exchange steps of the path on its stream: ncclAllReduce(sum) after every patch apply (the
 * PetscSF reduce + bcast around PCApply_PATCH) and a grouped ncclBroadcast of the owned rows
 * after every SpMV (the VecScatter of MatMult_MPIBAIJ).  `nccl_unique_id` is the 128-byte
 * ncclUniqueId made by alfib_comm_unique_id on rank 0 and broadcast by the host (mpi4py in a
 * deployment, torch.distributed here).  Call before alfib_cycle_setup; without it the ctx is
 * serial.                                                                                     */
int alfib_comm_unique_id(void* out128 /* 128 bytes, filled on rank 0 */);
int alfib_comm_init(alfib_ctx* ctx, const void* nccl_unique_id, int rank, int nranks);
/* Optional, after all levels exist: exchange over NVLink *peer memory* instead of NCCL.  Every
 * rank exports a 64-byte CUDA IPC handle of its symmetric buffer (alfib_comm_peer_handle), the
 * host all-gathers the handles (nranks * 64 bytes, rank order) and passes them to
 * alfib_comm_peer_open.  The patch-apply sum, the SpMV row gather and the coarse GEMV gather then
 * become one kernel each that pulls only the overlapping index ranges from the peers' buffers
 * (csrc/comm.cu).  Collective: call on every rank.                                            */
int alfib_comm_peer_handle(alfib_ctx* ctx, void* out64);
int alfib_comm_peer_open(alfib_ctx* ctx, const void* handles);

/* Distributed level vectors (the PetscSF pattern of the reference's parallel PCPATCH / MatMult on the
 * vertex-overlap partition, solver.py:604-605, 661-662).  A level that is given a halo
 * holds LOCAL vectors: this rank's owned dofs first (n_owned_dofs of them), then its ghosts, n_local
 * dofs = the size alfib_level_create was called with; the operator pattern, patches, Dirichlet lists
 * and cell patches of such a level are in local numbering, block rows complete for the owned nodes
 * (alfi_b200.halo.local_level builds exactly this).  send_idx[send_off[p] .. send_off[p+1]) are the
 * owned local dofs peer `peers[p]` holds as ghosts, recv_idx the matching ghost local dofs on this
 * rank, both in the order the two ranks agree on.  The library then runs
 *   owner->ghost update  before every gather of ghost entries (patch gather, SpMV, P_H),
 *   ghost->owner sum     after every patch scatter-add (peers in ascending rank order: reproducible),
 * between neighbouring ranks only (grouped ncclSend/ncclRecv of the packed entries), reduces the
 * FGMRES dots over the owned entries with one small ncclAllReduce, and stops replicating vectors
 * and sharding operator rows on that level.  which = 0: the level's own vectors.  which = 1: the
 * "transfer halo" of level `level` — the layout in which P_H of level `level` reads the vectors of
 * level-1 (same owned dofs as that level's halo, ghosts = the other coarse dofs in the columns of
 * this rank's rows of P_H); without it level-1 must be replicated (no halo), its restricted
 * right-hand side is then summed with one all-reduce.  Call after alfib_comm_init and
 * alfib_level_create and before alfib_transfer_set; vectors passed to alfib_smooth / alfib_prolong /
 * alfib_restrict / alfib_cycle_apply on such a level are local vectors whose owned part is valid
 * on entry and on return.  peer_send_off[p] / peer_recv_off[p] (optional) = where, in peer p's
 * own send_idx / recv_idx lists, the segment for THIS rank starts; with them and
 * alfib_comm_peer_open the exchanges run over NVLink peer memory — pack into this rank's symmetric
 * slot, then one kernel that pulls the neighbours' packed entries into place — instead of NCCL,
 * and the small all-reduces of the dots take the same route.                                    */
int alfib_level_set_halo(alfib_ctx* ctx, int level, int which, int32_t n_owned_dofs, int32_t n_local_dofs,
                         int32_t npeers, const int32_t* peers, const int64_t* send_off,
                         const int32_t* send_idx, const int64_t* recv_off, const int32_t* recv_idx,
                         const int64_t* peer_send_off /* [npeers] or NULL */,
                         const int64_t* peer_recv_off /* [npeers] or NULL */);

/* ---- level operator: replaces the BAIJ Mat PETSc holds for fieldsplit_0 on each level
 *      (parameters["default_sub_matrix_type"] = "baij", solver.py:512)                        */
int alfib_level_create(alfib_ctx* ctx, int level, int n_nodes, int bs);
int alfib_level_set_bsr_pattern(alfib_ctx* ctx, int level, int64_t nnzb,
                                const int32_t* rowptr, const int32_t* colidx);
/* once per Newton step: the single hand-over of UFL/TSFC assembly output.
 * block_col_major = 1 for PETSc's BAIJ in-block layout, 0 for row-major blocks (scipy).        */
int alfib_level_set_bsr_values(alfib_ctx* ctx, int level, const double* vals, int block_col_major);
/* global Dirichlet dofs of the level (DirichletBC.nodes x components; ldc*.py bcs)             */
int alfib_level_set_bc(alfib_ctx* ctx, int level, int32_t nbc, const int32_t* bc_dofs);

/* MatMult / residual on the level operator (PETSc MatMult_SeqBAIJ)          */
int alfib_spmv(alfib_ctx* ctx, int level, const double* x, double* y);
int alfib_residual(alfib_ctx* ctx, int level, const double* b, const double* x, double* r);

/* ---- patches: replaces PCPATCH's per-patch index sets, operators and factorisations
 *      (pc_python_type firedrake.PatchPC, solver.py:318-344; transfer.py:100-113).
 *      `which` selects the smoother's patches or the transfer's cell patches.
 *      offsets[npatch+1] index into dofs (scalar dof numbers in patch-local order);
 *      order[norder] is the iteration set (relaxation.py:141-149); colours[npatch] may be NULL,
 *      then a greedy colouring in iteration order is computed.                    */
int alfib_level_set_patches(alfib_ctx* ctx, int level, int which, int32_t npatch,
                            const int64_t* offsets, const int32_t* dofs, int32_t norder,
                            const int32_t* order, const int32_t* colours);
/* Optional, after alfib_level_set_patches and before the first factorisation: the block/separator
 * structure of every patch (csrc/condense.cu).  block_of_dof has one entry per entry of `dofs`:
 * < 0 for separator dofs, otherwise a block label local to the patch.  Blocks must be pairwise
 * decoupled in the level's BSR pattern (checked: a wrong hint is ALFIB_EINVAL, never a wrong
 * result) and have <= 64 dofs and <= 64 coupled separator dofs.  The inverse of patch i is then
 * held as X_SS (the separator block of A_i^-1) plus D_k = A_kk^-1, A_Nk D_k and D_k A_kN per block:
 * for the macro-star patches of Scott-Vogelius elements on barycentrically refined meshes
 * (relaxation.py:168-177, bary.py) — interiors of the macro cells = blocks — that is ~10x fewer
 * bytes to store and to stream per PCApply_PATCH than the dense inverse PETSc's
 * `patch_pc_patch_dense_inverse` (solver.py:602) keeps; the result is the same A_i^-1 r_i.
 * Blocks with the same dof set in several patches (a macro cell is in the macro stars of all its
 * vertices) are stored and applied once (D_k, A_Uk D_k, D_k A_kU with U_k the union of their
 * separator neighbourhoods) when they are pairwise disjoint and |U_k| <= 64; environment
 * ALFIB_CONDENSE_SHARED=0 keeps one copy per (patch, block).
 * NULL returns the set to dense inverses.                                                      */
int alfib_level_set_patch_blocks(alfib_ctx* ctx, int level, int which, const int32_t* block_of_dof);
/* Multiplicative composition (`patch_pc_patch_local_type multiplicative`, with `symmetrise_sweep` the backward
 * sweep after the forward one; alfi/solver.py:306-308,322,324,331-335 — PCApply_PATCH then runs the patches of the
 * iteration set one after the other, y += R_i^T A_i^-1 R_i (x - A y)).  The sequential sweep is executed as a
 * schedule of STAGES: stage_of_visit[k] (0 <= . < nstage) for the k-th entry of the iteration set, such that two
 * visits whose patches are coupled through the operator (A[I_i, I_j] != 0, in particular patches sharing a dof) lie
 * in different stages, the earlier visit in the lower one (alfi_b200.patches.sweep_stages; checked here against the
 * BSR pattern's node adjacency).  Patches of one stage commute exactly, so the result is that of the sequential
 * sweep, bit for bit, whatever the stage is executed with.  Dense inverses only (no patch blocks); nstage = 0
 * returns the set to additive composition.                                                                  */
int alfib_level_set_sweep_stages(alfib_ctx* ctx, int level, int which, int32_t nvisit, const int32_t* stage_of_visit,
                                 int32_t nstage, int symmetric);
/* Patch operators that are NOT sub-matrices of the level operator.  With Burman's interior-facet
 * stabilisation (alfi/stabilisation.py:139-162, --stabilisation-type burman: what the reference's Scott-Vogelius jobs
 * run with, examples/Makefile:12-16) PCPATCH integrates a patch over its cells and over the facets whose BOTH cells are
 * patch cells, while A[I_i, I_i] also holds the inside-inside part of the facet integrals on the patch boundary.  The
 * host hands the difference over: A_i = A[I_i, I_i] + C_i, C_i in COO form with patch-local indices (rows / cols index
 * the patch's dof list), entries of patch i = [corr_off[i], corr_off[i+1]), sorted by (row, col) and distinct.  The
 * pattern is set once (after alfib_level_set_patches); the values follow every alfib_level_set_bsr_values, before
 * alfib_level_factor (which refuses to run on stale ones).  Dense inverses only: the jump terms couple the macro-cell
 * interiors across macro faces, so there is no block structure to condense.  corr_off = NULL removes them.        */
int alfib_level_set_patch_corrections(alfib_ctx* ctx, int level, int which, const int64_t* corr_off,
                                      const int32_t* rows, const int32_t* cols);
int alfib_level_set_patch_correction_values(alfib_ctx* ctx, int level, int which, const double* vals);
/* algorithmic bytes of one application of that patch set: stored factors + index data + 16 N   */
int64_t alfib_patch_apply_bytes(alfib_ctx* ctx, int level, int which);
/* bytes of device storage the inverse factors of that patch set need                          */
int64_t alfib_patch_storage_bytes(alfib_ctx* ctx, int level, int which);
/* how that patch set holds its inverses: 0 dense tiles, 1 condensed with one V / [D | -W] pair per
 * (patch, block), 2 condensed with the blocks shared between the patches that contain them
 * (alfib_level_set_patch_blocks; csrc/condense.cu); negative on a bad handle                      */
int alfib_patch_storage_form(alfib_ctx* ctx, int level, int which);
/* optional: caller-owned device buffer (a torch tensor's data_ptr()) for the factors; if never
 * called the library allocates.                                                               */
int alfib_patch_bind_storage(alfib_ctx* ctx, int level, int which, void* dev_ptr, int64_t bytes);
/* per Newton step (PCSetUp_PATCH; patch_pc_patch_save_operators + sub_pc_type lu +
 * dense_inverse, solver.py:320,327,602): A_i = A[I_i,I_i] gathered from the level's BSR values,
 * inverted in place by blocked Gauss-Jordan with partial (row) pivoting.                       */
int alfib_level_factor(alfib_ctx* ctx, int level);
/* PCApply_PATCH additive, no partition of unity (solver.py:321-322): y = sum_i R_i^T A_i^-1 R_i x,
 * then y[bc] = x[bc].                                                                          */
int alfib_smoother_apply(alfib_ctx* ctx, int level, const double* x, double* y);
/* copy the colours in use back (npatch int32) — for the bit-exactness test                     */
int alfib_get_colours(alfib_ctx* ctx, int level, int which, int32_t* colours);
/* debugging / tests: dense inverse of patch p in row-major n x n (host pointer)                */
int alfib_get_patch_inverse(alfib_ctx* ctx, int level, int which, int32_t patch, double* out);

/* ---- robust transfer: replaces AutoSchoeberlTransfer.prolong/restrict (transfer.py:186-275)
 *      between `level-1` and `level`.  P is the scalar CSR of the standard prolongation
 *      (firedrake prolong, transfer.py:284-290) acting per component — or, with dof_level = 1,
 *      a CSR on scalar dofs (n_fine_nodes / n_coarse_nodes are then dof counts), which is what
 *      BubbleTransfer's flux-corrected prolongation of [P1+FB]^3 needs because it mixes the
 *      components on every facet (bubble.py:10-265, transfer.py:334-356); cell patches come through
 *      alfib_level_set_patches(which = TRANSFER); cb_dofs are the coarse-boundary dofs of
 *      fix_coarse_boundaries (transfer.py:121-158).                                            */
int alfib_transfer_set(alfib_ctx* ctx, int level, int32_t n_fine_nodes, int32_t n_coarse_nodes,
                       const int32_t* P_rowptr, const int32_t* P_colidx, const double* P_vals,
                       int32_t ncb, const int32_t* cb_dofs, int dof_level);
/* once per (nu, gamma) (transfer.py:173-184, 238-244): BSR values, on the level's pattern, of
 * A0 = nu(2 sym grad u, grad v) + gamma(div u, div v) and of gamma*D = gamma(div u, div v)
 * (transfer.py:295-309 / 319-332); gathers + inverts the cell patches.  A0_vals may be NULL to
 * keep the previous factors (only D changes), both NULL is an error.                           */
int alfib_transfer_update(alfib_ctx* ctx, int level, const double* A0_vals, const double* D_vals,
                          int block_col_major);
int alfib_prolong(alfib_ctx* ctx, int level, const double* coarse, double* fine);
int alfib_restrict(alfib_ctx* ctx, int level, const double* fine, double* coarse);

/* ---- level smoother and cycle: replaces the mg_levels KSP (fgmres, max_it = smoothing,
 *      convergence_test skip; solver.py:313-317) and fieldsplit_0 = richardson(1) + PCMG full
 *      with a direct coarse solve (solver.py:359-379).                                         */
int alfib_smooth(alfib_ctx* ctx, int level, int m, const double* b, double* x);
/* dense LU of the level-0 operator (replaces AssembledPC + telescope + superlu_dist)           */
int alfib_coarse_factor(alfib_ctx* ctx);
int alfib_coarse_solve(alfib_ctx* ctx, const double* b, double* x);
int alfib_cycle_setup(alfib_ctx* ctx, int nlevels, int smoothing);
/* one application of fieldsplit_0: x = Fcycle(b) on the finest level                           */
int alfib_cycle_apply(alfib_ctx* ctx, const double* b, double* x);

/* ---- outer Schur-complement fieldsplit: the pieces that bracket fieldsplit_0 in alfi's outer
 *      solver, so that one outer Krylov iteration costs one PCIe round trip instead of two velocity-block applications
 *      with their own.  Replaces PCFIELDSPLIT schur / full (solver.py:405-421) with fieldsplit_1 = alfi.solver.DGMassInv
 *      (solver.py:15-38): y1 = A^-1 r_u ; y_p = -(nu + gamma) M_p^-1 (r_p - B y1), constants removed if remove_constant
 *      (the pressure nullspace, problem.py:33-38) ; y_u = A^-1 (r_u - B^T y_p), A^-1 = alfib_cycle_apply.
 *      B (n_p x n_u: the assembled (div u, q) block of the Jacobian, Dirichlet columns removed as Firedrake's bcs do)
 *      and M_p^-1 (n_p x n_p, block diagonal for the discontinuous pressure spaces) are scalar CSR matrices on the
 *      dofs of the FINEST level; vectors are [velocity dofs ; pressure dofs].  One GPU; after alfib_cycle_setup.   */
int alfib_schur_set(alfib_ctx* ctx, int32_t n_p, const int32_t* B_rowptr, const int32_t* B_colidx, const double* B_vals,
                    const int32_t* Minv_rowptr, const int32_t* Minv_colidx, const double* Minv_vals,
                    int remove_constant);
int alfib_schur_apply(alfib_ctx* ctx, double nu, double gamma, const double* r, double* y);
/* MatMult of the saddle-point Jacobian [A B^T; B 0] (A = the finest level's BSR values)                          */
int alfib_jacobian_apply(alfib_ctx* ctx, const double* z, double* Jz);
/* The outer KSP of a Newton step (solver.py:463-474: fgmres, right preconditioning, classical Gram-Schmidt,
 * restart <= 32, zero initial guess; converged when the recurrence residual <= max(rtol |rhs|, atol)).
 * iterations: Krylov iterations taken; history[0 .. min(nhistory, iterations + 1)): residual norms (may be NULL). */
int alfib_outer_solve(alfib_ctx* ctx, double nu, double gamma, const double* rhs, double* x, double rtol, double atol,
                      int32_t maxit, int32_t restart, int32_t* iterations, double* history, int32_t nhistory);

/* ---- instrumentation: names follow the PETSc events alfi reports (driver.py:80)              */
enum {
  ALFIB_EV_PCPATCH_APPLY = 0, ALFIB_EV_MATMULT = 1, ALFIB_EV_PROLONG = 2, ALFIB_EV_RESTRICT = 3,
  ALFIB_EV_KSP_GMRES_ORTHOG = 4, ALFIB_EV_COARSE = 5, ALFIB_EV_PCSETUP_PATCH = 6,
  ALFIB_EV_HALO = 7 /* PetscSF bcast / reduce of ghost entries (distributed vectors) */, ALFIB_EV_COUNT = 8
};
/* enable/disable CUDA-event timing per kernel family and level.  Events are recorded on the ctx
 * stream without synchronising and resolved when read, so the timed region is not perturbed.  */
int alfib_profile(alfib_ctx* ctx, int enable);
/* accumulated milliseconds and call counts per event since the last reset; level = -1 sums
 * over all levels                                                                             */
int alfib_profile_get(alfib_ctx* ctx, int level, double* ms /*[ALFIB_EV_COUNT]*/,
                      int64_t* calls /*[ALFIB_EV_COUNT]*/);
int alfib_profile_reset(alfib_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ALFIB_H */
