"""Host-side driver: loads per-level hand-over data into the CUDA library and runs the cycle.

This is what the petsc4py plugins of :mod:`alfi_b200.pc` wrap.  The input is whatever produces
the per-level data — Firedrake/alfi in a deployment, :mod:`alfi_b200.synth` here.  Data handed
over per level (`LevelInput`): BSR pattern + values of the velocity operator, Dirichlet dofs,
smoother patch dof sets, and (levels >= 1) the standard prolongation, the cell patches, the
coarse-boundary dofs and the A0 / gamma*D values of the Schöberl transfer.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .lib import PATCHES_SMOOTHER, PATCHES_TRANSFER, Context

__all__ = ["LevelInput", "DeviceMultigrid", "DistributedMultigrid", "level_input_from_synth"]


@dataclass
class LevelInput:
    n_nodes: int
    bs: int
    rowptr: np.ndarray
    colidx: np.ndarray
    vals: np.ndarray                      # (nnzb, bs, bs) row-major blocks
    bc_dofs: np.ndarray
    patch_offsets: np.ndarray | None = None
    patch_dofs: np.ndarray | None = None
    patch_order: np.ndarray | None = None
    patch_colours: np.ndarray | None = None
    patch_blocks: np.ndarray | None = None   # condensed form: block label per patch dof, -1 = separator
    patch_stages: np.ndarray | None = None   # multiplicative composition: stage per entry of the iteration set
    symmetrise_sweep: bool = False           # ... with the backward sweep after the forward one
    # patch operators that are not sub-matrices (Burman's interior-facet term): A_i = A[I_i, I_i] + C_i,
    # C_i as COO entries in patch-local indices (pattern once, values with every operator hand-over)
    patch_corr_off: np.ndarray | None = None
    patch_corr_rows: np.ndarray | None = None
    patch_corr_cols: np.ndarray | None = None
    patch_corr_vals: np.ndarray | None = None
    P: object | None = None               # scipy CSR: scalar per node, or on dofs if P_dof_level
    P_dof_level: bool = False
    cell_offsets: np.ndarray | None = None
    cell_dofs: np.ndarray | None = None
    cell_blocks: np.ndarray | None = None
    cb_dofs: np.ndarray | None = None
    a0_vals: np.ndarray | None = None
    d_vals: np.ndarray | None = None
    # coarsest level only: its free dofs as ONE patch with macro-cell block labels -> condensed coarse inverse
    coarse_dofs: np.ndarray | None = None
    coarse_blocks: np.ndarray | None = None


def level_input_from_synth(ld) -> LevelInput:
    """alfi_b200.synth.problem.LevelData → LevelInput."""
    li = LevelInput(ld.V.nnodes, ld.V.bs, ld.A.rowptr, ld.A.colidx, ld.A.vals, ld.bc_dofs)
    if ld.patches is not None:
        ps = ld.patches
        li.patch_offsets, li.patch_dofs, li.patch_order, li.patch_colours = ps.offsets, ps.dofs, ps.order, ps.colours
        li.patch_blocks = ps.blocks
        li.patch_stages = getattr(ps, "stages", None)
        li.symmetrise_sweep = bool(getattr(ps, "symmetrise", False))
        if getattr(ps, "corrections", None) is not None:
            pc = ps.corrections
            li.patch_corr_off, li.patch_corr_rows, li.patch_corr_cols, li.patch_corr_vals = pc.off, pc.rows, pc.cols, ps.corr_vals
    if ld.patches is None and getattr(ld.level, "bary", False) and not hasattr(ld.pattern, "facet_cells"):
        from .patches import PatchSet, macro_interior_blocks
        free = np.setdiff1d(np.arange(ld.V.ndofs), ld.bc_dofs).astype(np.int32)
        one = PatchSet(offsets=np.array([0, free.size], np.int64), dofs=free, bs=ld.V.bs, order=np.zeros(1, np.int32))
        blocks = macro_interior_blocks(ld.level.plex, ld.V, one)
        if blocks is not None and (blocks >= 0).any():
            li.coarse_dofs, li.coarse_blocks = free, blocks
    if ld.P is not None:
        li.P = ld.P
        li.P_dof_level = bool(getattr(ld, "P_dof_level", False))
        if ld.cell_patches is not None:
            li.cell_offsets, li.cell_dofs = ld.cell_patches.offsets, ld.cell_patches.dofs
            li.cell_blocks = ld.cell_patches.blocks
            li.cb_dofs = ld.cb_dofs
            li.a0_vals, li.d_vals = ld.A0.vals, ld.D.vals
    return li


class DeviceMultigrid:
    """The velocity-block multigrid of alfi/solver.py:359-379 resident on one GPU."""

    def __init__(self, levels: list[LevelInput], smoothing: int, device: int = 0, deterministic: bool = False,
                 robust_restrict: bool = True, ctx: Context | None = None, torch_storage: bool = False,
                 rank: int = 0, nranks: int = 1, unique_id: bytes | None = None,
                 peer_memory: bool = False, condense: bool = True):
        """With nranks > 1 every rank passes the same global `levels`; this rank keeps the patches
        `alfi_b200.dist.partition_patches` assigns to it (in a deployment each rank would only
        ever see its own) and the library adds the exchange steps: NCCL collectives by default,
        NVLink peer-memory pull-reductions with ``peer_memory=True`` (measured equal within 2 % at
        2/4/8 GPUs in round 1, so the simpler NCCL path is the default).

        ``condense``: use the block/separator form of the patch inverses (csrc/condense.cu) for the
        patch sets whose LevelInput carries block labels; False keeps dense inverses everywhere."""
        from .dist import condensed_cost, partition_patches, shard_dof_array, shard_patch_arrays
        self.ctx = ctx or Context(device, deterministic)
        self.nlevels = len(levels)
        self.smoothing = smoothing
        self.sizes = [li.n_nodes * li.bs for li in levels]
        self.rank, self.nranks = rank, nranks
        self._storage = []
        self.local_patches = {}
        c = self.ctx
        c.set_option(3, robust_restrict)
        if nranks > 1:
            c.comm_init(unique_id, rank, nranks)
        for l, li in enumerate(levels):
            c.level_create(l, li.n_nodes, li.bs)
            c.set_bsr_pattern(l, li.rowptr, li.colidx)
            c.set_bc(l, li.bc_dofs)
            if l == 0 and condense and li.coarse_blocks is not None and os.environ.get("ALFIB_COARSE_CONDENSED", "1") != "0":
                # the coarse level as one patch with macro-cell blocks: coarse_factor keeps the condensed pieces of
                # the dense inverse (X_SS of the separator + block tiles) instead of the inverse itself
                c.set_patches(0, np.array([0, li.coarse_dofs.size], np.int64), li.coarse_dofs, np.zeros(1, np.int32),
                              np.zeros(1, np.int32), PATCHES_SMOOTHER)
                c.set_patch_blocks(0, li.coarse_blocks, PATCHES_SMOOTHER)
            if l > 0:
                off, dofs, order, cols = li.patch_offsets, li.patch_dofs, li.patch_order, li.patch_colours
                blocks = li.patch_blocks if (condense and li.patch_stages is None) else None   # sweeps: dense inverses
                if li.patch_stages is not None and nranks > 1:
                    raise NotImplementedError("multiplicative patch composition is single-GPU")
                if nranks > 1:
                    # balance what is streamed per application: condensed bytes where blocks are given
                    cost = condensed_cost(off, blocks) if blocks is not None else None
                    owner = partition_patches(off, dofs, nranks, cost)
                    goff = off
                    off, dofs, order, cols, mine = shard_patch_arrays(off, dofs, order, cols, owner, rank)
                    blocks = shard_dof_array(goff, blocks, mine)
                    self.local_patches[(l, PATCHES_SMOOTHER)] = mine
                c.set_patches(l, off, dofs, order, cols, PATCHES_SMOOTHER)
                if blocks is not None:
                    c.set_patch_blocks(l, blocks, PATCHES_SMOOTHER)
                if li.patch_corr_off is not None:
                    if nranks > 1:
                        raise NotImplementedError("patch corrections (Burman stabilisation) are single-GPU")
                    c.set_patch_corrections(l, li.patch_corr_off, li.patch_corr_rows, li.patch_corr_cols, PATCHES_SMOOTHER)
                if li.patch_stages is not None:
                    c.set_sweep_stages(l, li.patch_stages, li.symmetrise_sweep, PATCHES_SMOOTHER)
                if torch_storage:
                    self._bind(l, PATCHES_SMOOTHER)
                cb = li.cb_dofs if li.cb_dofs is not None else np.empty(0, np.int32)
                c.set_transfer(l, li.P, cb, li.P_dof_level)
                if li.cell_offsets is not None:
                    off, dofs = li.cell_offsets, li.cell_dofs
                    cols = np.zeros(off.size - 1, np.int32)
                    order = None
                    blocks = li.cell_blocks if condense else None
                    if nranks > 1:
                        owner = partition_patches(off, dofs, nranks)
                        goff = off
                        off, dofs, order, cols, mine = shard_patch_arrays(off, dofs, None, cols, owner, rank)
                        blocks = shard_dof_array(goff, blocks, mine)
                        self.local_patches[(l, PATCHES_TRANSFER)] = mine
                    c.set_patches(l, off, dofs, order, cols, PATCHES_TRANSFER)
                    if blocks is not None:
                        c.set_patch_blocks(l, blocks, PATCHES_TRANSFER)
                    if torch_storage:
                        self._bind(l, PATCHES_TRANSFER)
        if nranks > 1 and peer_memory:
            # map every rank's symmetric buffer: exchanges become NVLink peer loads (csrc/comm.cu)
            import torch.distributed as dist
            handles = [None] * nranks
            dist.all_gather_object(handles, c.comm_peer_handle())
            c.comm_peer_open(b"".join(handles))
        self.update_operators(levels)
        self.update_transfers(levels)
        c.cycle_setup(self.nlevels, smoothing)

    def _bind(self, level, which):
        """PyTorch owns the big factor buffers (north star: torch for buffer ownership only)."""
        import torch
        nbytes = self.ctx.patch_storage_bytes(level, which)
        buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda:%d" % self.ctx.device)
        self._storage.append(buf)
        self.ctx.bind_patch_storage(level, buf, which)

    def update_operators(self, levels, pin_values=False):
        """Once per Newton step: new BSR values on every level, patch factors, coarse LU.  pin_values: page-lock the
        value arrays on first sight (they must then be the SAME arrays, refilled, on later calls — as PETSc's are)."""
        c = self.ctx
        for l, li in enumerate(levels):
            if pin_values and isinstance(li.vals, np.ndarray) and li.vals.flags.c_contiguous and li.vals.nbytes >= (1 << 20):
                seen = self.__dict__.setdefault("_pinned", set())
                if li.vals.ctypes.data not in seen:
                    c.host_register(li.vals)
                    seen.add(li.vals.ctypes.data)
            c.set_bsr_values(l, li.vals)
            if l > 0:
                if li.patch_corr_off is not None:
                    c.set_patch_correction_values(l, li.patch_corr_vals, PATCHES_SMOOTHER)
                c.factor(l)
        c.coarse_factor()

    def update_transfers(self, levels):
        """Once per (nu, gamma): transfer.py:238-244."""
        for l, li in enumerate(levels):
            if l > 0 and li.a0_vals is not None:
                self.ctx.transfer_update(l, li.a0_vals, li.d_vals)

    def apply(self, b, x):
        """x = one fieldsplit_0 application (F-cycle) of b."""
        return self.ctx.cycle_apply(b, x)


class DistributedMultigrid:
    """The same cycle on `nranks` GPUs with DISTRIBUTED level vectors: every level >= 1
    is handed to the library as this rank's `alfi_b200.halo.LocalLevel` — owned dofs first, then ghosts, operator
    rows / patches / cell patches / P_H rows in local numbering — with the exchange lists of its halo
    (`alfib_level_set_halo`); level 0 stays replicated (redundant solve after one small all-reduce).  Per Krylov
    iteration the library then moves only ghost entries between neighbouring ranks (two owner->ghost updates, one
    ghost->owner sum) and reduces the dots with one small all-reduce each, instead of the full-vector all-reduce /
    all-gather of the replicated design (`DeviceMultigrid` with nranks > 1).

    Every rank passes the same global `levels` here (a deployment would build the LocalLevels from its own mesh
    partition).  Vectors given to `apply` are LOCAL (`scatter` / `local_dofs`); `gather` assembles the global
    result with torch.distributed.  oracle/distributed.py is the CPU statement of the same algorithm."""

    def __init__(self, levels: list[LevelInput], smoothing: int, rank: int, nranks: int, unique_id: bytes | None,
                 device: int = 0, deterministic: bool = False, robust_restrict: bool = True, ctx=None,
                 torch_storage: bool = False, condense: bool = True, peer_memory: bool = False):
        """``peer_memory``: ghost exchanges and the dots' all-reduces over NVLink peer memory (one pull kernel per
        exchange, csrc/comm.cu) instead of NCCL send/recv; collective (torch.distributed all-gather of the IPC
        handles)."""
        import scipy.sparse as sp

        from .dist import condensed_cost, partition_patches
        from .halo import build_layout, local_level, transfer_halo
        self.ctx = c = ctx or Context(device, deterministic)
        self.nlevels, self.smoothing = len(levels), smoothing
        self.rank, self.nranks = rank, nranks
        self._storage = []
        c.set_option(3, robust_restrict)
        c.comm_init(unique_id, rank, nranks)
        # ---- layouts (identical on every rank: pure functions of the global data)
        self.layouts, self.halos = [None], [None]
        for l in range(1, self.nlevels):
            li = levels[l]
            blocks = li.patch_blocks if condense else None
            cost = condensed_cost(li.patch_offsets, blocks) if blocks is not None else None
            owner = partition_patches(li.patch_offsets, li.patch_dofs, nranks, cost)
            extra = (li.cell_offsets, li.cell_dofs) if li.cell_offsets is not None else None
            self.layouts.append(build_layout(li.patch_offsets, li.patch_dofs, li.patch_order, owner, li.rowptr, li.colidx,
                                             li.bs, li.n_nodes * li.bs, extra_sets=extra, nranks=nranks))
        for l in range(1, self.nlevels):
            li = levels[l]
            if l == 1:
                self.halos.append(None)
            else:
                P = li.P.tocsr() if li.P_dof_level else sp.kron(li.P, sp.identity(li.bs), format="csr")
                self.halos.append(transfer_halo(P, self.layouts[l], self.layouts[l - 1].owner))
        self.local = [None] + [local_level(levels[l], self.layouts[l], rank, self.halos[l]) for l in range(1, self.nlevels)]
        thalo = [None] * self.nlevels
        for l in range(2, self.nlevels):
            hr = self.halos[l].ranks[rank]
            thalo[l] = (hr.n_owned, hr.n_local, hr.send, {q: hr.n_owned + v for q, v in hr.recv.items()})
        peer_off = {}
        if peer_memory:
            for l in range(1, self.nlevels):
                peer_off[(l, 0)] = self._peer_offsets(self.layouts[l])
                if self.halos[l] is not None:
                    peer_off[(l, 1)] = self._peer_offsets(self.halos[l])
        self._levels0 = levels[0]
        self._handover(levels[0], thalo, peer_off, condense, torch_storage, peer_memory)
        self.update_operators(levels)
        self.update_transfers(levels)
        c.cycle_setup(self.nlevels, smoothing)

    @classmethod
    def from_local(cls, problem, smoothing: int, unique_id: bytes | None, device: int = 0, deterministic: bool = False,
                   robust_restrict: bool = True, ctx=None, torch_storage: bool = False, condense: bool = True,
                   peer_memory: bool = False):
        """From a rank-locally generated problem (`alfi_b200.synth.bricks.build_rank_local`): nothing global besides
        the replicated level 0 exists on this rank.  The transfer halo of level l is the halo of level l-1 (P_H reads
        that level in its own local set).  With ``peer_memory`` the neighbours' list offsets are all-gathered."""
        self = cls.__new__(cls)
        self.ctx = c = ctx or Context(device, deterministic)
        self.nlevels, self.smoothing = len(problem.local), smoothing
        self.rank, self.nranks = problem.rank, problem.nranks
        self._storage = []
        self.layouts = self.halos = None
        self.local = problem.local
        c.set_option(3, robust_restrict)
        c.comm_init(unique_id, self.rank, self.nranks)
        thalo = [None] * self.nlevels
        for l in range(2, self.nlevels):
            lc = self.local[l - 1]
            thalo[l] = (lc.n_owned, lc.n_local, lc.send, lc.recv)
        peer_off = {}
        if peer_memory and self.nranks > 1:
            import torch.distributed as dist
            mine = {l: Context.halo_peer_list(self.local[l].send, self.local[l].recv) for l in range(1, self.nlevels)}
            everyone = [None] * self.nranks
            dist.all_gather_object(everyone, {l: (v[0], v[1].tolist(), v[2].tolist()) for l, v in mine.items()})
            for l in range(1, self.nlevels):
                off = {}
                for q in mine[l][0]:
                    peers_q, s_off, r_off = everyone[q][l]
                    k = peers_q.index(self.rank)
                    off[q] = (int(s_off[k]), int(r_off[k]))
                peer_off[(l, 0)] = off
                if l + 1 < self.nlevels:
                    peer_off[(l + 1, 1)] = off
        self._levels0 = problem.level0
        self._handover(problem.level0, thalo, peer_off, condense, torch_storage, peer_memory)
        self.update_operators(None)
        self.update_transfers(None)
        c.cycle_setup(self.nlevels, smoothing)
        return self

    def _handover(self, li, thalo, peer_off, condense, torch_storage, peer_memory):
        c = self.ctx
        c.level_create(0, li.n_nodes, li.bs)
        c.set_bsr_pattern(0, li.rowptr, li.colidx)
        c.set_bc(0, li.bc_dofs)
        if condense and li.coarse_blocks is not None and os.environ.get("ALFIB_COARSE_CONDENSED", "1") != "0":
            c.set_patches(0, np.array([0, li.coarse_dofs.size], np.int64), li.coarse_dofs, np.zeros(1, np.int32),
                          np.zeros(1, np.int32), PATCHES_SMOOTHER)
            c.set_patch_blocks(0, li.coarse_blocks, PATCHES_SMOOTHER)
        for l in range(1, self.nlevels):
            ll = self.local[l]
            c.level_create(l, ll.n_local_nodes, ll.bs)
            c.set_halo(l, ll.n_owned, ll.n_local, ll.send, ll.recv, 0, peer_off.get((l, 0)))
            if thalo[l] is not None:
                c.set_halo(l, thalo[l][0], thalo[l][1], thalo[l][2], thalo[l][3], 1, peer_off.get((l, 1)))
            c.set_bsr_pattern(l, ll.rowptr, ll.colidx)
            c.set_bc(l, ll.bc_dofs)
            c.set_patches(l, ll.patch_offsets, ll.patch_dofs, ll.patch_order, ll.patch_colours, PATCHES_SMOOTHER)
            if condense and ll.patch_blocks is not None:
                c.set_patch_blocks(l, ll.patch_blocks, PATCHES_SMOOTHER)
            if torch_storage:
                self._bind(l, PATCHES_SMOOTHER)
            c.set_transfer(l, ll.P, ll.cb_dofs if ll.cb_dofs is not None else np.empty(0, np.int32), True)
            if ll.cell_offsets is not None:
                c.set_patches(l, ll.cell_offsets, ll.cell_dofs, None, np.zeros(ll.cell_offsets.size - 1, np.int32),
                              PATCHES_TRANSFER)
                if condense and ll.cell_blocks is not None:
                    c.set_patch_blocks(l, ll.cell_blocks, PATCHES_TRANSFER)
                if torch_storage:
                    self._bind(l, PATCHES_TRANSFER)
        if peer_memory and self.nranks > 1:
            import torch.distributed as dist
            handles = [None] * self.nranks
            dist.all_gather_object(handles, c.comm_peer_handle())
            c.comm_peer_open(b"".join(handles))

    _bind = DeviceMultigrid._bind

    def _host_barrier(self):
        """Ranks leave the rank-local setup phases (seconds of factorisation) at different times; the device-side
        exchanges spin on their neighbours with a time-out, so the hosts meet before the next exchange is enqueued."""
        if self.nranks > 1:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    self.ctx.synchronize()
                    dist.barrier()
            except ImportError:
                pass

    def _peer_offsets(self, layout):
        """{peer: (start of this rank's segment in the peer's packed send list, in its packed ghost list)} — in a
        deployment two integers per neighbour exchanged at setup; here read off the global layout."""
        me = layout.ranks[self.rank]
        out = {}
        for q in sorted(set(me.send) | set(me.recv)):
            rq = layout.ranks[q]
            peers_q, s_off, r_off = Context.halo_peer_list(rq.send, rq.recv)
            k = peers_q.index(self.rank)
            out[q] = (int(s_off[k]), int(r_off[k]))
        return out

    def update_operators(self, levels):
        """Once per Newton step: this rank's blocks of the new operator values, patch factors, coarse inverse.
        `levels` = the global LevelInputs, or None to (re)use the values the LocalLevels carry (rank-local problems)."""
        c = self.ctx
        c.set_bsr_values(0, (levels[0] if levels is not None else self._levels0).vals)
        for l in range(1, self.nlevels):
            ll = self.local[l]
            vals = ll.vals if levels is None else np.asarray(levels[l].vals)[ll.vals_sel]
            c.set_bsr_values(l, np.ascontiguousarray(vals))
            c.factor(l)
        c.coarse_factor()
        self._host_barrier()

    def update_transfers(self, levels):
        for l in range(1, self.nlevels):
            ll = self.local[l]
            if levels is None:
                if ll.a0_vals is not None:
                    self.ctx.transfer_update(l, np.ascontiguousarray(ll.a0_vals), np.ascontiguousarray(ll.d_vals))
            elif levels[l].a0_vals is not None:
                self.ctx.transfer_update(l, np.ascontiguousarray(np.asarray(levels[l].a0_vals)[ll.vals_sel]),
                                         np.ascontiguousarray(np.asarray(levels[l].d_vals)[ll.vals_sel]))

    # ---- vectors
    @property
    def local_dofs(self):
        """Global dof of every local dof of the finest level (owned first)."""
        return self.local[-1].local_dofs if self.nlevels > 1 else None

    @property
    def n_owned(self):
        return self.local[-1].n_owned

    def scatter(self, x_global):
        """Global finest-level vector -> this rank's local vector (ghosts consistent)."""
        return np.ascontiguousarray(np.asarray(x_global)[self.local_dofs])

    def gather(self, x_local):
        """This rank's local vector -> the global vector on every rank (sum of the owned parts over the ranks)."""
        import torch
        import torch.distributed as dist
        ll = self.local[-1]
        xl = x_local if isinstance(x_local, torch.Tensor) else torch.from_numpy(np.asarray(x_local))
        out = torch.zeros(self.layouts[-1].ndofs, dtype=torch.float64, device=xl.device)
        out[torch.from_numpy(ll.local_dofs[:ll.n_owned]).to(xl.device)] = xl[:ll.n_owned]
        if self.nranks > 1:
            dist.all_reduce(out)
        return out

    def apply(self, b_local, x_local):
        """x = one fieldsplit_0 application of b, both LOCAL vectors of the finest level (owned part valid)."""
        return self.ctx.cycle_apply(b_local, x_local)


class DeviceBackend:
    """`fieldsplit_0` backend for alfi_b200.synth.outer.ContinuationSolver on the GPU."""

    def __init__(self, smoothing, device=0, deterministic=False, **kw):
        self.smoothing, self.device, self.deterministic, self.kw = smoothing, device, deterministic, kw
        self.mg = None

    def setup(self, levels):
        self.mg = DeviceMultigrid(levels, self.smoothing, device=self.device, deterministic=self.deterministic, **self.kw)
        self._out = np.empty(levels[-1].n_nodes * levels[-1].bs)

    def update_operators(self, levels):
        self.mg.update_operators(levels)

    def update_transfers(self, levels):
        self.mg.update_transfers(levels)

    def apply(self, b):
        out = np.empty_like(self._out)
        self.mg.apply(np.ascontiguousarray(b), out)
        return out

    # -- the outer pieces on the device (include/alfib.h "outer Schur-complement fieldsplit")
    def setup_outer(self, B, Minv, bc_dofs, remove_constant=True):
        """B = (div u, q) block of the Jacobian (pressure x velocity dofs), Minv = inverse pressure mass matrix,
        bc_dofs = Dirichlet velocity dofs: their columns of B are removed, as Firedrake's bcs do for the
        off-diagonal blocks of the assembled Jacobian."""
        import scipy.sparse as sp
        keep = np.ones(B.shape[1])
        keep[np.asarray(bc_dofs, dtype=np.int64)] = 0.0
        Bz = (B.tocsr() @ sp.diags(keep)).tocsr()
        Bz.eliminate_zeros()
        self.mg.ctx.schur_set(Bz, Minv, remove_constant)
        self._n_outer = B.shape[0] + B.shape[1]

    def schur_apply(self, nu, gamma, r):
        """One application of the Schur-complement fieldsplit preconditioner to r = [r_u; r_p]."""
        return self.mg.ctx.schur_apply(nu, gamma, np.ascontiguousarray(r), np.empty(self._n_outer))

    def jacobian_apply(self, z):
        return self.mg.ctx.jacobian_apply(np.ascontiguousarray(z), np.empty(self._n_outer))

    def outer_solve(self, nu, gamma, rhs, rtol, atol, maxit=500, restart=30):
        """The whole outer FGMRES solve of one Newton step; returns (x, iterations, residual history)."""
        return self.mg.ctx.outer_solve(nu, gamma, np.ascontiguousarray(rhs), np.empty(self._n_outer), rtol, atol, maxit, restart)
