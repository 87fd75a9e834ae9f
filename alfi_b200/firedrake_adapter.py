"""HostAdapter for a real Firedrake/petsc4py deployment.

NOT exercised in this repository's CI: firedrake, petsc4py and mpi4py are absent from the image
(SURVEY fact 3).  It is written against the documented petsc4py / Firedrake API that alfi itself
uses (alfi/solver.py:15-38, alfi/transfer.py:121-158) and mirrors, method for method, the
`SynthAdapter` that *is* tested (alfi_b200/synth/fakepetsc.py).  Imports are lazy so that the
rest of the package never needs Firedrake.
"""
from __future__ import annotations

import numpy as np

from .pc import HostAdapter

__all__ = ["FiredrakeAdapter", "attach", "transfer_backend"]


def baij_csr_to_blocks(indptr, indices, data, bs):
    """Scalar CSR of a BAIJ matrix (every block fully stored, as MatGetValuesCSR returns it) ->
    (block rowptr, block colidx, values (nnzb, bs, bs) row-major blocks, block_col_major=False)."""
    n = indptr.size - 1
    counts = np.diff(indptr)
    assert n % bs == 0 and (counts % bs == 0).all(), "not a fully stored block matrix"
    bcount = counts[::bs] // bs
    rowptr = np.concatenate(([0], np.cumsum(bcount))).astype(np.int32)
    first = indptr[:-1:bs]                                     # first scalar row of every block row
    lead = np.repeat(first, bcount * bs) + np.concatenate([np.arange(c * bs) for c in bcount]) if n else np.empty(0, int)
    colidx = (indices[lead][::bs] // bs).astype(np.int32)
    row_of = np.repeat(np.arange(n), counts)                   # scalar row of every stored entry
    pos = np.arange(indices.size) - indptr[row_of]             # position inside its row
    blk = rowptr[row_of // bs] + pos // bs
    vals = np.empty((colidx.size, bs, bs))
    vals[blk, row_of % bs, pos % bs] = data
    return rowptr, colidx, vals, False


class _Space:
    """The view of a Firedrake FunctionSpace the patch builders need."""

    def __init__(self, V):
        self.V = V
        self.bs = V.value_size
        self.cell_nodes = np.asarray(V.cell_node_list, dtype=np.int64)
        self.nnodes = V.dof_dset.total_size

    def node_points(self, plex):
        """node -> DMPlex point, from the section of the scalar space (transfer.py:127-144)."""
        section = self.V.dm.getDefaultSection()
        pStart, pEnd = section.getChart()
        out = np.full(self.nnodes, -1, dtype=np.int64)
        for p in range(pStart, pEnd):
            dof, off = section.getDof(p), section.getOffset(p)
            if dof:
                out[off // self.bs:(off + dof) // self.bs] = p
        return out


class FiredrakeAdapter(HostAdapter):
    def __init__(self, device=0, deterministic=False):
        self.device, self.deterministic = device, deterministic

    def options(self, pc):
        from firedrake.petsc import PETSc
        return PETSc.Options(pc.getOptionsPrefix())

    def operator(self, pc):
        _, P = pc.getOperators()
        bs = P.getBlockSize()
        indptr, indices, data = P.getValuesCSR()          # scalar CSR of the BAIJ matrix
        return baij_csr_to_blocks(np.asarray(indptr), np.asarray(indices), np.asarray(data), bs)


    def function_space(self, pc):
        from firedrake import dmhooks
        return _Space(dmhooks.get_function_space(pc.getDM()))

    def bc_nodes(self, pc):
        from firedrake.dmhooks import get_appctx
        ctx = get_appctx(pc.getDM())
        bcs = getattr(ctx, "bcs", None) or getattr(getattr(ctx, "_problem", None), "bcs", ())
        nodes = [np.asarray(bc.nodes) for bc in bcs]
        return np.unique(np.concatenate(nodes)) if nodes else np.empty(0, np.int32)


def attach(solver, device=None):
    """Put an adapter on every level's DM so `alfi_b200.PatchPC` finds it."""
    rank = solver.mesh.comm.rank
    ad = FiredrakeAdapter(device=rank % 8 if device is None else device)
    for mesh in solver.mh:
        mesh._topology_dm.setAttr("alfi_b200_adapter", ad)
    return ad


def transfer_backend(solver):
    """kwargs for `alfi_b200.SVSchoeberlTransfer(..., **transfer_backend(solver))`: the device
    context that holds the levels and a callback re-assembling (A0, gamma*D) values when
    (nu, gamma) change (alfi/transfer.py:238-244)."""
    raise NotImplementedError("needs a live Firedrake solver: assemble transfer.form / bform per level "
                              "and return {'backend': ctx, 'values_for': callback}")
