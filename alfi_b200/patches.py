"""PCPATCH dof-set construction and patch colouring (host side, vectorised numpy).

Restates what PETSc's PCPATCH derives from the point sets handed to it by the constructors of
:mod:`alfi_b200.relaxation` / :mod:`alfi_b200.transfer` (selected in alfi/solver.py:318-344 and
alfi/transfer.py:100-113); PETSc's source is not in the reference tree, the semantics are
those written down in SURVEY.md Appendix A.1:

* ``ht``  = the user's point set; ``cht`` = closure of every cell in the star of a point of ht;
* patch cells = cells of ``cht`` (ascending);
* patch dofs = dofs attached to points of ``ht`` minus the global Dirichlet dofs; dofs seen on
  ``cht`` but not attached to ``ht`` are artificial boundary conditions and are dropped;
* local numbering = first encounter walking patch cells in order, nodes in ``cell_node_list``
  order, components innermost;
* patches with no dofs are kept (empty) so patch indices stay aligned with the iteration set.

The literal, loop-based restatement used as the checker is ``oracle/pcpatch.py``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

__all__ = ["PatchSet", "patch_dofs_from_points", "points_to_csr", "greedy_colouring", "macro_interior_blocks",
           "MAX_BLOCK_DOFS"]

MAX_BLOCK_DOFS = 64        # limit of the condensed form (csrc/condense.cu: one warp tile per block)


@dataclass
class PatchSet:
    offsets: np.ndarray        # int64 (npatch+1)
    dofs: np.ndarray           # int32 scalar dof indices, patch-local order
    order: np.ndarray          # int32 iteration set (indices into patches)
    bs: int
    colours: np.ndarray | None = None     # int32 (npatch,), greedy colouring in iteration order
    blocks: np.ndarray | None = None      # int32 per dof entry: -1 separator, else block label (condensed form)
    centres: np.ndarray | None = None     # (npatch, dim) coordinates of the patches' entities (partitioning only)
    stages: np.ndarray | None = None      # multiplicative composition: stage per entry of `order` (sweep_stages)
    symmetrise: bool = False              # ... followed by the backward sweep (patch_pc_patch_symmetrise_sweep)

    @property
    def npatch(self):
        return self.offsets.size - 1

    @property
    def sizes(self):
        return np.diff(self.offsets)

    def patch(self, i):
        return self.dofs[self.offsets[i]:self.offsets[i + 1]]


def points_to_csr(point_lists, npoints):
    """list of int arrays (possibly with duplicates, as the reference produces) → CSR boolean."""
    lens = np.fromiter((len(p) for p in point_lists), dtype=np.int64, count=len(point_lists))
    rows = np.repeat(np.arange(len(point_lists)), lens)
    cols = np.concatenate([np.asarray(p, dtype=np.int64) for p in point_lists]) if len(point_lists) else np.empty(0, np.int64)
    H = sp.csr_matrix((np.ones(rows.size, dtype=np.int32), (rows, cols)), shape=(len(point_lists), npoints))
    H.sum_duplicates()
    H.data[:] = 1
    H.sort_indices()
    return H


def patch_dofs_from_points(plex, V, H, bc_nodes=None, order=None) -> PatchSet:
    """Patch dof lists for point sets H (CSR npatch x npoints) on space V (see module doc)."""
    npatch = H.shape[0]
    bs = V.bs
    nn = np.int64(V.nnodes)
    H = H.tocsr().astype(np.int32)
    # patch cells: cells in the star of any point of ht
    Hc = (H @ plex.star.astype(np.int32)).tocsr()[:, plex.cStart:plex.cEnd].tocsr()
    Hc.sort_indices()
    # owned nodes: nodes attached to points of ht
    npt = plex.node_points(V)
    NP = sp.csr_matrix((np.ones(V.nnodes, dtype=np.int32), (npt, np.arange(V.nnodes))),
                       shape=(plex.npoints, V.nnodes))
    O = (H @ NP).tocsr()
    O.sort_indices()
    okeys = np.repeat(np.arange(npatch, dtype=np.int64), np.diff(O.indptr)) * nn + O.indices
    # walk (patch, cell, local node)
    nl = V.cell_nodes.shape[1]
    pc_patch = np.repeat(np.arange(npatch, dtype=np.int64), np.diff(Hc.indptr))
    nodes = V.cell_nodes[Hc.indices]                            # (npairs, nl)
    keys = (pc_patch[:, None] * nn + nodes).ravel()
    pos = np.searchsorted(okeys, keys)
    pos[pos == okeys.size] = 0
    keep = okeys[pos] == keys if okeys.size else np.zeros(keys.size, dtype=bool)
    if bc_nodes is not None and len(bc_nodes):
        isbc = np.zeros(V.nnodes, dtype=bool)
        isbc[np.asarray(bc_nodes)] = True
        keep &= ~isbc[nodes.ravel()]
    kept = keys[keep]
    _, first = np.unique(kept, return_index=True)
    first.sort()
    sel = kept[first]                                           # first-encounter order, grouped by patch
    p_of = sel // nn
    node = sel % nn
    counts = np.bincount(p_of, minlength=npatch).astype(np.int64)
    offsets = np.zeros(npatch + 1, dtype=np.int64)
    np.cumsum(counts * bs, out=offsets[1:])
    dofs = (node[:, None] * bs + np.arange(bs)[None, :]).ravel().astype(np.int32)
    if order is None:
        order = np.arange(npatch, dtype=np.int32)
    return PatchSet(offsets, dofs, np.asarray(order, dtype=np.int32), bs)


def greedy_colouring(ps: PatchSet, ndofs: int) -> np.ndarray:
    """Greedy colouring: patches in iteration-set order, lowest free colour, conflict = shared dof.

    The reference has no colouring (PETSc applies patches sequentially); this definition is
    ours and is what makes the device scatter-add race-free and deterministic.
    Patches that appear several times in the iteration set (multi-sweep sort orders) keep the
    colour of their first visit.
    """
    used = np.zeros(ndofs, dtype=np.uint64)
    colours = np.full(ps.npatch, -1, dtype=np.int32)
    one = np.uint64(1)
    for p in ps.order:
        if colours[p] >= 0:
            continue
        d = ps.dofs[ps.offsets[p]:ps.offsets[p + 1]]
        if d.size == 0:
            colours[p] = 0
            continue
        m = int(np.bitwise_or.reduce(used[d]))
        c = 0
        while (m >> c) & 1:
            c += 1
        if c >= 64:
            raise RuntimeError("more than 64 colours needed")
        colours[p] = c
        used[d] |= one << np.uint64(c)
    ps.colours = colours
    return colours


def sweep_stages(ps: PatchSet, rowptr, colidx) -> np.ndarray:
    """Schedule of the SEQUENTIAL (multiplicative) patch sweep of PCApply_PATCH (`pc_patch_local_type multiplicative`,
    alfi/solver.py:306-308,322): stage of every entry of the iteration set, such that two visits whose patches are
    coupled through the operator (block pattern ``rowptr`` / ``colidx``; patches sharing a node are coupled through the
    diagonal block) lie in different stages, the earlier visit in the lower one:
    ``stage[k] = 1 + max(stage[k'] : k' < k, visit k' coupled with visit k)``.  Visits of one stage commute exactly,
    so executing stage after stage reproduces the sequential sweep — forwards, and with the stages reversed the
    backward sweep of `symmetrise_sweep` (handed over with ``alfib_level_set_sweep_stages``)."""
    import scipy.sparse as sp
    npatch, bs = ps.npatch, ps.bs
    nn = len(rowptr) - 1
    nodes = ps.dofs[::bs] // bs                                   # patches are node-wise: bs consecutive dofs per node
    counts = np.diff(ps.offsets) // bs
    Pn = sp.csr_matrix((np.ones(nodes.size, dtype=np.int32), nodes, np.concatenate(([0], np.cumsum(counts)))), shape=(npatch, nn))
    An = sp.csr_matrix((np.ones(len(colidx), dtype=np.int32), colidx, rowptr), shape=(nn, nn))
    C = ((Pn @ An) @ Pn.T).tocsr()                                # C[i, j] != 0: patch i reads what patch j writes
    C = (C + C.T).tocsr()
    stage_of_patch_last = np.full(npatch, -1, dtype=np.int64)     # stage of the latest visit of each patch so far
    stages = np.zeros(ps.order.size, dtype=np.int32)
    for k, p in enumerate(ps.order):
        nb = C.indices[C.indptr[p]:C.indptr[p + 1]]
        s = int(stage_of_patch_last[nb].max()) + 1 if nb.size else 0
        stages[k] = s
        stage_of_patch_last[p] = s
    return stages


def macro_interior_blocks(plex, V, ps: PatchSet, label: str = "MacroVertices") -> np.ndarray | None:
    """Block/separator structure of patches on a barycentrically refined mesh, for the condensed
    form of the patch inverses (``alfib_level_set_patch_blocks``, csrc/condense.cu).

    A point lies in the interior of a macro cell iff it is in the star of that cell's barycentre,
    i.e. of a vertex *not* carrying the ``MacroVertices`` label that alfi/bary.py:16-27 sets and
    ``MacroStar`` already relies on (alfi/relaxation.py:168-177).  Dofs attached to such points
    couple only to dofs of the same macro cell, so per patch they form decoupled blocks (label =
    the barycentre's point number); every other dof is separator (-1).  Returns None when the mesh
    has no macro structure or a block would exceed the 64-dof limit of the condensed kernels (the
    caller then keeps dense inverses).  The library re-checks decoupling against the operator's
    sparsity pattern, so this is a hint that cannot change results.
    """
    lab = plex.labels.get(label) if hasattr(plex, "labels") else None
    if lab is None or ps.dofs.size == 0:
        return None
    verts = np.arange(plex.vStart, plex.vEnd)
    bary = verts[lab[verts] != 1]
    if bary.size == 0:
        return None
    star = plex.star
    block_of_point = np.full(plex.npoints, -1, dtype=np.int64)
    counts = np.diff(star.indptr)[bary]
    pts = np.concatenate([star.indices[star.indptr[b]:star.indptr[b + 1]] for b in bary])
    block_of_point[pts] = np.repeat(bary, counts)
    block_of_node = block_of_point[plex.node_points(V)]
    blocks = block_of_node[ps.dofs // ps.bs].astype(np.int32)
    # size limit: dofs per (patch, label)
    patch_of = np.repeat(np.arange(ps.npatch, dtype=np.int64), np.diff(ps.offsets))
    inb = blocks >= 0
    if inb.any():
        key = patch_of[inb] * np.int64(plex.npoints) + blocks[inb]
        _, cnt = np.unique(key, return_counts=True)
        if cnt.max() > MAX_BLOCK_DOFS:
            return None
    return blocks
