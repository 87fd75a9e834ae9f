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
    cells: object | None = None           # scipy CSR (npatch x ncells): the patch cells PCPATCH integrates over
    corrections: "PatchCorrections | None" = None   # A_i = A[I_i, I_i] + C_i (forms with interior-facet integrals)
    corr_vals: np.ndarray | None = None   # ... the entries of C for the current operator values (per Newton step)

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


def patch_cells(plex, H):
    """CSR (npatch x ncells): the cells in the star of any point of a patch's point set — the cells PCPATCH
    integrates over."""
    Hc = (H.tocsr().astype(np.int32) @ plex.star.astype(np.int32)).tocsr()[:, plex.cStart:plex.cEnd].tocsr()
    Hc.sort_indices()
    return Hc


def patch_dofs_from_points(plex, V, H, bc_nodes=None, order=None) -> PatchSet:
    """Patch dof lists for point sets H (CSR npatch x npoints) on space V (see module doc)."""
    npatch = H.shape[0]
    bs = V.bs
    nn = np.int64(V.nnodes)
    H = H.tocsr().astype(np.int32)
    # patch cells: cells in the star of any point of ht
    Hc = patch_cells(plex, H)
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
    return PatchSet(offsets, dofs, np.asarray(order, dtype=np.int32), bs, cells=Hc)


@dataclass
class PatchCorrections:
    """What separates PCPATCH's patch operator from the sub-matrix of the assembled operator when the form has
    interior-facet integrals (Burman's stabilisation, alfi/stabilisation.py:156-162): PCPATCH integrates
    over the patch cells and over the facets whose BOTH cells are patch cells, while A[I_i, I_i] also holds the
    inside-inside part of the facets on the patch boundary.  A_i = A[I_i, I_i] + C_i with C_i in COO form, patch-local
    indices, sorted by (patch, row, col); `values(S)` turns the macro-element tensors of the interior facets
    (`synth.fem.burman_facet_tensors`) into the entries: minus the sum of the boundary facets' inside-inside blocks."""
    off: np.ndarray            # int64 (npatch + 1)
    rows: np.ndarray           # int32 patch-local row of every entry
    cols: np.ndarray
    src_f: np.ndarray          # entry slot[e] -= S[src_f[e], src_i[e], src_j[e]]
    src_i: np.ndarray
    src_j: np.ndarray
    slot: np.ndarray

    def values(self, S, scale=1.0):
        return -scale * np.bincount(self.slot, weights=S[self.src_f, self.src_i, self.src_j], minlength=self.rows.size)


def facet_corrections(V, ps: PatchSet, facet_cells) -> PatchCorrections:
    """Patch-boundary facets of every patch of `ps` (interior facets of the mesh with exactly one of their two cells
    among the patch cells `ps.cells`) and the entries their inside-inside blocks touch.  `facet_cells` (nF, 2): the
    cells of the interior facets, in the numbering of the facet tensors."""
    Hc = ps.cells.tocsr()
    npatch, nc = Hc.shape
    bs, nl = ps.bs, V.cell_nodes.shape[1]
    nn = np.int64(V.nnodes)
    nF = facet_cells.shape[0]
    # (cell, side) -> interior facets
    cs = np.concatenate([facet_cells[:, 0], facet_cells[:, 1]])
    fs = np.concatenate([np.arange(nF), np.arange(nF)])
    ss = np.concatenate([np.zeros(nF, np.int64), np.ones(nF, np.int64)])
    o = np.argsort(cs, kind="stable")
    cs, fs, ss = cs[o], fs[o], ss[o]
    cptr = np.searchsorted(cs, np.arange(nc + 1))
    # every (patch, patch cell, interior facet of the cell)
    p_of = np.repeat(np.arange(npatch, dtype=np.int64), np.diff(Hc.indptr))
    cells = Hc.indices.astype(np.int64)
    cnt = cptr[cells + 1] - cptr[cells]
    pp = np.repeat(p_of, cnt)
    idx = np.repeat(cptr[cells] - np.concatenate(([0], np.cumsum(cnt)[:-1])), cnt) + np.arange(cnt.sum())
    ff, sd = fs[idx], ss[idx]
    other = facet_cells[ff, 1 - sd]
    pairs = p_of * nc + cells                                     # sorted: CSR rows ascending, indices sorted
    pos = np.searchsorted(pairs, pp * nc + other)
    pos[pos == pairs.size] = 0
    boundary = pairs[pos] != pp * nc + other
    pp, ff, sd = pp[boundary], ff[boundary], sd[boundary]
    # patch-local position of the inside cell's nodes
    pnode = ps.dofs[::bs].astype(np.int64) // bs                  # patch nodes in patch-local order
    ppatch = np.repeat(np.arange(npatch, dtype=np.int64), np.diff(ps.offsets) // bs)
    plocal = np.arange(pnode.size) - np.repeat(ps.offsets[:-1] // bs, np.diff(ps.offsets) // bs)
    pk = ppatch * nn + pnode
    ok = np.argsort(pk)
    pk_s, pl_s = pk[ok], plocal[ok]
    inside = facet_cells[ff, sd]
    nodes = V.cell_nodes[inside]                                  # (nb, nl)
    key = pp[:, None] * nn + nodes
    q = np.searchsorted(pk_s, key)
    q[q == pk_s.size] = 0
    loc = np.where(pk_s[q] == key, pl_s[q], -1)                   # (nb, nl): -1 = not a dof of the patch
    I, J = np.meshgrid(np.arange(nl), np.arange(nl), indexing="ij")
    li, lj = loc[:, I.ravel()], loc[:, J.ravel()]                 # (nb, nl*nl)
    keep = (li >= 0) & (lj >= 0)
    b_idx, e_idx = np.nonzero(keep)
    li, lj = li[keep], lj[keep]
    si = sd[b_idx] * nl + I.ravel()[e_idx]
    sj = sd[b_idx] * nl + J.ravel()[e_idx]
    ent_p, ent_f = pp[b_idx], ff[b_idx]
    # one entry per velocity component (the form is the same scalar matrix for each)
    comp = np.arange(bs)
    rows = (li[:, None] * bs + comp[None, :]).ravel()
    cols = (lj[:, None] * bs + comp[None, :]).ravel()
    ent_p, ent_f, si, sj = (np.repeat(a, bs) for a in (ent_p, ent_f, si, sj))
    nmax = np.int64(max(int(np.diff(ps.offsets).max()) if npatch else 1, 1))
    k = (ent_p * nmax + rows) * nmax + cols
    uk, slot = np.unique(k, return_inverse=True)
    up = uk // (nmax * nmax)
    off = np.zeros(npatch + 1, dtype=np.int64)
    np.cumsum(np.bincount(up, minlength=npatch), out=off[1:])
    return PatchCorrections(off, ((uk // nmax) % nmax).astype(np.int32), (uk % nmax).astype(np.int32),
                            ent_f, si, sj, slot)


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
