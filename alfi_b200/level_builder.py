"""The coarse-grained hand-over (`LevelInput` per level, coarsest first) assembled from what a host framework can provide
level by level — the part of `FiredrakeAdapter.levels()` / `firedrake_adapter.transfer_backend()` that is not a Firedrake
call, so that it runs, and is tested, against a stand-in over the synthetic problems (tests/test_level_builder.py).

``access`` protocol (`alfi_b200.firedrake_adapter.FiredrakeAccess` for a live alfi solver):

    nlevels()                       number of levels of the velocity hierarchy, coarsest = 0
    space(l)                        object with ``nnodes``, ``bs``, ``cell_nodes`` (cells x nodes per cell) and what
                                    ``plex(l).node_points`` needs (the PetscSection of the space)
    plex(l)                         DMPlexView-like: closure / star CSR, labels, strata, node_points(V), point_coords(p)
    coarse_to_fine_cells(l)         (cells of level l) x (children) cell numbers of level l + 1   (l < nlevels - 1;
                                    HierarchyBase.coarse_to_fine_cells, alfi/bary.py:113-119, 130-170)
    bc_nodes(l)                     nodes of the homogeneous Dirichlet condition of the velocity on level l
    operator_blocks(l)              (rowptr, colidx, vals (nnzb, bs, bs)) of the rediscretised velocity block
    transfer_blocks(l, nu, gamma)   (A0 vals, gamma D vals) on the same pattern (alfi/transfer.py:293-332)
    prolong(l, coarse)              the solver's standard transfer level l - 1 -> l (alfi/transfer.py:284-290; BubbleTransfer
                                    for [P1+FB]^3) applied to an array: nodal (n_coarse_nodes,) -> (n_fine_nodes,), or with
                                    ``dof_level_transfer`` (n_coarse_dofs,) -> (n_fine_dofs,)
    parameters()                    (nu, gamma) as floats
    dof_level_transfer              True if the standard transfer couples the vector components (BubbleTransfer)

The standard prolongation is not a matrix anywhere in Firedrake (it is a par_loop of point evaluations), so it is
recovered by coloured probing: coarse nodes that can never influence the same fine node share a probe vector, one
application of the framework's own ``prolong`` per colour, once per mesh.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .multigrid import LevelInput
from .patches import greedy_colouring, macro_interior_blocks, patch_dofs_from_points, sweep_stages
from .relaxation import iteration_order, macro_star_points, star_points
from .transfer import cell_patch_set

__all__ = ["prolongation_candidates", "colour_candidates", "probe_prolongation", "build_level_inputs"]


def prolongation_candidates(fine_cell_nodes, coarse_cell_nodes, c2f, n_fine, n_coarse):
    """Boolean CSR (fine nodes x coarse nodes): the coarse nodes the prolonged value at a fine node can depend on =
    the nodes of every coarse cell that is a parent of a fine cell holding the node."""
    c2f = np.asarray(c2f)
    ncc, nchild = c2f.shape
    kf, kc = fine_cell_nodes.shape[1], coarse_cell_nodes.shape[1]
    fnodes = fine_cell_nodes[c2f.ravel()].reshape(ncc, nchild * kf)               # fine nodes under every coarse cell
    rows = np.repeat(fnodes, kc, axis=1).ravel()
    cols = np.tile(coarse_cell_nodes[:ncc], (1, nchild * kf)).ravel()
    C = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n_fine, n_coarse))
    C.sum_duplicates()
    C.data[:] = 1
    C.sort_indices()
    return C


def colour_candidates(C):
    """Greedy colouring of the coarse nodes (in index order) such that two candidates of one fine node never share a
    colour.  Returns (colour per coarse node, number of colours); nodes that are nobody's candidate get colour 0."""
    C = C.tocsr()
    CT = C.T.tocsr()
    n_coarse = C.shape[1]
    colour = np.full(n_coarse, -1, dtype=np.int32)
    ncol = 0
    for c in range(n_coarse):
        rows = CT.indices[CT.indptr[c]:CT.indptr[c + 1]]
        if rows.size == 0:
            colour[c] = 0
            continue
        starts, ends = C.indptr[rows], C.indptr[rows + 1]
        lens = ends - starts
        idx = np.repeat(starts - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(lens.sum())
        used = colour[C.indices[idx]]
        used = np.unique(used[used >= 0])
        k = 0
        for u in used:                          # smallest colour not in the (sorted) used set
            if u != k:
                break
            k += 1
        colour[c] = k
        ncol = max(ncol, k + 1)
    return colour, max(ncol, 1)


def probe_prolongation(apply, C, drop=1e-13):
    """The matrix of the linear map ``apply`` (coarse array -> fine array) whose sparsity lies inside the candidate
    pattern ``C``: one application per colour of `colour_candidates`.  Entries below ``drop`` x the largest entry are
    removed (a point evaluation returns rounding noise where a basis function vanishes).  Raises if the map has
    entries outside ``C`` (checked with one more application on a random vector)."""
    C = C.tocsr()
    n_fine, n_coarse = C.shape
    colour, ncol = colour_candidates(C)
    rows = np.repeat(np.arange(n_fine), np.diff(C.indptr))
    cols = C.indices
    vals = np.zeros(cols.size)
    ecol = colour[cols]
    for k in range(ncol):
        x = (colour == k).astype(np.float64)
        y = np.asarray(apply(x), dtype=np.float64).reshape(-1)
        if y.size != n_fine:
            raise ValueError("prolong returned %d values for %d fine entries" % (y.size, n_fine))
        sel = ecol == k
        vals[sel] = y[rows[sel]]
    keep = np.abs(vals) > drop * max(np.abs(vals).max(), 1e-300)
    P = sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n_fine, n_coarse))
    P.sort_indices()
    x = np.random.default_rng(0).standard_normal(n_coarse)
    y = np.asarray(apply(x), dtype=np.float64).reshape(-1)
    err = np.abs(P @ x - y).max()
    if err > 1e-10 * max(np.abs(y).max(), 1.0):
        raise ValueError("the standard prolongation has entries outside the candidate pattern (defect %.2e)" % err)
    return P


class _LevelView:
    """What `alfi_b200.transfer.cell_patch_set` reads of a level."""

    def __init__(self, plex, c2f):
        self.plex, self.c2f = plex, c2f


def _smoother_patches(plex, V, bc_nodes, construct, sort_order, macro_expand, condensed, composition, rowptr, colidx):
    name = (construct or "star").rpartition(".")[2]
    if name == "MacroStar":
        H, ents = macro_star_points(plex, macro_expand)
    elif name in ("star", "Star"):
        H, ents = star_points(plex)
    else:
        raise NotImplementedError("patch construction %r" % construct)
    order = None
    if sort_order:
        coords = np.array([plex.point_coords(p) for p in ents])
        order = iteration_order(coords, sort_order)
    ps = patch_dofs_from_points(plex, V, H, bc_nodes=bc_nodes, order=order)
    greedy_colouring(ps, V.nnodes * V.bs)
    if condensed:
        ps.blocks = macro_interior_blocks(plex, V, ps)
    if composition == "multiplicative":
        ps.stages = sweep_stages(ps, rowptr, colidx)
    return ps


def build_level_inputs(access, construct="star", sort_order=None, macro_expand="all", bary=False,
                       composition="additive", smoother=True, transfer=True, prolongations=None):
    """`LevelInput` of every level, coarsest first.  ``construct`` / ``sort_order`` / ``composition`` are what
    `alfi_b200.pc.fieldsplit0_config` reads from the reference's dictionary; ``bary`` = barycentric hierarchy
    (macro-cell patches for the transfer, condensed block structure, alfi/transfer.py:111).  ``smoother=False``
    leaves out the operator values and the smoother's patches (what a transfer backend needs); ``prolongations``
    caches the probed ``P_H`` per level across calls (they depend on the meshes only)."""
    nl = access.nlevels()
    nu, gamma = access.parameters()
    views = []
    out = []
    for l in range(nl):
        V, plex = access.space(l), access.plex(l)
        views.append(_LevelView(plex, access.coarse_to_fine_cells(l) if l < nl - 1 else None))
        rowptr, colidx, vals = access.operator_blocks(l)
        bc_nodes = np.asarray(access.bc_nodes(l), dtype=np.int32)
        bs = V.bs
        bc_dofs = (bc_nodes[:, None] * bs + np.arange(bs)[None, :]).ravel().astype(np.int32)
        li = LevelInput(V.nnodes, bs, np.asarray(rowptr, np.int32), np.asarray(colidx, np.int32),
                        np.asarray(vals, np.float64) if smoother else None, bc_dofs)
        if smoother and l > 0:
            ps = _smoother_patches(plex, V, bc_nodes, construct, sort_order, macro_expand, bary, composition,
                                   li.rowptr, li.colidx)
            li.patch_offsets, li.patch_dofs, li.patch_order, li.patch_colours = ps.offsets, ps.dofs, ps.order, ps.colours
            li.patch_blocks = ps.blocks
            li.patch_stages = getattr(ps, "stages", None)
            li.symmetrise_sweep = composition == "multiplicative"
        if smoother and l == 0 and bary:
            from .patches import PatchSet
            free = np.setdiff1d(np.arange(V.nnodes * bs), bc_dofs).astype(np.int32)
            one = PatchSet(offsets=np.array([0, free.size], np.int64), dofs=free, bs=bs, order=np.zeros(1, np.int32))
            blocks = macro_interior_blocks(plex, V, one)
            if blocks is not None and (blocks >= 0).any():
                li.coarse_dofs, li.coarse_blocks = free, blocks
        if l > 0:
            Vc = access.space(l - 1)
            key = l
            if prolongations is not None and key in prolongations:
                li.P = prolongations[key]
            else:
                C = prolongation_candidates(np.asarray(V.cell_nodes), np.asarray(Vc.cell_nodes), views[l - 1].c2f,
                                            V.nnodes, Vc.nnodes)
                if getattr(access, "dof_level_transfer", False):
                    C = sp.kron(C, np.ones((bs, bs), dtype=np.int8), format="csr")
                li.P = probe_prolongation(lambda x, l=l: access.prolong(l, x), C)
                if prolongations is not None:
                    prolongations[key] = li.P
            li.P_dof_level = bool(getattr(access, "dof_level_transfer", False))
            if transfer:
                cells, cb_nodes = cell_patch_set(views, l, V, bary)
                if bary:
                    cells.blocks = macro_interior_blocks(plex, V, cells)
                li.cell_offsets, li.cell_dofs, li.cell_blocks = cells.offsets, cells.dofs, cells.blocks
                li.cb_dofs = (np.asarray(cb_nodes)[:, None] * bs + np.arange(bs)[None, :]).ravel().astype(np.int32)
                li.a0_vals, li.d_vals = access.transfer_blocks(l, nu, gamma)
        out.append(li)
    return out
