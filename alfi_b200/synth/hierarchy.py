"""Uniform and barycentric mesh hierarchies + the standard (point-evaluation) prolongation.

Stands in for `alfi.bary.BaryMeshHierarchy` (alfi/bary.py:29-194) / Firedrake `MeshHierarchy`
(alfi/problem.py:10-24) and for Firedrake's `prolong` (alfi/transfer.py:284-290), host side.

Conventions restated from the reference:

* the uniform hierarchy is refined first, every level is then Alfeld-split (bary.py:89);
* ``coarse_to_fine_cells[l]`` maps every coarse *bary* cell to all ``(d+1)*2^d`` fine bary cells
  of the same macro cell (bary.py:141-157);
* facets of the uniform mesh of level ``j`` carry the label ``prolongation = j+1`` and refined
  facets inherit it (solver.py:101-108), so on the fine mesh of level ``L`` the facets with
  ``0 <= label <= L`` are exactly those lying on facets of the level ``L-1`` uniform mesh;
* the hierarchy is *not nested* for bary meshes (bary.py:193), so the standard prolongation
  evaluates the coarse function at each fine node inside the candidate coarse cell that
  contains it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .fem import LagrangeElement, VectorSpace
from .mesh import SimplexMesh, alfeld_split, kuhn_mesh, locate_in_kuhn, refine_uniform
from .plex import SynthPlex

__all__ = ["Level", "build_hierarchy", "build_hierarchy_from", "prolongation_matrix"]


@dataclass
class Level:
    index: int
    macro: SimplexMesh                   # uniform (Kuhn) mesh of this level
    mesh: SimplexMesh                    # the mesh the spaces live on (Alfeld split or == macro)
    plex: SynthPlex
    bary: bool
    # to the next finer level (None on the finest)
    macro_c2f: np.ndarray | None = None  # (nc_macro, 2^d) fine macro cells per macro cell
    c2f: np.ndarray | None = None        # (nc, .) fine mesh cells per mesh cell


def _label_prolongation(level: Level, coarse: Level | None):
    """Set the `prolongation` label on the facets of level.mesh (see module docstring)."""
    m, plex, d = level.mesh, level.plex, level.mesh.dim
    nfac = m.facets.shape[0]
    f0 = plex.getHeightStratum(1)[0]
    lab = np.full(plex.npoints, -1, dtype=np.int64)
    # every facet of the uniform mesh of this level gets index+1 ...
    if level.bary:
        on_macro = m.macro_vertex[m.facets].all(axis=1)
    else:
        on_macro = np.ones(nfac, dtype=bool)
    lab[f0:f0 + nfac][on_macro] = level.index + 1
    # ... unless it descends from a coarser facet, whose (smaller) value it inherits
    if coarse is not None:
        # parent coarse macro cell of every cell of this mesh
        parent_of_macro = np.empty(level.macro.nc, dtype=np.int64)
        parent_of_macro[coarse.macro_c2f.ravel()] = np.repeat(np.arange(coarse.macro.nc),
                                                               coarse.macro_c2f.shape[1])
        cell_macro = np.arange(m.nc) // (d + 1) if level.bary else np.arange(m.nc)
        parent = parent_of_macro[cell_macro]
        cf = m.cell_facets
        # facet -> (min parent, max parent, count) over its support cells
        lo = np.full(nfac, np.iinfo(np.int64).max)
        hi = np.full(nfac, -1)
        cnt = np.zeros(nfac, dtype=np.int64)
        par = np.repeat(parent, cf.shape[1])
        np.minimum.at(lo, cf.ravel(), par)
        np.maximum.at(hi, cf.ravel(), par)
        np.add.at(cnt, cf.ravel(), 1)
        inherited = on_macro & ((lo != hi) | (cnt == 1))
        coarse_lab = coarse.plex.labels["prolongation"]
        # all coarse macro facets carry <= coarse.index+1; the exact inherited value only matters
        # through the test `0 <= value <= level` (transfer.py:36-38,132-133)
        lab[f0:f0 + nfac][inherited] = min(coarse.index + 1, int(coarse_lab.max()))
    plex.labels["prolongation"] = lab


def build_hierarchy(dim: int, N: int, nref: int, bary: bool, length: float = 2.0, shape: tuple = (),
                    counts: tuple = (), origin: tuple = ()) -> list[Level]:
    """`counts` / `origin` (cells of the coarsest level, size length / N): the hierarchy over a sub-box of the grid."""
    levels: list[Level] = []
    for l in range(nref + 1):
        macro = kuhn_mesh(dim, N * 2 ** l, length, shape, tuple(c * 2 ** l for c in counts), tuple(o * 2 ** l for o in origin))
        mesh = alfeld_split(macro) if bary else macro
        levels.append(Level(l, macro, mesh, SynthPlex(mesh), bary))
    d = dim
    for c, f in zip(levels[:-1], levels[1:]):
        cent = f.macro.coords[f.macro.cells].mean(axis=1)
        par = locate_in_kuhn(c.macro, cent)
        order = np.argsort(par, kind="stable")
        c.macro_c2f = order.reshape(c.macro.nc, 2 ** d)
        assert (par[c.macro_c2f] == np.arange(c.macro.nc)[:, None]).all()
        if bary:
            fine = (c.macro_c2f[:, :, None] * (d + 1) + np.arange(d + 1)[None, None, :]).reshape(c.macro.nc, -1)
            c.c2f = np.repeat(fine, d + 1, axis=0)          # same list for the d+1 sub-cells
        else:
            c.c2f = c.macro_c2f
    for i, lev in enumerate(levels):
        _label_prolongation(lev, levels[i - 1] if i else None)
    return levels


def build_hierarchy_from(base: SimplexMesh, nref: int, bary: bool) -> list[Level]:
    """The same hierarchy over a general base mesh (a Gmsh file, examples/bfs2d/bfs2d.py:14-17): `nref`
    uniform refinements (alfi/problem.py:10-24), every level Alfeld-split for `bary` (alfi/bary.py:89)."""
    d = base.dim
    macros, c2fs = [base], []
    for _ in range(nref):
        fine, c2f = refine_uniform(macros[-1])
        macros.append(fine)
        c2fs.append(c2f)
    levels = []
    for l, macro in enumerate(macros):
        mesh = alfeld_split(macro) if bary else macro
        levels.append(Level(l, macro, mesh, SynthPlex(mesh), bary))
    for c, c2f in zip(levels[:-1], c2fs):
        c.macro_c2f = c2f
        if bary:
            fine = (c2f[:, :, None] * (d + 1) + np.arange(d + 1)[None, None, :]).reshape(c.macro.nc, -1)
            c.c2f = np.repeat(fine, d + 1, axis=0)
        else:
            c.c2f = c2f
    for i, lev in enumerate(levels):
        _label_prolongation(lev, levels[i - 1] if i else None)
    return levels


def prolongation_matrix(Vc: VectorSpace, Vf: VectorSpace, c2f: np.ndarray, drop_tol: float = 1e-13):
    """Scalar CSR ``P`` (fine nodes x coarse nodes) of the standard prolongation.

    ``c2f[c]`` lists the fine cells that are candidates for coarse cell ``c`` (the inverse map
    is Firedrake's ``fine_to_coarse_cells``, bary.py:173-184).  For every fine node the
    candidate coarse cell with the largest minimal barycentric coordinate is used (ties:
    lowest cell id), and the coarse basis is evaluated there.
    """
    mc = Vc.mesh
    d = mc.dim
    nlf = Vf.cell_nodes.shape[1]
    ncc, nfc = c2f.shape
    # candidate (coarse cell, fine node) pairs
    cc = np.repeat(np.arange(ncc), nfc * nlf)
    fn = Vf.cell_nodes[c2f.ravel()].ravel()
    key = np.unique(fn * np.int64(ncc) + cc)
    fn, cc = key // ncc, key % ncc
    # barycentric coordinates of the fine nodes in their candidate coarse cells
    X = mc.coords[mc.cells[cc]]                                   # (np, d+1, d)
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
    xi = np.linalg.solve(J, (Vf.node_coords[fn] - X[:, 0, :])[:, :, None])[:, :, 0]
    lam = np.concatenate([1.0 - xi.sum(axis=1, keepdims=True), xi], axis=1)
    score = np.round(lam.min(axis=1), 10)
    order = np.lexsort((cc, -score, fn))                          # by node, best score, cell id
    first = np.flatnonzero(np.concatenate(([True], fn[order][1:] != fn[order][:-1])))
    sel = order[first]
    assert sel.size == Vf.nnodes, "every fine node needs a candidate coarse cell"
    assert score[sel].min() > -1e-8, "fine node outside all candidate coarse cells"
    el = Vc.element
    vals = el.tabulate(xi[sel])                                   # (nf, nlc)
    cols = Vc.cell_nodes[cc[sel]]
    rows = np.repeat(fn[sel], vals.shape[1])
    keep = np.abs(vals.ravel()) > drop_tol
    P = sp.csr_matrix((vals.ravel()[keep], (rows[keep], cols.ravel()[keep])),
                      shape=(Vf.nnodes, Vc.nnodes))
    P.sum_duplicates()
    P.sort_indices()
    return P
