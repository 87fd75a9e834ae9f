"""Robust (Schöberl) transfer — host side; drop-in surface of ``alfi.transfer``.

Mirrors alfi/transfer.py:

* ``CoarseCellPatches`` / ``CoarseCellMacroPatches`` (transfer.py:13-88): patch constructors
  ``obj(pc) -> (patches, iterationSet)``, one patch per coarse (macro) cell, made of the closure
  points of its fine cells minus points whose ``prolongation`` label is in ``[0, level]``;
* ``fix_coarse_boundaries`` (transfer.py:121-158): all dofs in the closure of every fine facet
  inherited from a coarser level;
* ``AutoSchoeberlTransfer.prolong/restrict`` (transfer.py:186-275) with the rebuild-on-
  parameter-change logic of transfer.py:173-184, 238-244;
* ``SVSchoeberlTransfer`` / ``PkP0SchoeberlTransfer`` forms (transfer.py:293-332) and
  ``NullTransfer`` (transfer.py:359-366).

The arithmetic (``P_H`` SpMV, ``gamma*D`` SpMV, block-diagonal cell-patch solve, axpy) runs in
the CUDA library through :class:`alfi_b200.lib.Context`; this module only prepares index sets
and forwards buffers.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .patches import PatchSet, patch_dofs_from_points, points_to_csr
from .relaxation import _make_is

__all__ = ["CoarseCellPatches", "CoarseCellMacroPatches", "coarse_cell_points",
           "fix_coarse_boundaries", "fix_coarse_boundaries_loop", "AutoSchoeberlTransfer",
           "SVSchoeberlTransfer", "PkP0SchoeberlTransfer", "NullTransfer", "device_transfer_backend"]


def _hierarchy_of(pc):
    """(list of levels, fine level index) for the PC's DM.

    With Firedrake this is ``get_level(ctx._x.ufl_domain())`` (transfer.py:19-22); the synthetic
    stand-in stores the same information on the ``ctx`` attribute.
    """
    ctx = pc.getAttr("ctx")
    return ctx.hierarchy, ctx.level


class CoarseCellPatches:
    """One patch per coarse cell (transfer.py:13-46)."""
    macro = False

    def __call__(self, pc):
        dmf = pc.getDM()
        levels, level = _hierarchy_of(pc)
        c2f = levels[level - 1].c2f
        tdim = dmf.getDimension()
        patches = []
        for i, fine_cells in enumerate(c2f):
            # d+1 coarse bary cells map to the same fine cells: build that patch once
            # (transfer.py:69-73; synthetic numbering has firedrake == plex cell ids)
            if self.macro and i % (tdim + 1) != 0:
                continue
            entities = []
            for fp in fine_cells:
                pts, _ = dmf.getTransitiveClosure(int(fp), True)
                for pt in pts:
                    value = dmf.getLabelValue("prolongation", pt)
                    if not (value > -1 and value <= level):
                        entities.append(pt)
            patches.append(_make_is(np.unique(entities)))
        return patches, _make_is(np.arange(len(patches)))


class CoarseCellMacroPatches(CoarseCellPatches):
    """One patch per coarse *macro* cell (transfer.py:49-88)."""
    macro = True


def coarse_cell_points(levels, level, macro):
    """Vectorised CoarseCell[Macro]Patches: CSR boolean (npatch x npoints of the fine plex)."""
    fine, coarse = levels[level], levels[level - 1]
    plex = fine.plex
    d = plex.dim
    c2f = coarse.c2f[::d + 1] if macro else coarse.c2f
    npatch, nf = c2f.shape
    G = sp.csr_matrix((np.ones(c2f.size, dtype=np.int32),
                       (np.repeat(np.arange(npatch), nf), c2f.ravel())), shape=(npatch, plex.npoints))
    H = (G @ plex.closure.astype(np.int32)).tocsr()
    lab = plex.labels["prolongation"]
    keep = ~((lab > -1) & (lab <= level))
    H = (H @ sp.diags(keep.astype(np.int32), format="csr", dtype=np.int32)).tocsr()
    H.eliminate_zeros()
    H.data[:] = 1
    H.sort_indices()
    return H


def fix_coarse_boundaries(plex, V, level):
    """Node list of the homogeneous Dirichlet condition on coarse-level facets, vectorised.

    transfer.py:121-158 walks every facet; if its ``prolongation`` label is in ``[0, level]``
    all dofs of all points in its closure are constrained.  Returns sorted unique *node*
    indices (every component of a node is constrained).
    """
    f0, f1 = plex.getHeightStratum(1)
    lab = plex.labels["prolongation"][f0:f1]
    fac = np.flatnonzero((lab > -1) & (lab <= level)) + f0
    pts = np.unique(plex.closure[fac].indices)
    onpt = np.zeros(plex.npoints, dtype=bool)
    onpt[pts] = True
    return np.flatnonzero(onpt[plex.node_points(V)]).astype(np.int32)


def fix_coarse_boundaries_loop(plex, V, level):
    """Literal loop form of transfer.py:128-144 (used by the tests as the checker)."""
    npt = plex.node_points(V)
    by_point = {}
    for n, p in enumerate(npt):
        by_point.setdefault(int(p), []).append(n)
    nodes = []
    for p in range(*plex.getHeightStratum(1)):
        value = plex.getLabelValue("prolongation", p)
        if value > -1 and value <= level:
            closure, _ = plex.getTransitiveClosure(p)
            for c in closure:
                nodes.extend(by_point.get(int(c), []))
    return np.unique(nodes).astype(np.int32)


def cell_patch_set(levels, level, V, macro) -> tuple[PatchSet, np.ndarray]:
    """(cell-patch dof sets, coarse-boundary node list) for the transfer onto ``level``."""
    plex = levels[level].plex
    H = coarse_cell_points(levels, level, macro)
    cb = fix_coarse_boundaries(plex, V, level)
    return patch_dofs_from_points(plex, V, H, bc_nodes=cb), cb


def _values(x):
    """Read-only flat float64 view of a numpy array or of a Firedrake Function's owned values."""
    if hasattr(x, "dat"):
        return np.ascontiguousarray(x.dat.data_ro, dtype=np.float64).reshape(-1)
    return x


def _buffer(x):
    """The array a device call writes into: the array itself, or a flat buffer for a Firedrake Function."""
    if hasattr(x, "dat"):
        return np.empty(int(np.prod(x.dat.data_ro.shape)), dtype=np.float64)
    return x


def _store(x, buf):
    if hasattr(x, "dat"):
        x.dat.data[...] = buf.reshape(x.dat.data_ro.shape)


class AutoSchoeberlTransfer:
    """``prolong(coarse, fine)`` / ``restrict(fine, coarse)`` on device (transfer.py:91-290).

    ``parameters = (nu, gamma)`` are objects convertible with ``float()`` (Firedrake Constants
    in a deployment).  ``backend`` is an :class:`alfi_b200.lib.Context` whose levels already hold
    the transfer operators (`Context.set_transfer`); ``values_for(level, nu, gamma)`` must return
    the re-assembled ``(A0 values, D values)`` when the parameters changed.
    """

    def __init__(self, parameters, tdim, hierarchy, backend=None, values_for=None):
        self.parameters = parameters
        self.tdim = tdim
        self.hierarchy = hierarchy
        self.prev_parameters = {}
        self.force_rebuild_d = {}
        self.backend = backend
        self.values_for = values_for
        self.patch_constructor = (CoarseCellMacroPatches if hierarchy == "bary" else CoarseCellPatches)

    def force_rebuild(self):
        self.force_rebuild_d = {k: True for k in self.prev_parameters}

    def rebuild(self, key):
        if self.force_rebuild_d.get(key, False):
            self.force_rebuild_d[key] = False
            return True
        prev = self.prev_parameters.get(key, [])
        return any(float(p) != q for q, p in zip(prev, self.parameters))

    def _ensure(self, level):
        first = level not in self.prev_parameters
        if first or self.rebuild(level):
            nu, gamma = (float(p) for p in self.parameters)
            if self.values_for is not None:
                a0, dvals = self.values_for(level, nu, gamma)
                self.backend.transfer_update(level, a0, dvals)
            self.prev_parameters[level] = [float(p) for p in self.parameters]

    def prolong(self, coarse, fine, level=None):
        """fine <- (I - A0^-1 gamma D) P_H coarse   (transfer.py:246-259).  ``coarse`` / ``fine`` are numpy arrays or
        Firedrake Functions (what the TransferManager passes, solver.py:593-596): those are read / written through
        ``.dat.data_ro`` / ``.dat.data`` like the reference does (transfer.py:256)."""
        c, f = _values(coarse), _buffer(fine)
        level = self._level_of(f) if level is None else level
        self._ensure(level)
        self.backend.prolong(level, c, f)
        _store(fine, f)

    def restrict(self, fine, coarse, level=None):
        """coarse <- P_H^T (I - gamma D A0^-1) fine   (transfer.py:261-275)."""
        f, c = _values(fine), _buffer(coarse)
        level = self._level_of(f) if level is None else level
        self._ensure(level)
        self.backend.restrict(level, f, c)
        _store(coarse, c)

    def _level_of(self, fine):
        n = fine.size
        for lev, ndofs in self.backend.level_sizes().items():
            if ndofs == n:
                return lev
        raise KeyError("no level with %d dofs" % n)


class SVSchoeberlTransfer(AutoSchoeberlTransfer):
    """A0 = nu(2 sym grad u, grad v) + gamma(div u, div v);  D = (div u, div v)
    (transfer.py:293-309)."""
    divform = "sv"


class PkP0SchoeberlTransfer(AutoSchoeberlTransfer):
    """Same with cell-averaged divergence (transfer.py:312-332)."""
    divform = "pkp0"


class NullTransfer:
    """Fills the destination with NaN (transfer.py:359-366); used for the pressure space."""

    def transfer(self, src, dest):
        dest[...] = np.nan

    inject = transfer
    prolong = transfer
    restrict = transfer


def device_transfer_backend(levels, values_for=None, device=0, deterministic=False, condense=True):
    """``backend=`` / ``values_for=`` of the transfer classes for a hierarchy handed over as
    :class:`alfi_b200.multigrid.LevelInput` (coarsest first): a device context whose levels hold what
    ``restrict_or_prolong`` needs and nothing else — the operator pattern, the Dirichlet dofs, ``P_H`` with the
    coarse-boundary dofs (`alfib_transfer_set`), the cell patches (with their macro-cell blocks) and, through
    `alfib_transfer_update`, the ``A0`` / ``gamma D`` values.  ``values_for(level, nu, gamma)`` must return the
    re-assembled ``(A0 values, D values)``; None keeps the values the levels carry (no parameter change possible).
    This is what `alfi_b200.firedrake_adapter.transfer_backend` returns for a live solver and what the tests build from
    the synthetic hierarchy: the reference's ``SVSchoeberlTransfer((nu, gamma), tdim, hierarchy)`` line then only
    gains ``**device_transfer_backend(...)``."""
    from .lib import PATCHES_TRANSFER, Context
    ctx = Context(device, deterministic)
    for l, li in enumerate(levels):
        ctx.level_create(l, li.n_nodes, li.bs)
        ctx.set_bsr_pattern(l, li.rowptr, li.colidx)
        ctx.set_bc(l, li.bc_dofs)
        if l > 0:
            cb = li.cb_dofs if li.cb_dofs is not None else np.empty(0, np.int32)
            ctx.set_transfer(l, li.P, cb, li.P_dof_level)
            if li.cell_offsets is not None:
                ctx.set_patches(l, li.cell_offsets, li.cell_dofs, None, np.zeros(li.cell_offsets.size - 1, np.int32),
                                PATCHES_TRANSFER)
                if condense and li.cell_blocks is not None:
                    ctx.set_patch_blocks(l, li.cell_blocks, PATCHES_TRANSFER)
    carried = {l: (li.a0_vals, li.d_vals) for l, li in enumerate(levels) if l > 0 and li.a0_vals is not None}

    def default_values(level, nu, gamma):
        return carried[level]
    return {"backend": ctx, "values_for": values_for or default_values}
