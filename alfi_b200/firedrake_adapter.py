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

__all__ = ["FiredrakeAdapter", "DMPlexView", "attach", "transfer_backend"]


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
    """The view of a Firedrake FunctionSpace the patch builders need (``nnodes``, ``bs``, ``cell_nodes``) plus the
    PetscSection that attaches its nodes to DMPlex points."""

    def __init__(self, V):
        self.V = V
        self.bs = V.value_size
        self.cell_nodes = np.asarray(V.cell_node_list, dtype=np.int64)
        self.nnodes = V.dof_dset.total_size

    @property
    def section(self):
        return self.V.dm.getDefaultSection()


class _Labels:
    """``plex.labels`` of the synthetic DMPlex look-alike (name -> int array over the chart, -1 = unlabelled) read
    lazily from a petsc4py DMPlex with ``getLabelValue`` (the call alfi itself uses: relaxation.py:34-50,
    transfer.py:36-38,132)."""

    def __init__(self, dm, npoints):
        self.dm, self.npoints, self._cache = dm, npoints, {}

    def get(self, name, default=None):
        if name not in self._cache:
            has = self.dm.hasLabel(name) if hasattr(self.dm, "hasLabel") else True
            if not has:
                self._cache[name] = None
            else:
                arr = np.fromiter((self.dm.getLabelValue(name, p) for p in range(self.npoints)), dtype=np.int64,
                                  count=self.npoints)
                self._cache[name] = arr if (arr != -1).any() else None
        out = self._cache[name]
        return default if out is None else out

    def __getitem__(self, name):
        out = self.get(name)
        if out is None:
            raise KeyError(name)
        return out

    def __contains__(self, name):
        return self.get(name) is not None


class DMPlexView:
    """What the vectorised builders (`star_points`, `macro_star_points`, `patch_dofs_from_points`,
    `macro_interior_blocks`, `coarse_cell_points`, `fix_coarse_boundaries`) read from a mesh, built ONCE from a
    petsc4py ``DMPlex`` through the calls alfi itself makes (``getChart``, ``getCone``, ``getDepthStratum`` /
    ``getHeightStratum``, ``getLabelValue``, ``getTransitiveClosure``): the cone relation as CSR, its transitive
    closure / star as sparse boolean matrices, the strata bounds and the labels.  Everything else
    (``getTransitiveClosure`` for the per-entity callbacks of `alfi_b200.relaxation`, ...) is forwarded to the DM.
    Tested against a petsc4py-shaped stand-in in tests/test_firedrake_adapter.py (petsc4py is absent here)."""

    def __init__(self, dm):
        import scipy.sparse as sp
        self.dm = dm
        pStart, pEnd = dm.getChart()
        if pStart != 0:
            raise ValueError("DMPlex chart must start at 0")
        n = self.npoints = int(pEnd)
        self.dim = int(dm.getDimension())
        self.cStart, self.cEnd = (int(v) for v in dm.getHeightStratum(0))
        self.vStart, self.vEnd = (int(v) for v in dm.getDepthStratum(0))
        rows, cols = [], []
        for p in range(n):
            cone = np.asarray(dm.getCone(p), dtype=np.int64)
            if cone.size:
                rows.append(np.full(cone.size, p, dtype=np.int64))
                cols.append(cone)
        rows = np.concatenate(rows) if rows else np.empty(0, np.int64)
        cols = np.concatenate(cols) if cols else np.empty(0, np.int64)
        self.cone = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n, n))
        self.cone.sort_indices()
        eye = sp.identity(n, dtype=np.int32, format="csr")
        D = self.cone.astype(np.int32)
        C = eye + D
        for _ in range(self.dim - 1):
            C = eye + D @ C
        C.data[:] = 1
        C.sort_indices()
        self.closure = C.tocsr()
        self.star = C.T.tocsr()
        self.star.sort_indices()
        self.labels = _Labels(dm, n)

    def __getattr__(self, name):                 # getDepthStratum, getTransitiveClosure, getLabelValue, getSupport, ...
        return getattr(self.dm, name)

    def node_points(self, V):
        """node -> DMPlex point from the section of the space.  Firedrake's sections count NODES (alfi uses
        ``section.getOffset(p) + d`` directly as node numbers, transfer.py:138-144)."""
        section = V.section
        out = np.full(V.nnodes, -1, dtype=np.int64)
        for p in range(self.npoints):
            dof = section.getDof(p)
            if dof:
                off = section.getOffset(p)
                out[off:off + dof] = p
        if (out < 0).any():
            raise ValueError("the section does not attach every node to a mesh point")
        return out

    def point_coords(self, p):
        """Mean of the vertex coordinates in the closure of p, read as alfi/relaxation.py:61-67 does."""
        dm = self.dm
        dim = dm.getCoordinateDM().getDimension()
        return np.asarray(dm.getVecClosure(dm.getCoordinateSection(), dm.getCoordinatesLocal(), p)).reshape(-1, dim).mean(axis=0)


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

    def plex(self, pc):
        """The DMPlex of the PC's level as a `DMPlexView` (cone / closure / star relations read once per mesh)."""
        dm = pc.getDM()
        view = dm.getAttr("alfi_b200_plex_view") if hasattr(dm, "getAttr") else None
        if view is None:
            view = DMPlexView(dm)
            if hasattr(dm, "setAttr"):
                dm.setAttr("alfi_b200_plex_view", view)
        return view

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
