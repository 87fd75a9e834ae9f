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

__all__ = ["FiredrakeAdapter", "FiredrakeAccess", "DMPlexView", "attach", "transfer_backend"]


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

    def __init__(self, V, plex_to_firedrake_cells=None):
        self.V = V
        self.bs = V.value_size
        # rows in DMPlex cell order when the renumbering is given (the coarse-to-fine cell tables and the patch-cell
        # lists of the builders are in plex numbering; Firedrake numbers its cells differently: transfer.py:25-31)
        cells = np.asarray(V.cell_node_list, dtype=np.int64)
        self.cell_nodes = cells if plex_to_firedrake_cells is None else cells[np.asarray(plex_to_firedrake_cells)]
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


def c2f_to_plex(c2f, firedrake_to_plex_coarse, firedrake_to_plex_fine):
    """A coarse-to-fine cell table in Firedrake cell numbers (row i = coarse Firedrake cell i) -> the same table in
    DMPlex cell numbers with row r = coarse plex cell r (cells are the first stratum of an interpolated DMPlex)."""
    c2f = np.asarray(c2f)
    f2p_c = np.asarray(firedrake_to_plex_coarse)[:c2f.shape[0]]
    out = np.empty_like(c2f)
    out[f2p_c] = np.asarray(firedrake_to_plex_fine)[c2f]
    return out


class FiredrakeAccess:
    """Every Firedrake call behind the coarse-grained hand-over (`alfi_b200.level_builder` protocol) for a live alfi
    solver (alfi/solver.py NavierStokesSolver: ``mh``, ``Z``, ``nu``, ``gamma``, ``bcs``, ``smoothing``, ``hierarchy``).
    Serial meshes (one rank); kept to one Firedrake idiom per method so that each can be checked against the lines of
    the reference it copies.  NOT executed here (no Firedrake in the image): the builders that consume it are tested
    over a stand-in with the same methods (tests/test_level_builder.py)."""

    dof_level_transfer = False

    def __init__(self, solver):
        self.solver = solver
        self._V, self._plex, self._renum, self._forms = {}, {}, {}, {}
        V = solver.Z.sub(0)
        # PkP0SchoeberlTransfer.standard_transfer switches to BubbleTransfer for 3-D [P1+FB]^3 (transfer.py:334-356)
        self.bubble = V.ufl_element().value_shape()[0] == 3 and "CG1" in V.ufl_element().shortstr()
        self.dof_level_transfer = bool(self.bubble)
        self._bubbles = {}

    def nlevels(self):
        return len(self.solver.mh)

    def function_space(self, l):
        from firedrake import FunctionSpace
        if l not in self._V:
            self._V[l] = FunctionSpace(self.solver.mh[l], self.solver.Z.sub(0).ufl_element())
        return self._V[l]

    def _cells(self, l):
        """(plex -> firedrake, firedrake -> plex) cell numbers of level l (transfer.py:25-26)."""
        from firedrake.cython.mgimpl import get_entity_renumbering
        if l not in self._renum:
            mesh = self.solver.mh[l]
            self._renum[l] = get_entity_renumbering(mesh._topology_dm, mesh._cell_numbering, "cell")
        return self._renum[l]

    def space(self, l):
        return _Space(self.function_space(l), self._cells(l)[0])

    def plex(self, l):
        if l not in self._plex:
            self._plex[l] = DMPlexView(self.solver.mh[l]._topology_dm)
        return self._plex[l]

    def coarse_to_fine_cells(self, l):
        """HierarchyBase.coarse_to_fine_cells[l] (firedrake numbers, rows = coarse firedrake cells) in plex numbering,
        rows in coarse plex order — CoarseCellMacroPatches keeps the coarse cells whose PLEX number is a multiple of
        d + 1 (transfer.py:60-71), which `alfi_b200.transfer.coarse_cell_points` does with ``c2f[::d + 1]``."""
        from firedrake.mg.utils import get_level
        hierarchy, _ = get_level(self.solver.mh[l])
        return c2f_to_plex(hierarchy.coarse_to_fine_cells[l], self._cells(l)[1], self._cells(l + 1)[1])

    def bc_nodes(self, l):
        """Homogeneous Dirichlet nodes of the velocity: the solver's bcs re-applied on the level's space, as Firedrake's
        coarsening of the problem does (`firedrake.mg.ufl_utils.coarsen` of a DirichletBC keeps ``sub_domain``)."""
        import ufl
        from firedrake import DirichletBC
        V = self.function_space(l)
        zero = ufl.zero(V.ufl_element().value_shape())                 # only .nodes is read (as transfer.py:155)
        nodes = [np.asarray(DirichletBC(V, zero, bc.sub_domain).nodes)
                 for bc in self.solver.bcs if bc.function_space().index == 0]
        return np.unique(np.concatenate(nodes)) if nodes else np.empty(0, np.int32)

    def _blocks(self, form, V, bcs=None):
        from firedrake import assemble
        mat = assemble(form, bcs=bcs, mat_type="baij").petscmat
        indptr, indices, data = mat.getValuesCSR()
        rowptr, colidx, vals, _ = baij_csr_to_blocks(np.asarray(indptr), np.asarray(indices), np.asarray(data), mat.getBlockSize())
        return rowptr, colidx, vals

    def operator_blocks(self, l):
        """The velocity block of the Jacobian rediscretised on level l: the (0, 0) split of the coarsened
        SNES context, i.e. what PCMG's level KSPs get from Firedrake's dmhooks (``get_appctx(dm).J``)."""
        from firedrake.mg.ufl_utils import coarsen
        ctx = self.solver.solver._ctx                                  # _SNESContext of the NonlinearVariationalSolver
        for _ in range(self.nlevels() - 1 - l):
            ctx = coarsen(ctx, coarsen)
        ctx0, = ctx.split([(0,)])
        return self._blocks(ctx0.J, self.function_space(l), bcs=ctx0._problem.bcs)     # identity rows on the Dirichlet dofs

    def transfer_blocks(self, l, nu, gamma):
        """A0 and gamma D of alfi/transfer.py:293-332 (no bcs: the library keeps the coarse-boundary dofs itself)."""
        from firedrake import Constant, TestFunction, TrialFunction, cell_avg, div, dx, grad, inner, sym
        V = self.function_space(l)
        u, v = TrialFunction(V), TestFunction(V)
        if self.solver.hierarchy == "bary":                 # SVSchoeberlTransfer
            d = inner(div(u), div(v)) * dx
        else:                                               # PkP0SchoeberlTransfer
            d = inner(cell_avg(div(u)), div(v)) * dx(metadata={"mode": "vanilla"})
        a0 = Constant(nu) * inner(2 * sym(grad(u)), grad(v)) * dx + Constant(gamma) * d
        return self._blocks(a0, V)[2], self._blocks(Constant(gamma) * d, V)[2]

    def prolong(self, l, coarse):
        """The standard transfer of the reference applied to an array: firedrake.prolong (transfer.py:284-290), or
        BubbleTransfer for 3-D [P1+FB]^3 (transfer.py:334-356).  Nodal arrays probe component 0 (P = P_node x I)."""
        from firedrake import Function, prolong
        uc, uf = Function(self.function_space(l - 1)), Function(self.function_space(l))
        if self.dof_level_transfer:
            uc.dat.data[...] = np.asarray(coarse).reshape(uc.dat.data_ro.shape)
        else:
            uc.dat.data[...] = 0.0
            uc.dat.data[:, 0] = coarse
        if self.bubble:
            from alfi.bubble import BubbleTransfer
            if l not in self._bubbles:
                self._bubbles[l] = BubbleTransfer(uc.function_space(), uf.function_space())
            self._bubbles[l].prolong(uc, uf)
        else:
            prolong(uc, uf)
        return uf.dat.data_ro.reshape(-1).copy() if self.dof_level_transfer else uf.dat.data_ro[:, 0].copy()

    def parameters(self):
        return float(self.solver.nu), float(self.solver.gamma)

    def pressure_operators(self):
        """(B, M_p^-1): the assembled (div u, q) block and the inverse of the DG pressure mass matrix that
        DGMassInv.initialize assembles (solver.py:21-31)."""
        import scipy.sparse as sp
        from firedrake import FunctionSpace, Tensor, TestFunction, TrialFunction, assemble, div, dx, inner
        fine = self.solver.mh[-1]
        V = self.function_space(self.nlevels() - 1)
        Q = FunctionSpace(fine, self.solver.Z.sub(1).ufl_element())
        u, q = TrialFunction(V), TestFunction(Q)
        ip, ix, dv = assemble(-div(u) * q * dx, mat_type="aij").petscmat.getValuesCSR()     # the (1, 0) block of the residual's Jacobian
        B = sp.csr_matrix((dv, ix, ip), shape=(Q.dim(), V.dim()))
        p = TrialFunction(Q)
        ip, ix, dv = assemble(Tensor(inner(p, q) * dx).inv).petscmat.getValuesCSR()          # solver.py:24
        return B, sp.csr_matrix((dv, ix, ip), shape=(Q.dim(), Q.dim()))


class FiredrakeAdapter(HostAdapter):
    """``access`` (a `FiredrakeAccess`, or any object with its methods) is what the coarse-grained PCs need on top of
    the per-level methods: `levels(pc)` / `parameters(pc)` / ``smoothing`` / ``restriction`` for
    `alfi_b200.VelocityMGPC`, `pressure_operators(pc)` for `alfi_b200.ALFieldsplitPC`.  ``fieldsplit_0`` is the
    reference's own dictionary (solver.py:359-379), read by `alfi_b200.pc.fieldsplit0_config`."""

    def __init__(self, device=0, deterministic=False, access=None, fieldsplit_0=None, hierarchy="bary",
                 macro_expand="all", restriction=True):
        self.device, self.deterministic = device, deterministic
        self.access, self.hierarchy, self.macro_expand, self.restriction = access, hierarchy, macro_expand, restriction
        self.smoothing = None
        self._construct, self._sort_order, self._composition = "star", None, "additive"
        self._prolongations = {}
        if fieldsplit_0 is not None:
            from .pc import fieldsplit0_config
            cfg = fieldsplit0_config(fieldsplit_0)
            self.smoothing = cfg["smoothing"]
            self._construct, self._sort_order = cfg["construct"], cfg["sort_order"]
            self._composition = cfg["local_type"]

    def levels(self, pc):
        """`LevelInput` per level, coarsest first, for `alfi_b200.VelocityMGPC` (alfi_b200.level_builder)."""
        from .level_builder import build_level_inputs
        if self.access is None:
            raise RuntimeError("FiredrakeAdapter.levels needs an access object: attach(solver) provides it")
        return build_level_inputs(self.access, construct=self._construct, sort_order=self._sort_order,
                                  macro_expand=self.macro_expand, bary=self.hierarchy == "bary",
                                  composition=self._composition, prolongations=self._prolongations)

    def parameters(self, pc):
        return self.access.parameters()

    def pressure_operators(self, pc):
        B, Minv = self.access.pressure_operators()
        bs = self.access.space(self.access.nlevels() - 1).bs
        bc = np.asarray(self.access.bc_nodes(self.access.nlevels() - 1), dtype=np.int64)
        return B, Minv, (bc[:, None] * bs + np.arange(bs)[None, :]).ravel()

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


def attach(solver, device=None, fieldsplit_0=None):
    """Put an adapter on every level's DM so that `alfi_b200.PatchPC` / `VelocityMGPC` / `ALFieldsplitPC` find it
    (INTEGRATION.md §1).  ``fieldsplit_0``: the reference's dictionary for the coarse-grained PCs (smoothing, patch
    construction); the per-level `PatchPC` reads its options from PETSc and does not need it."""
    rank = solver.mesh.comm.rank
    ad = FiredrakeAdapter(device=rank % 8 if device is None else device, access=FiredrakeAccess(solver),
                          fieldsplit_0=fieldsplit_0, hierarchy=solver.hierarchy,
                          restriction=getattr(solver, "restriction", True))
    if ad.smoothing is None:
        ad.smoothing = getattr(solver, "smoothing", None) or (10 if solver.tdim > 2 else 6)      # solver.py:309-310
    for mesh in solver.mh:
        mesh._topology_dm.setAttr("alfi_b200_adapter", ad)
    return ad


def transfer_backend(solver, access=None, make_backend=None, device=0):
    """kwargs for `alfi_b200.SVSchoeberlTransfer(..., **transfer_backend(solver))`: a device context
    whose levels hold what ``restrict_or_prolong`` needs — operator pattern, Dirichlet dofs, the standard prolongation
    (probed once per mesh), the coarse-boundary dofs and the cell patches — and the callback that re-assembles the
    (A0, gamma D) values when (nu, gamma) change (alfi/transfer.py:173-184, 238-244).  ``access`` defaults to
    `FiredrakeAccess(solver)`; ``make_backend`` to `alfi_b200.transfer.device_transfer_backend` (tests replace both)."""
    from .level_builder import build_level_inputs
    from .transfer import device_transfer_backend
    access = FiredrakeAccess(solver) if access is None else access
    levels = build_level_inputs(access, bary=getattr(solver, "hierarchy", "bary") == "bary", smoother=False)

    def values_for(level, nu, gamma):
        return access.transfer_blocks(level, float(nu), float(gamma))
    return (make_backend or device_transfer_backend)(levels, values_for=values_for, device=device)
