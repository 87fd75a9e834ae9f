"""petsc4py python-PC plugins — the drop-in surface for alfi's solver dictionaries.

Two granularities:

* :class:`PatchPC` replaces ``firedrake.PatchPC`` one to one: alfi selects it by changing the
  single string at alfi/solver.py:319 (``"pc_python_type": "alfi_b200.PatchPC"``).  It honours the
  sibling keys of alfi/solver.py:320-344, 599-602, 655-659 under the ``patch_`` prefix and follows
  the Firedrake ``PCBase`` protocol shown by ``DGMassInv`` (alfi/solver.py:15-38):
  ``initialize(pc)``, ``update(pc)``, ``apply(pc, x, y)``, ``applyTranspose(pc, x, y)``.
* :class:`VelocityMGPC` replaces the whole ``fieldsplit_0`` sub-dictionary
  (alfi/solver.py:359-379) by one python PC that runs richardson(1) + PCMG-full + FGMRES(m)
  smoothing + Schöberl transfers + coarse solve on the device, so one application costs one
  H2D/D2H pair instead of one per smoother application.

PETSc objects are reached only through the handful of methods used below (``getOperators``,
``getDM``, ``getOptionsPrefix``, ``getAttr``; ``Vec.array_r`` / ``Vec.array_w``).  What differs
between a Firedrake deployment and the synthetic stand-in is *how the operator and the function
space are read*; that is isolated in the ``HostAdapter`` the PC finds on the DM
(``dm.getAttr("alfi_b200_adapter")`` / appctx), see INTEGRATION.md.
"""
from __future__ import annotations

import importlib

import numpy as np

from .lib import PATCHES_SMOOTHER, Context
from .multigrid import DeviceMultigrid, LevelInput
from .patches import greedy_colouring, patch_dofs_from_points, points_to_csr, sweep_stages
from .relaxation import _Options, star_points

__all__ = ["fieldsplit0_config", "PatchPC", "VelocityMGPC", "ALFieldsplitPC", "HostAdapter"]


class HostAdapter:
    """How a PC reads its inputs from the host framework.  Subclass per framework.

    `operator(pc)`      -> (rowptr, colidx, vals (nnzb, bs, bs), block_col_major) of ``P``
    `function_space(pc)`-> object with ``nnodes``, ``bs``, ``cell_nodes`` and, for ``plex``,
                           ``node_points`` (the PetscSection of the space)
    `plex(pc)`          -> DMPlex-like object (petsc4py DMPlex or SynthPlex)
    `bc_nodes(pc)`      -> node indices of the global Dirichlet conditions
    `options(pc)`       -> PETSc.Options-like object for the PC's prefix
    `patch_corrections(pc, patches)` (optional) -> None, or (off, rows, cols, vals): what separates PCPATCH's patch
                           operators from sub-matrices of ``P`` when the form has interior-facet integrals (Burman
                           stabilisation; include/alfib.h alfib_level_set_patch_corrections).  Called at every set-up;
                           the pattern must not change.
    """

    def operator(self, pc):
        raise NotImplementedError

    def function_space(self, pc):
        raise NotImplementedError

    def plex(self, pc):
        return pc.getDM()

    def bc_nodes(self, pc):
        raise NotImplementedError

    def options(self, pc):
        return _Options(pc.getOptionsPrefix(), getattr(pc, "options", {}))


def _adapter(pc) -> HostAdapter:
    ad = pc.getAttr("alfi_b200_adapter") if hasattr(pc, "getAttr") else None
    if ad is None:
        dm = pc.getDM()
        ad = dm.getAttr("alfi_b200_adapter") if hasattr(dm, "getAttr") else getattr(dm, "alfi_b200_adapter", None)
    if ad is None:
        raise RuntimeError("no alfi_b200 HostAdapter attached to the PC or its DM (see INTEGRATION.md)")
    return ad


def _truthy(v):
    return str(v).lower() in ("1", "true", "yes", "on") if not isinstance(v, bool) else v


class PatchPC:
    """Additive-Schwarz patch smoother on the GPU (PCPATCH semantics)."""

    _prefix = "patch_"

    def setUp(self, pc):                      # petsc4py calls setUp; Firedrake's PCBase dispatches
        if getattr(self, "initialized", False):
            self.update(pc)
        else:
            self.initialize(pc)
            self.initialized = True

    # -- PCBase protocol ------------------------------------------------------------------------
    def initialize(self, pc):
        ad = _adapter(pc)
        opts = ad.options(pc)
        g = lambda key, default=None: opts.getString(self._prefix + key, default=default)   # noqa: E731
        if _truthy(g("pc_patch_partition_of_unity", "false")):
            raise NotImplementedError("partition_of_unity weighting is off in alfi (solver.py:321)")
        self.local_type = g("pc_patch_local_type", "additive")
        if self.local_type not in ("additive", "multiplicative"):
            raise NotImplementedError("patch composition %r" % self.local_type)
        self.symmetrise = _truthy(g("pc_patch_symmetrise_sweep", "false"))
        if g("sub_pc_type", "lu") != "lu" or g("sub_ksp_type", "preonly") != "preonly":
            raise NotImplementedError("patch sub-solver must be preonly + lu (solver.py:326-327)")
        # accepted and implied by the implementation: save_operators, precompute_element_tensors,
        # sub_mat_type, dense_inverse (always an explicit inverse), factor_mat_solver_type, statistics
        self.options_seen = {k: g(k) for k in ("pc_patch_save_operators", "pc_patch_precompute_element_tensors",
                                                "pc_patch_sub_mat_type", "pc_patch_dense_inverse",
                                                "sub_pc_factor_mat_solver_type", "pc_patch_statistics")}
        plex, V = ad.plex(pc), ad.function_space(pc)
        ctype = g("pc_patch_construct_type", "star")
        if ctype == "star":
            dim = opts.getInt(self._prefix + "pc_patch_construct_dim", default=0)
            if dim != 0:
                raise NotImplementedError("builtin star construction on vertices only (solver.py:337-338)")
            H, _ = star_points(plex)
            order = None
        elif ctype == "python":
            dotted = g("pc_patch_construct_python_type")
            mod, _, cls = dotted.rpartition(".")
            mod = {"alfi": "alfi_b200.relaxation", "alfi.relaxation": "alfi_b200.relaxation",
                   "alfi.transfer": "alfi_b200.transfer"}.get(mod, mod)
            ctor = getattr(importlib.import_module(mod), cls)()
            patches, iterset = ctor(_PrefixedPC(pc, self._prefix))
            H = points_to_csr([np.asarray(getattr(p, "indices", p)) for p in patches], plex.npoints)
            order = np.asarray(getattr(iterset, "indices", iterset), dtype=np.int32)
        else:
            raise NotImplementedError("patch construct_type %r" % ctype)
        self.bc_nodes = np.asarray(ad.bc_nodes(pc), dtype=np.int32)
        self.patches = patch_dofs_from_points(plex, V, H, bc_nodes=self.bc_nodes, order=order)
        greedy_colouring(self.patches, V.nnodes * V.bs)
        rowptr, colidx, vals, colmajor = ad.operator(pc)
        self.ctx = Context(getattr(ad, "device", 0), deterministic=getattr(ad, "deterministic", False))
        self.n = V.nnodes * V.bs
        c = self.ctx
        c.level_create(0, V.nnodes, V.bs)
        c.set_bsr_pattern(0, rowptr, colidx)
        bs = V.bs
        c.set_bc(0, (self.bc_nodes[:, None] * bs + np.arange(bs)[None, :]).ravel())
        ps = self.patches
        c.set_patches(0, ps.offsets, ps.dofs, ps.order, ps.colours, PATCHES_SMOOTHER)
        if self.local_type == "multiplicative":
            # the sequential sweep as a schedule of stages of mutually uncoupled patches (solver.py:322-335)
            self.stages = sweep_stages(ps, rowptr, colidx)
            c.set_sweep_stages(0, self.stages, self.symmetrise, PATCHES_SMOOTHER)
        c.set_bsr_values(0, vals, colmajor)
        self._corrections(pc, first=True)
        c.factor(0)

    def _corrections(self, pc, first=False):
        ad = _adapter(pc)
        corr = ad.patch_corrections(pc, self.patches) if hasattr(ad, "patch_corrections") else None
        if corr is None:
            return
        off, rows, cols, vals = corr
        if first:
            self.ctx.set_patch_corrections(0, off, rows, cols, PATCHES_SMOOTHER)
        self.ctx.set_patch_correction_values(0, vals, PATCHES_SMOOTHER)

    def update(self, pc):
        """Called on every PCSetUp after the first, i.e. once per Newton step: new operator values
        → re-gather and re-invert every patch (PCSetUp_PATCH)."""
        rowptr, colidx, vals, colmajor = _adapter(pc).operator(pc)
        self.ctx.set_bsr_values(0, vals, colmajor)
        self._corrections(pc)
        self.ctx.factor(0)

    def apply(self, pc, x, y):
        xa = np.ascontiguousarray(x.array_r)
        out = np.empty(self.n)
        self.ctx.smoother_apply(0, xa, out)
        y.array_w[:] = out

    def applyTranspose(self, pc, x, y):
        raise NotImplementedError("Sorry!")        # as alfi/solver.py:37-38

    def view(self, pc, viewer=None):
        ps = self.patches
        print("alfi_b200.PatchPC: %d patches, max %d dofs, %d colours, %.1f MB of inverses"
              % (ps.npatch, int(ps.sizes.max()), int(ps.colours.max()) + 1,
                 self.ctx.patch_storage_bytes(0) / 1e6))


def fieldsplit0_config(fs0: dict) -> dict:
    """Read the reference's own ``fieldsplit_0`` dictionary (alfi/solver.py:359-379 with ``mg_levels`` from
    :313-344) and return what `VelocityMGPC` / `DeviceMultigrid` need: ``smoothing`` (FGMRES iterations per
    level), the patch construction (``construct``: "star" or a python class name, ``sort_order``) and the
    patch sub-matrix options.  The dictionary is taken unchanged; anything the device path does not implement
    raises NotImplementedError instead of being silently ignored."""
    def need(d, key, *allowed):
        if d.get(key) not in allowed:
            raise NotImplementedError("fieldsplit_0: %s = %r (supported: %s)" % (key, d.get(key), ", ".join(map(repr, allowed))))
    need(fs0, "ksp_type", "richardson")
    need(fs0, "ksp_max_it", 1)
    need(fs0, "ksp_richardson_self_scale", False, None)
    need(fs0, "pc_type", "mg")
    need(fs0, "pc_mg_type", "full")
    lv = fs0["mg_levels"]
    need(lv, "ksp_type", "fgmres")
    need(lv, "ksp_norm_type", "unpreconditioned")
    need(lv, "ksp_convergence_test", "skip")
    need(lv, "pc_type", "python")
    need(lv, "pc_python_type", "firedrake.PatchPC", "alfi_b200.PatchPC")
    need(lv, "patch_pc_patch_partition_of_unity", False, None)
    need(lv, "patch_pc_patch_local_type", "additive", "multiplicative")
    need(lv, "patch_sub_ksp_type", "preonly")
    need(lv, "patch_sub_pc_type", "lu")
    ctype = lv.get("patch_pc_patch_construct_type", "star")
    if ctype == "star":
        need(lv, "patch_pc_patch_construct_dim", 0, None)
        construct, sort_order = "star", None
    elif ctype == "python":
        construct = lv["patch_pc_patch_construct_python_type"]
        name = construct.rpartition(".")[2]
        sort_order = lv.get("patch_pc_patch_construction_%s_sort_order" % name)
    else:
        raise NotImplementedError("fieldsplit_0: patch construct_type %r" % ctype)
    coarse = fs0.get("mg_coarse_assembled", {})
    if coarse and coarse.get("telescope_pc_type", coarse.get("pc_type")) != "lu":
        raise NotImplementedError("fieldsplit_0: the coarse solve must be a direct LU (solver.py:369-378)")
    return {"smoothing": int(lv["ksp_max_it"]), "construct": construct, "sort_order": sort_order,
            "local_type": lv["patch_pc_patch_local_type"], "symmetrise_sweep": bool(lv.get("patch_pc_patch_symmetrise_sweep", False)),
            "sub_mat_type": lv.get("patch_pc_patch_sub_mat_type"), "dense_inverse": bool(lv.get("patch_pc_patch_dense_inverse", False)),
            "patch_lu": lv.get("patch_sub_pc_factor_mat_solver_type")}


class _PrefixedPC:
    """What PCPATCH hands to a python patch constructor: the PC with the ``patch_`` prefix added
    (so ``pc_patch_construction_<Name>_sort_order`` resolves as in alfi/relaxation.py:80-91)."""

    def __init__(self, pc, prefix):
        self._pc, self._prefix = pc, prefix
        self.options = getattr(pc, "options", {})

    def getOptionsPrefix(self):
        return (self._pc.getOptionsPrefix() or "") + self._prefix

    def __getattr__(self, name):
        return getattr(self._pc, name)


class VelocityMGPC:
    """The whole ``fieldsplit_0`` of alfi/solver.py:359-379 as one python PC.

    The adapter provides ``levels(pc) -> list[LevelInput]`` (coarsest first) and ``smoothing``;
    `update` re-uploads the operator values of every level (rediscretised by the host,
    SURVEY A.6) and re-factors; the transfer operators are rebuilt only when (nu, gamma) change
    (alfi/transfer.py:173-184)."""

    def setUp(self, pc):
        if getattr(self, "initialized", False):
            self.update(pc)
        else:
            self.initialize(pc)
            self.initialized = True

    def initialize(self, pc):
        ad = _adapter(pc)
        levels = ad.levels(pc)
        self.mg = DeviceMultigrid(levels, ad.smoothing, device=getattr(ad, "device", 0),
                                  deterministic=getattr(ad, "deterministic", False),
                                  robust_restrict=getattr(ad, "restriction", True))
        self.n = levels[-1].n_nodes * levels[-1].bs
        self._params = ad.parameters(pc) if hasattr(ad, "parameters") else None

    def update(self, pc):
        ad = _adapter(pc)
        levels = ad.levels(pc)
        self.mg.update_operators(levels)
        params = ad.parameters(pc) if hasattr(ad, "parameters") else None
        if params != self._params:
            self.mg.update_transfers(levels)
            self._params = params

    def apply(self, pc, x, y):
        out = np.empty(self.n)
        self.mg.apply(np.ascontiguousarray(x.array_r), out)
        y.array_w[:] = out

    def applyTranspose(self, pc, x, y):
        raise NotImplementedError("Sorry!")


class ALFieldsplitPC(VelocityMGPC):
    """The whole ``outer_fieldsplit`` preconditioner of alfi/solver.py:405-421 as one python PC: PCFIELDSPLIT schur
    with ``full`` factorisation, ``fieldsplit_0`` = the multigrid cycle above, ``fieldsplit_1`` =
    ``alfi.solver.DGMassInv`` (solver.py:15-38, i.e. -(nu + gamma) M_p^-1) — one H2D/D2H pair per OUTER Krylov
    iteration instead of one per velocity-block application (two per iteration).  Selected by replacing the
    ``"pc_type": "fieldsplit"`` block by ``"pc_type": "python", "pc_python_type": "alfi_b200.ALFieldsplitPC"``;
    the outer KSP (fgmres) stays PETSc's.

    Besides what `VelocityMGPC` needs, the adapter provides ``pressure_operators(pc) -> (B, Minv, bc_dofs)`` (scipy
    sparse: the assembled (div u, q) block and the inverse pressure mass matrix — `DGMassInv.initialize` assembles the
    same, solver.py:21-31 — and the Dirichlet velocity dofs) and ``parameters(pc) -> (nu, gamma)``.  Vectors are the
    monolithic [velocity; pressure] layout of the nest matrix."""

    def initialize(self, pc):
        super().initialize(pc)
        ad = _adapter(pc)
        B, Minv, bc_dofs = ad.pressure_operators(pc)
        import scipy.sparse as sp
        keep = np.ones(B.shape[1])
        keep[np.asarray(bc_dofs, dtype=np.int64)] = 0.0
        Bz = (B.tocsr() @ sp.diags(keep)).tocsr()
        Bz.eliminate_zeros()
        self.mg.ctx.schur_set(Bz, Minv, getattr(ad, "remove_constant_pressure", True))
        self.n_total = self.n + B.shape[0]

    def apply(self, pc, x, y):
        nu, gamma = self._params[:2]
        out = np.empty(self.n_total)
        self.mg.ctx.schur_apply(nu, gamma, np.ascontiguousarray(x.array_r), out)
        y.array_w[:] = out
