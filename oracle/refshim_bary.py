"""Run the reference's `BaryMeshHierarchy` (alfi/bary.py:29-194) VERBATIM over stand-ins.

Test infrastructure (see oracle/__init__.py).  The function refines a mesh uniformly, Alfeld-splits every level
and — the part that matters to the hot path — composes the coarse-to-fine *cell* maps of the barycentric meshes
from those of the uniform meshes (bary.py:141-171): coarse bary cell c belongs to uniform cell c // (d+1) and is
mapped to all (d+1)·2^d bary cells of that uniform cell's children.  `CoarseCellMacroPatches` (transfer.py:49-88)
and the standard prolongation consume exactly that table.

Stand-ins: `dm.refine()` = alfi_b200.synth.mesh refinement (Kuhn or general red refinement),
`PETSc.DMPlexTransform(REFINEALFELD)` = `alfeld_split`, `firedrake.Mesh(dm)` = a wrapper with identity
cell numbering, `impl.coarse_to_fine_cells` = the synthetic uniform table, `HierarchyBase` = a record of its
arguments.  So the *composition* is the reference's; the refinement itself is ours.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import types

import numpy as np

from . import refshim


class _DM:
    """DMPlex of one uniform or barycentric mesh."""

    def __init__(self, world, kind, level):
        self.world, self.kind, self.level = world, kind, level     # kind: "uniform" | "bary"
        self.labels = {}
        self.refine_level = None

    @property
    def mesh(self):
        lev = self.world.levels[self.level]
        return lev.macro if self.kind == "uniform" else lev.mesh

    def setRefinementUniform(self, flag):
        assert flag is True

    def refine(self):
        assert self.kind == "uniform"
        return self.world.dm("uniform", self.level + 1)

    def removeLabel(self, name):
        self.labels.pop(name, None)

    def getHeightStratum(self, h):
        m = self.mesh
        return (0, m.nc) if h == 0 else (m.nc + m.nv, m.nc + m.nv + m.facets.shape[0])

    def getDepthStratum(self, d):
        m = self.mesh
        assert d == 0
        return (m.nc, m.nc + m.nv)

    def setLabelValue(self, name, p, value):
        self.labels.setdefault(name, {})[p] = value

    def setRefineLevel(self, i):
        self.refine_level = i

    def getComm(self):
        return None


class _Mesh:
    def __init__(self, dm, **kw):
        self._topology_dm, self.kw = dm, kw
        self._cell_numbering = None
        self.initialised = False

    def init(self):
        self.initialised = True


class World:
    """A synthetic hierarchy (alfi_b200.synth.hierarchy levels) presented as the objects bary.py walks."""

    def __init__(self, levels):
        self.levels = levels
        self._dms = {}
        self.alfeld_applied = []

    def dm(self, kind, level):
        if (kind, level) not in self._dms:
            self._dms[(kind, level)] = _DM(self, kind, level)
        return self._dms[(kind, level)]

    def base_mesh(self):
        d = self.levels[0].macro.dim
        m = _Mesh(self.dm("uniform", 0))
        m.comm = types.SimpleNamespace(size=1)
        m._grown_halos = False
        m._distribution_parameters = {"partition": True}
        m.ufl_cell = lambda: types.SimpleNamespace(geometric_dimension=lambda: d)
        m.topological_dimension = lambda: d
        return m

    def namespace(self):
        world = self

        class Transform:                                  # PETSc.DMPlexTransform, bary.py:21-25
            def create(self, comm=None):
                return self

            def setType(self, t):
                assert t == "REFINEALFELD"

            def setDM(self, dm):
                self.dm = dm

            def setUp(self):
                pass

            def apply(self, dm):
                assert dm is self.dm and dm.kind == "uniform"
                # bary.py:18-19: every vertex of the uniform mesh was labelled MacroVertices = 1 beforehand
                nv = dm.mesh.nv
                assert sorted(dm.labels["MacroVertices"]) == list(range(*dm.getDepthStratum(0))) and nv
                world.alfeld_applied.append(dm.level)
                return world.dm("bary", dm.level)

        def ident(dm):
            return np.arange(dm.mesh.nc)
        impl = types.SimpleNamespace(
            filter_labels=lambda dm, stratum, *labels: None,
            create_lgmap=lambda dm: None,
            get_entity_renumbering=lambda dm, numbering, kind: (ident(dm), ident(dm)),
            coarse_to_fine_cells=lambda coarse, fine, cl, fl: (
                world.levels[coarse._topology_dm.level].macro_c2f,
                None))
        petsc = types.SimpleNamespace(DMPlexTransform=Transform, IntType=np.int32,
                                      DMPlexTransformType=types.SimpleNamespace(REFINEALFELD="REFINEALFELD"))

        def HierarchyBase(meshes, c2f, f2c, refinements_per_level, nested=None):
            return types.SimpleNamespace(meshes=meshes, coarse_to_fine_cells=c2f, fine_to_coarse_cells=f2c,
                                         refinements_per_level=refinements_per_level, nested=nested)
        firedrake = dict(Mesh=lambda dm, **kw: _Mesh(dm, **kw), HierarchyBase=HierarchyBase, np=np)
        extra = {
            "firedrake.cython.mgimpl": refshim._module("firedrake.cython.mgimpl", **impl.__dict__),
            "firedrake.cython.dmcommon": refshim._module("firedrake.cython.dmcommon", FACE_SETS_LABEL="Face Sets"),
            "firedrake.petsc": refshim._module("firedrake.petsc", PETSc=petsc),
        }
        return firedrake, extra, impl

    @contextlib.contextmanager
    def reference_function(self):
        firedrake, extra, impl = self.namespace()
        with refshim.reference_modules(extra_firedrake=firedrake, extra_modules=extra):
            import sys
            # `import firedrake` + `firedrake.Mesh`, `from firedrake.cython import mgimpl as impl`
            sys.modules["firedrake.cython"].mgimpl = sys.modules["firedrake.cython.mgimpl"]
            path = os.path.join(refshim.REFERENCE, "alfi", "bary.py")
            spec = importlib.util.spec_from_file_location("_alfi_reference_bary", path)
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
            yield module.BaryMeshHierarchy
