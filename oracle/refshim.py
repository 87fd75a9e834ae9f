"""Run the REFERENCE'S OWN Python for the index-set rows of the hot path on top of stand-ins.

Test infrastructure (see oracle/__init__.py): only tests/ and the fixture generator
tests/golden/make_reference_golden.py use it; nothing in alfi_b200/ imports it.

alfi's modules start with ``from firedrake import *`` and Firedrake / petsc4py are not installed, but
the parts of alfi that decide *which mesh points form a patch and in which order patches are visited*
are plain Python over a handful of DMPlex / PETSc.IS / PETSc.Options calls:

* alfi/relaxation.py:8-19,21-150   ``select_entity``, ``OrderedRelaxation.__call__`` (+ ``keyfuncs``)
* alfi/relaxation.py:153-177        ``Star.callback``, ``MacroStar.callback``
* alfi/transfer.py:13-46,49-88     ``CoarseCellPatches.__call__``, ``CoarseCellMacroPatches.__call__``
* alfi/transfer.py:121-158         ``AutoSchoeberlTransfer.fix_coarse_boundaries``

`reference_modules()` loads those two source files *from where they lie under /root/reference* (never
copied) with stub ``firedrake`` / ``firedrake.petsc`` / ``pyop2`` / ``matplotlib`` modules in
``sys.modules`` that provide exactly the names the code above touches; the DMPlex is our
:class:`alfi_b200.synth.plex.SynthPlex`.  What this pins is the reference's patch-construction logic
(rows P1-P3, T1, T2 of SURVEY §8a) executed verbatim; what it cannot pin is DMPlex's own semantics
(closure/star/strata), which SynthPlex restates, and anything numerical (PETSc's arithmetic).
The outputs are stored as fixtures (tests/golden/reference_index_sets.npz) because /root/reference does
not exist on the GPU box.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE = os.environ.get("ALFI_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE, "alfi", "relaxation.py"))


# ------------------------------------------------------------------------------------------ PETSc stand-ins
class IS:
    """petsc4py.PETSc.IS as far as relaxation.py:129,142,149 / transfer.py:40-45 use it."""

    def __init__(self):
        self.indices = np.empty(0, dtype=np.int32)

    def createGeneral(self, indices, comm=None):
        self.indices = np.asarray(list(indices), dtype=np.int32)
        return self

    def createStride(self, size, first=0, step=1, comm=None):
        self.indices = (first + step * np.arange(size)).astype(np.int32)
        return self

    def getIndices(self):
        return self.indices

    def getSize(self):
        return self.indices.size


class Options:
    """petsc4py.PETSc.Options(prefix): reads a process-wide table set through `set_options`."""
    table: dict = {}

    def __init__(self, prefix=None):
        self.prefix = prefix or ""

    def _get(self, name, default):
        return Options.table.get(self.prefix + name, default)

    def getInt(self, name, default=None):
        v = self._get(name, default)
        return v if v is default else int(v)

    def getString(self, name, default=None):
        v = self._get(name, default)
        return v if v is default else str(v)


class _Section:
    """PetscSection of a function space: getDof / getOffset of every plex point.  Firedrake numbers the nodes
    attached to one point consecutively and the section offset is the first of them; the synthetic spaces
    number nodes by first encounter, which also keeps the nodes of a point together (checked)."""

    def __init__(self, plex, V):
        pts = plex.node_points(V)                      # point of every node
        self.count = np.bincount(pts, minlength=plex.npoints)
        self.first = np.full(plex.npoints, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(self.first, pts, np.arange(pts.size))
        last = np.full(plex.npoints, -1, dtype=np.int64)
        np.maximum.at(last, pts, np.arange(pts.size))
        has = self.count > 0
        assert (last[has] - self.first[has] + 1 == self.count[has]).all(), "nodes of a point are not consecutive"

    def getDof(self, p):
        return int(self.count[p])

    def getOffset(self, p):
        return int(self.first[p])


class _CoordPlex:
    """Adds the coordinate queries of relaxation.py:59-65 to a SynthPlex (delegates everything else)."""

    def __init__(self, plex):
        self._plex = plex

    def __getattr__(self, name):
        return getattr(self._plex, name)

    def getTransitiveClosure(self, p, useCone=True):
        return self._plex.getTransitiveClosure(p, useCone)

    def getCoordinateSection(self):
        return None

    def getCoordinateDM(self):
        return types.SimpleNamespace(getDimension=lambda: self._plex.dim)

    def getCoordinatesLocal(self):
        return None

    def getVecClosure(self, section, vec, p):
        # coordinates of the vertices in the closure of p, flattened (DMPlexVecGetClosure on the coordinate vector)
        pts = self._plex.getTransitiveClosure(p, True)[0]
        v0, v1 = self._plex.getDepthStratum(0)
        vs = [q for q in pts if v0 <= q < v1]
        return np.concatenate([self._plex.mesh.coords[q - v0] for q in vs])


class FakePC:
    """petsc4py.PC as the patch constructors see it (relaxation.py:110-113, transfer.py:17-19)."""

    def __init__(self, dm, prefix="", ctx=None):
        self._dm, self._prefix, self._ctx = dm, prefix, ctx

    def getDM(self):
        return self._dm

    def getOptionsPrefix(self):
        return self._prefix

    def getAttr(self, name):
        assert name == "ctx"
        return self._ctx


def set_options(table: dict):
    Options.table = dict(table)


# ------------------------------------------------------------------------------------------ Firedrake stand-ins
class _FakeMesh:
    """firedrake mesh of one level of our hierarchy: what transfer.py:21-28,57-64,122-123 read."""

    def __init__(self, hierarchy, level):
        self._hierarchy, self._level = hierarchy, level
        self._topology_dm = _CoordPlex(hierarchy.levels[level].plex)
        self._cell_numbering = None

    def topological_dimension(self):
        return self._hierarchy.levels[self._level].mesh.dim


class FakeHierarchy:
    """firedrake MeshHierarchy / alfi.bary.BaryMeshHierarchy: indexable, with coarse_to_fine_cells."""

    def __init__(self, levels):
        self.levels = levels
        self.coarse_to_fine_cells = [l.c2f for l in levels[:-1]]
        self.meshes = [_FakeMesh(self, i) for i in range(len(levels))]

    def __getitem__(self, i):
        return self.meshes[i]


class FakeFunctionSpace:
    """V of transfer.py:121-158: mesh(), dm.getDefaultSection(), ufl_element().value_shape()."""

    def __init__(self, hierarchy, level, V):
        self._mesh = hierarchy[level]
        self.V = V
        self.section = _Section(hierarchy.levels[level].plex, V)
        self.dm = types.SimpleNamespace(getDefaultSection=lambda: self.section)

    def mesh(self):
        return self._mesh

    def ufl_element(self):
        return types.SimpleNamespace(value_shape=lambda: (self.V.bs,))


def _stub_modules(extra_firedrake=None):
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    class DirichletBC:                                   # transfer.py:146-153 subclasses it
        def __init__(self, V, g, sub_domain):
            self.V, self.g, self.sub_domain = V, g, sub_domain

        def apply(self, fn):                             # transfer.py:266: impose g (= 0) on the bc nodes
            fn.dat.data[np.asarray(self.nodes, dtype=np.int64)] = self.g

    class _cached_property:                              # firedrake.utils.cached_property
        def __init__(self, fn):
            self.fn = fn

        def __get__(self, obj, cls):
            return self.fn(obj)

    def get_level(mesh):                                 # firedrake.mg.utils.get_level
        return mesh._hierarchy, mesh._level

    def get_entity_renumbering(dm, numbering, kind):     # firedrake.cython.mgimpl: identity in the synthetic numbering
        n = dm.getHeightStratum(0)[1]
        ident = np.arange(n)
        return ident, ident

    def timed_function(name):                            # pyop2.profiling.timed_function
        return lambda fn: fn

    class PCBase:                                        # firedrake.PCBase: solver.py:15 derives DGMassInv from it
        pass

    class DistributedMeshOverlapType:                    # solver.py:604-605,661-662
        VERTEX, FACET, NONE = "VERTEX", "FACET", "NONE"

    petsc = types.SimpleNamespace(IS=IS, Options=Options, COMM_SELF=object())
    utils = types.SimpleNamespace(cached_property=_cached_property)
    ufl = types.SimpleNamespace(zero=lambda shape: np.zeros(shape))
    mg_utils = mod("firedrake.mg.utils", get_level=get_level)
    mods = {
        "firedrake": mod("firedrake", DirichletBC=DirichletBC, utils=utils, ufl=ufl, PCBase=PCBase,
                         DistributedMeshOverlapType=DistributedMeshOverlapType, parameters={},
                         **(extra_firedrake or {})),
        "firedrake.petsc": mod("firedrake.petsc", PETSc=petsc),
        "firedrake.dmhooks": mod("firedrake.dmhooks", get_appctx=lambda dm: None),
        "firedrake.mg": mod("firedrake.mg", utils=mg_utils),
        "firedrake.mg.utils": mg_utils,
        "firedrake.cython": mod("firedrake.cython"),
        "firedrake.cython.mgimpl": mod("firedrake.cython.mgimpl", get_entity_renumbering=get_entity_renumbering),
        "pyop2": mod("pyop2"),
        "pyop2.datatypes": mod("pyop2.datatypes", IntType=np.int32),
        "pyop2.profiling": mod("pyop2.profiling", timed_function=timed_function),
        "alfi": mod("alfi", __path__=[]),
        "alfi.stabilisation": mod("alfi.stabilisation"),
        "alfi.bubble": mod("alfi.bubble", BubbleTransfer=type("BubbleTransfer", (), {})),
    }
    try:
        import mpi4py  # noqa: F401
    except ImportError:
        mods["mpi4py"] = mod("mpi4py", MPI=types.SimpleNamespace(), __path__=[])
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        mods["matplotlib"] = mod("matplotlib", __path__=[])
        mods["matplotlib.pyplot"] = mod("matplotlib.pyplot")
    return mods


@contextlib.contextmanager
def reference_modules(with_solver=False, extra_firedrake=None, extra_modules=None):
    """Context manager yielding (relaxation, transfer[, solver]): the reference's modules loaded from their
    source files."""
    if not available():
        raise FileNotFoundError("reference tree not found at %s" % REFERENCE)
    saved = {}
    stubs = _stub_modules(extra_firedrake)
    stubs.update(extra_modules or {})
    for name, m in stubs.items():
        saved[name] = sys.modules.get(name)
        sys.modules[name] = m
    try:
        out = []
        for name in ("relaxation", "transfer") + (("solver",) if with_solver else ()):
            path = os.path.join(REFERENCE, "alfi", name + ".py")
            spec = importlib.util.spec_from_file_location("_alfi_reference_" + name, path)
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
            out.append(module)
            if name == "transfer":                       # solver.py:9 does `from alfi.transfer import *`
                saved.setdefault("alfi.transfer", sys.modules.get("alfi.transfer"))
                sys.modules["alfi.transfer"] = module
        yield tuple(out)
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


def coord_plex(plex):
    return _CoordPlex(plex)


def reference_solver_parameters(solver="ScottVogeliusSolver", tdim=3, patch="macro", solver_type="almg",
                                patch_composition="additive", smoothing=None, relaxation_direction="0+:1-",
                                use_mkl=False, high_accuracy=False, comm_size=1):
    """The option dictionary the reference hands to PETSc: `<solver>.get_parameters()` (alfi/solver.py:305-510,
    with `configure_patch_solver` :599-602 / :655-659) executed on an object that carries just the attributes
    the method reads.  Returns (outer parameters, firedrake `parameters` side effects)."""
    with reference_modules(with_solver=True) as (_, _, sol):
        cls = getattr(sol, solver)
        me = types.SimpleNamespace(
            patch_composition=patch_composition, smoothing=smoothing, tdim=tdim, patch=patch, use_mkl=use_mkl,
            solver_type=solver_type, high_accuracy=high_accuracy,
            problem=types.SimpleNamespace(relaxation_direction=lambda: relaxation_direction),
            mesh=types.SimpleNamespace(mpi_comm=lambda: types.SimpleNamespace(size=comm_size)))
        me.configure_patch_solver = lambda opts: cls.configure_patch_solver(me, opts)
        outer = cls.get_parameters(me)
        side = dict(sol.parameters)
        return outer, side, me.smoothing


def flatten_options(params, prefix=""):
    """Nested solver_parameters -> flat PETSc option names, the way Firedrake's OptionsManager does it
    (nested dicts contribute their key + "_" as a prefix)."""
    out = {}
    for key, val in params.items():
        if isinstance(val, dict):
            out.update(flatten_options(val, prefix + key + "_"))
        else:
            out[prefix + key] = val
    return out


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m
