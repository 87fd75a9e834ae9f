"""Patch constructors for the additive-Schwarz smoother — drop-in for ``alfi.relaxation``.

Same public surface as the reference (alfi/relaxation.py): ``Star`` and ``MacroStar`` are
callables ``obj(pc) -> (patches, iterationSet)`` selected by dotted name through
``patch_pc_patch_construct_python_type`` (alfi/solver.py:334,341), reading
``pc_patch_construction_<Name>_{dim,codim,sort_order}`` from the PC's options prefix
(alfi/relaxation.py:77-108).  They talk to the DM through the DMPlex methods the reference
uses (``getDepthStratum``, ``getHeightStratum``, ``getTransitiveClosure``, ``getLabelValue``),
so they run against a petsc4py DMPlex or against :class:`alfi_b200.synth.plex.SynthPlex`.

``patches`` are returned as ``PETSc.IS`` objects when petsc4py is importable, otherwise as
int32 numpy arrays (the only property PCPATCH uses is the index list).

For large synthetic meshes ``star_points``/``macro_star_points`` build the same point sets for
*all* vertices at once from the sparse star/closure relations; tests check they agree with the
per-entity callbacks bit for bit.
"""
from __future__ import annotations

from functools import partial

import numpy as np
import scipy.sparse as sp

try:                                    # pragma: no cover - petsc4py is absent in this image
    from petsc4py import PETSc
except Exception:                       # noqa: BLE001
    PETSc = None

__all__ = ["select_entity", "OrderedRelaxation", "Star", "MacroStar",
           "star_points", "macro_star_points", "parse_sort_order", "iteration_order"]


def select_entity(p, dm=None, exclude=None):
    """True unless label ``exclude`` marks point p (alfi/relaxation.py:8-19)."""
    if exclude is None:
        return True
    return dm.getLabelValue(exclude, p) == -1


def _make_is(indices):
    idx = np.asarray(indices, dtype=np.int32)
    if PETSc is not None:               # pragma: no cover
        return PETSc.IS().createGeneral(idx, comm=PETSc.COMM_SELF)
    return idx


class _Options:
    """Tiny stand-in for ``PETSc.Options(prefix)`` over a plain dict."""

    def __init__(self, prefix, table):
        self.prefix, self.table = prefix or "", table or {}

    def _get(self, name, default):
        return self.table.get(self.prefix + name, self.table.get(name, default))

    def getInt(self, name, default=None):
        v = self._get(name, default)
        return v if v is default else int(v)

    def getString(self, name, default=None):
        v = self._get(name, default)
        return v if v is default else (None if v is None else str(v))


def _options_for(pc):
    prefix = pc.getOptionsPrefix()
    if PETSc is not None and not hasattr(pc, "options"):    # pragma: no cover
        return PETSc.Options(prefix)
    return _Options(prefix, getattr(pc, "options", {}))


def parse_sort_order(sortorders):
    """``"0+:1-|1+"`` → [[(0, +1), (1, -1)], [(1, +1)]]  (alfi/relaxation.py:88-108)."""
    if sortorders is None or sortorders in ("None", ""):
        return None
    res = []
    for sortorder in sortorders.split("|"):
        sortdata = []
        for axis in sortorder.split(":"):
            ax = int(axis[0])
            sgn = {"+": 1, "-": -1}[axis[1]] if len(axis) > 1 else 1
            sortdata.append((ax, sgn))
        res.append(sortdata)
    return res


def iteration_order(coords, sortorders, literal=True):
    """Concatenated stable sorts of patch indices by signed coordinates
    (alfi/relaxation.py:141-149); identity when no sort order is given.

    ``literal=True`` (default) reproduces what the reference actually computes for several sweeps
    ``"a|b"``: its key functions close over the loop variable ``sortdata`` (relaxation.py:96-107), so when
    they are called every sweep sorts by the keys of the *last* sweep.  Found by running the reference's
    own code (tests/test_reference_code.py); single-sweep orders — all the examples use "0+:1-" — are
    unaffected.  ``literal=False`` gives each sweep its own keys, which is what the syntax suggests."""
    n = len(coords)
    sweeps = parse_sort_order(sortorders)
    if sweeps is None:
        return np.arange(n, dtype=np.int32)
    if literal:
        sweeps = [sweeps[-1]] * len(sweeps)
    coords = np.asarray(coords, dtype=np.float64).reshape(n, -1)
    out = []
    for sortdata in sweeps:
        keys = [sgn * coords[:, ax] for ax, sgn in sortdata]
        out.append(np.lexsort(tuple(reversed(keys))))         # first key is primary; stable
    return np.concatenate(out).astype(np.int32)


class OrderedRelaxation:
    def __init__(self):
        self.name = None

    def callback(self, dm, entity):
        raise NotImplementedError

    def set_options(self, dm, opts, name):
        pass

    @staticmethod
    def star(dm, p):
        return dm.getTransitiveClosure(p, useCone=False)[0]

    @staticmethod
    def closure(dm, p):
        return dm.getTransitiveClosure(p, useCone=True)[0]

    @staticmethod
    def coords(dm, p):
        if hasattr(dm, "point_coords"):
            return dm.point_coords(p)
        sec = dm.getCoordinateSection()                        # pragma: no cover - petsc4py path
        dim = dm.getCoordinateDM().getDimension()
        return dm.getVecClosure(sec, dm.getCoordinatesLocal(), p).reshape(-1, dim).mean(axis=0)

    @staticmethod
    def get_entities(opts, name, dm):
        sentinel = object()
        codim = opts.getInt("pc_patch_construction_%s_codim" % name, default=sentinel)
        if codim is sentinel:
            dim = opts.getInt("pc_patch_construction_%s_dim" % name, default=0)
            return range(*dm.getDepthStratum(dim))
        return range(*dm.getHeightStratum(codim))

    def __call__(self, pc):
        dm = pc.getDM()
        opts = _options_for(pc)
        self.opts = opts
        name = self.name
        assert name is not None
        self.set_options(dm, opts, name)

        select = partial(select_entity, dm=dm, exclude="pyop2_ghost")
        patches, kept = [], []
        for entity in filter(select, self.get_entities(opts, name, dm)):
            sub = self.callback(dm, entity)
            if sub is None:
                continue
            patches.append(_make_is(sub))
            kept.append(entity)
        sortorders = opts.getString("pc_patch_construction_%s_sort_order" % name, default=None)
        if parse_sort_order(sortorders) is None:
            order = np.arange(len(patches), dtype=np.int32)
        else:
            order = iteration_order([self.coords(dm, p) for p in kept], sortorders)
        self.entities = kept
        return patches, _make_is(order)


class Star(OrderedRelaxation):
    """Patch = topological star of an entity (alfi/relaxation.py:153-160)."""

    def __init__(self):
        super().__init__()
        self.name = "Star"

    def callback(self, dm, vertex):
        return list(self.star(dm, vertex))


class MacroStar(OrderedRelaxation):
    """Star of a macro vertex plus the stars of the Alfeld barycentres in its closure
    (alfi/relaxation.py:163-177); non-macro vertices get no patch."""

    def __init__(self):
        super().__init__()
        self.name = "MacroStar"
        self.expand = "all"

    def set_options(self, dm, opts, name):
        # extension (not in the reference): "all" = the reference's literal behaviour,
        # "vertices" = expand barycentre vertices only (the open macro star in 3-D as well)
        self.expand = opts.getString("pc_patch_construction_%s_expand" % name, default="all")
        assert self.expand in ("all", "vertices")

    def callback(self, dm, vertex):
        if dm.getLabelValue("MacroVertices", vertex) != 1:
            return None
        s = list(self.star(dm, vertex))
        closures = []
        for e in s:
            closures.extend(self.closure(dm, e))
        # Literal restatement of relaxation.py:173: the label test runs over *every* point of
        # closure(star(v)), and edges/faces/cells are never labelled, so they pass it too.
        expand = [p for p in closures if dm.getLabelValue("MacroVertices", p) != 1]
        if self.expand == "vertices":
            (vlo, vhi) = dm.getDepthStratum(0)
            expand = [p for p in expand if vlo <= p < vhi]
        their_star = []
        for p in expand:
            their_star.extend(self.star(dm, p))
        return s + their_star


# --------------------------------------------------------------------------- vectorised builders
def star_points(plex, entities=None):
    """CSR boolean (npatch x npoints): row i = star of vertex i (all owned vertices)."""
    vlo, vhi = plex.getDepthStratum(0)
    ents = np.arange(vlo, vhi) if entities is None else np.asarray(entities)
    ghost = plex.labels.get("pyop2_ghost")
    if ghost is not None:
        ents = ents[ghost[ents] == -1]
    H = plex.star[ents]
    H.sort_indices()
    return H.tocsr(), ents


def macro_star_points(plex, expand="all"):
    """CSR boolean (npatch x npoints): MacroStar point sets of all macro vertices.

    Literal semantics of relaxation.py:168-177: every point of closure(star(v)) that is not a
    macro vertex is expanded to its star.  In 2-D this is exactly the open macro star; in 3-D
    the macro edges on the link of v are expanded too, so the patch reaches into the
    neighbouring macro cells around those edges ("MacroStar in 3-D").
    """
    vlo, vhi = plex.getDepthStratum(0)
    mv = plex.labels["MacroVertices"]
    ents = np.flatnonzero(mv[vlo:vhi] == 1) + vlo
    ghost = plex.labels.get("pyop2_ghost")
    if ghost is not None:
        ents = ents[ghost[ents] == -1]
    S = plex.star[ents].astype(np.int32)                       # star(v)
    cl = (S @ plex.closure.astype(np.int32)).tocsr()           # closure(star(v))
    notmacro = (mv != 1).astype(np.int32)
    if expand == "vertices":
        notmacro[:vlo] = 0
        notmacro[vhi:] = 0
    B = (cl @ sp.diags(notmacro, format="csr", dtype=np.int32)).tocsr()   # drop macro vertices
    H = S + B @ plex.star.astype(np.int32)
    H = H.tocsr()
    H.eliminate_zeros()
    H.data[:] = 1
    H.sort_indices()
    return H, ents
