"""The Firedrake adapter cannot run here (no firedrake / petsc4py), but its pure index arithmetic
can: BAIJ `getValuesCSR()` -> block CSR + (nnzb, bs, bs) values, exercised through a fake Mat."""
import numpy as np
import pytest
import scipy.sparse as sp

from alfi_b200.firedrake_adapter import FiredrakeAdapter


class FakeBAIJ:
    def __init__(self, A_bsr):
        self.A = A_bsr

    def getBlockSize(self):
        return self.A.blocksize[0]

    def getValuesCSR(self):
        # PETSc returns the scalar CSR of a BAIJ matrix with every block fully stored
        rows = np.repeat(np.arange(self.A.shape[0] // self.A.blocksize[0]), np.diff(self.A.indptr))
        bs = self.A.blocksize[0]
        n = self.A.shape[0]
        indptr = [0]
        indices, data = [], []
        for i in range(n):
            bi, r = divmod(i, bs)
            for k in range(self.A.indptr[bi], self.A.indptr[bi + 1]):
                cj = self.A.indices[k]
                indices.extend(range(cj * bs, cj * bs + bs))
                data.extend(self.A.data[k][r])
            indptr.append(len(indices))
        return np.array(indptr, np.int32), np.array(indices, np.int32), np.array(data)


class FakePC:
    def __init__(self, mat):
        self.mat = mat

    def getOperators(self):
        return None, self.mat


@pytest.mark.parametrize("bs", [2, 3])
def test_baij_csr_to_block_csr(bs):
    rng = np.random.default_rng(bs)
    nb = 9
    pat = (sp.random(nb, nb, density=0.3, random_state=bs) + sp.identity(nb)).tocsr()
    pat.sort_indices()
    vals = rng.standard_normal((pat.indices.size, bs, bs))
    A = sp.bsr_matrix((vals, pat.indices, pat.indptr), shape=(nb * bs, nb * bs))
    rowptr, colidx, v, colmajor = FiredrakeAdapter().operator(FakePC(FakeBAIJ(A)))
    assert not colmajor
    assert np.array_equal(rowptr, pat.indptr) and np.array_equal(colidx, pat.indices)
    assert np.array_equal(v, vals)


# ---- DMPlexView: the vectorised builders over a petsc4py-shaped DMPlex ---------------------------------------
class _FakeSection:
    """PetscSection of a Firedrake function space: dof / offset per mesh point, in NODES."""

    def __init__(self, node_points, npoints):
        order = np.argsort(node_points, kind="stable")
        self.dof = np.bincount(node_points, minlength=npoints)
        self.off = np.full(npoints, -1, dtype=np.int64)
        first = np.flatnonzero(np.r_[True, np.diff(node_points[order]) != 0])
        self.off[node_points[order][first]] = order[first]
        for p in np.flatnonzero(self.dof > 1):       # nodes of one point must be numbered consecutively
            nodes = np.flatnonzero(node_points == p)
            assert nodes.max() - nodes.min() + 1 == nodes.size

    def getDof(self, p):
        return int(self.dof[p])

    def getOffset(self, p):
        return int(self.off[p])

    def getChart(self):
        return 0, self.dof.size


class _FakeDMPlex:
    """Only the petsc4py DMPlex methods alfi and the adapter call, backed by a SynthPlex (none of its attributes)."""

    def __init__(self, plex):
        self._p = plex
        self._attrs = {}

    def getChart(self):
        return self._p.getChart()

    def getDimension(self):
        return self._p.getDimension()

    def getCone(self, p):
        return self._p.getCone(p)

    def getSupport(self, p):
        return self._p.getSupport(p)

    def getDepthStratum(self, d):
        return self._p.getDepthStratum(d)

    def getHeightStratum(self, h):
        return self._p.getHeightStratum(h)

    def getTransitiveClosure(self, p, useCone=True):
        return self._p.getTransitiveClosure(p, useCone)

    def hasLabel(self, name):
        return name in self._p.labels

    def getLabelValue(self, name, p):
        return self._p.getLabelValue(name, p)

    def getAttr(self, name):
        return self._attrs.get(name)

    def setAttr(self, name, value):
        self._attrs[name] = value


class _FakeSpace:
    def __init__(self, V, section):
        self.nnodes, self.bs, self.cell_nodes, self.section = V.nnodes, V.bs, V.cell_nodes, section


@pytest.mark.parametrize("name,level", [("ldc2d-sv-k2-tiny", 1), ("ldc3d-sv-k3-tiny", 1), ("ldc2d-pkp0-tiny", 2)])
def test_dmplex_view_reproduces_the_patch_sets(problems, name, level):
    """FiredrakeAdapter.plex(): the view built from getCone / getLabelValue / the section gives the same closure and
    star relations, labels, node attachment — hence the same patch dof sets — as the synthetic DMPlex look-alike."""
    from alfi_b200.firedrake_adapter import DMPlexView, FiredrakeAdapter
    from alfi_b200.patches import macro_interior_blocks, patch_dofs_from_points
    from alfi_b200.relaxation import macro_star_points, star_points
    from alfi_b200.synth.fakepetsc import FakePC
    prob = problems(name)
    ld = prob.levels[level]
    sp_plex = ld.level.plex
    dm = _FakeDMPlex(sp_plex)
    view = FiredrakeAdapter().plex(FakePC(dm))
    assert isinstance(view, DMPlexView) and FiredrakeAdapter().plex(FakePC(dm)) is view       # cached on the DM
    assert (view.closure != sp_plex.closure).nnz == 0 and (view.star != sp_plex.star).nnz == 0
    assert (view.cStart, view.cEnd, view.vStart, view.vEnd, view.npoints) == (
        sp_plex.cStart, sp_plex.cEnd, sp_plex.vStart, sp_plex.vEnd, sp_plex.npoints)
    npt = sp_plex.node_points(ld.V)
    V = _FakeSpace(ld.V, _FakeSection(npt, sp_plex.npoints))
    assert np.array_equal(view.node_points(V), npt)
    macro = prob.config.patch == "macro"
    Hs, _ = macro_star_points(sp_plex, prob.config.macro_expand) if macro else star_points(sp_plex)
    Hv, _ = macro_star_points(view, prob.config.macro_expand) if macro else star_points(view)
    assert (Hs != Hv).nnz == 0
    a = patch_dofs_from_points(sp_plex, ld.V, Hs, bc_nodes=ld.bc_nodes)
    b = patch_dofs_from_points(view, V, Hv, bc_nodes=ld.bc_nodes)
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.dofs, b.dofs)
    if macro:
        assert np.array_equal(macro_interior_blocks(sp_plex, ld.V, a), macro_interior_blocks(view, V, b))


def test_patchpc_over_a_petsc4py_shaped_dm(problems, monkeypatch):
    """`alfi_b200.PatchPC.initialize` with the PC's DM being a petsc4py-shaped DMPlex (no SynthPlex attribute is
    reachable) and the adapter handing out the DMPlexView: the python patch constructor (alfi.MacroStar, the
    reference's per-entity protocol) and PCPATCH's dof-set construction give the reference index sets."""
    import alfi_b200
    from alfi_b200 import pc as pcmod
    from alfi_b200.firedrake_adapter import DMPlexView
    from alfi_b200.synth.fakepetsc import FakePC, SynthAdapter
    from tests.test_pc_protocol import LEVEL_OPTS, Recorder

    class ViewAdapter(SynthAdapter):
        def plex(self, pc):
            dm = pc.getDM()
            if dm.getAttr("view") is None:
                dm.setAttr("view", DMPlexView(dm))
            return dm.getAttr("view")

        def function_space(self, pc):
            ld = self._ld()
            return _FakeSpace(ld.V, _FakeSection(ld.level.plex.node_points(ld.V), ld.level.plex.npoints))

    monkeypatch.setattr(pcmod, "Context", Recorder)
    prob = problems("ldc2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    dm = _FakeDMPlex(prob.levels[1].level.plex)
    # the reference's key functions read vertex coordinates with getVecClosure over the coordinate section
    # (relaxation.py:61-67): the petsc4py calls, answered from the synthetic mesh
    plex = prob.levels[1].level.plex
    dm.getCoordinateSection = lambda: "coordinate-section"
    dm.getCoordinatesLocal = lambda: "coordinate-vector"
    dm.getCoordinateDM = lambda: type("CDM", (), {"getDimension": staticmethod(lambda: plex.dim)})()

    def vec_closure(sec, vec, q):
        assert (sec, vec) == ("coordinate-section", "coordinate-vector")
        pts = plex.closure.indices[plex.closure.indptr[q]:plex.closure.indptr[q + 1]]
        v = pts[(pts >= plex.vStart) & (pts < plex.vEnd)] - plex.vStart
        return plex.mesh.coords[v].ravel()
    dm.getVecClosure = vec_closure
    pc = FakePC(dm, options=dict(LEVEL_OPTS), attrs={"alfi_b200_adapter": ViewAdapter(prob, 1)})
    p = alfi_b200.PatchPC()
    p.initialize(pc)
    ref = prob.levels[1].patches
    assert np.array_equal(p.patches.offsets, ref.offsets) and np.array_equal(p.patches.dofs, ref.dofs)
    assert np.array_equal(p.patches.order, ref.order)
