"""Patch index sets: the vectorised host builders against the
literal per-entity restatements of the reference callbacks, and against the analytic dof-count
histograms of SURVEY §8a (row P4 / T1).  Integer work: bit-exact."""
import numpy as np
import pytest

from alfi_b200.patches import greedy_colouring, patch_dofs_from_points, points_to_csr
from alfi_b200.relaxation import MacroStar, Star, iteration_order, macro_star_points, parse_sort_order, star_points
from alfi_b200.synth.fem import VectorSpace
from alfi_b200.synth.hierarchy import build_hierarchy
from alfi_b200.synth.mesh import alfeld_split, kuhn_mesh
from alfi_b200.synth.plex import SynthPlex
from alfi_b200.transfer import (CoarseCellMacroPatches, CoarseCellPatches, coarse_cell_points,
                                fix_coarse_boundaries, fix_coarse_boundaries_loop)
from oracle import pcpatch


class FakePC:
    """The slice of petsc4py.PC the patch constructors use (relaxation.py:110-113, transfer.py:17-19)."""

    def __init__(self, dm, options=None, ctx=None, prefix=""):
        self.dm, self.options, self.ctx, self.prefix = dm, options or {}, ctx, prefix

    def getDM(self):
        return self.dm

    def getOptionsPrefix(self):
        return self.prefix

    def getAttr(self, name):
        assert name == "ctx"
        return self.ctx


class Ctx:
    def __init__(self, hierarchy, level):
        self.hierarchy, self.level = hierarchy, level


def hist(ps):
    v, c = np.unique(ps.sizes, return_counts=True)
    return dict(zip(v.tolist(), c.tolist()))


@pytest.mark.parametrize("dim,M,k", [(2, 3, 2), (3, 2, 3)])
def test_star_vectorised_equals_callback(dim, M, k):
    mesh = alfeld_split(kuhn_mesh(dim, M))
    plex = SynthPlex(mesh)
    H, ents = star_points(plex)
    patches, order = Star()(FakePC(plex))
    assert (points_to_csr(patches, plex.npoints) != H).nnz == 0
    assert np.array_equal(order, np.arange(len(patches)))


@pytest.mark.parametrize("dim,M", [(2, 3), (3, 2)])
@pytest.mark.parametrize("expand", ["all", "vertices"])
def test_macrostar_vectorised_equals_callback(dim, M, expand):
    mesh = alfeld_split(kuhn_mesh(dim, M))
    plex = SynthPlex(mesh)
    H, ents = macro_star_points(plex, expand)
    ms = MacroStar()
    patches, _ = ms(FakePC(plex, {"pc_patch_construction_MacroStar_expand": expand}))
    assert ms.entities == list(ents)
    assert (points_to_csr(patches, plex.npoints) != H).nnz == 0


def test_macrostar_2d_literal_equals_open_macro_star():
    """In 2-D the reference's literal expansion is exactly the open macro star."""
    plex = SynthPlex(alfeld_split(kuhn_mesh(2, 4)))
    Ha, _ = macro_star_points(plex, "all")
    Hv, _ = macro_star_points(plex, "vertices")
    assert (Ha != Hv).nnz == 0


def test_patch_sizes_match_survey_histograms():
    # 2-D SV k=2 macro star: {8, 18, 28, 62}; 2-D P2 star: {2, 4, 14}
    mesh = alfeld_split(kuhn_mesh(2, 4))
    plex, V = SynthPlex(mesh), VectorSpace(mesh, 2)
    ps = patch_dofs_from_points(plex, V, macro_star_points(plex)[0], V.boundary_nodes())
    assert set(hist(ps)) == {8, 18, 28, 62} and hist(ps)[62] == 9
    mesh = kuhn_mesh(2, 4)
    plex, V = SynthPlex(mesh), VectorSpace(mesh, 2)
    ps = patch_dofs_from_points(plex, V, star_points(plex)[0], V.boundary_nodes())
    assert set(hist(ps)) - {0} == {2, 4, 14} and hist(ps)[14] == 9
    # 3-D SV k=3 macro star (vertex expansion): {93, 189, 294, 399, 609, 1275}
    mesh = alfeld_split(kuhn_mesh(3, 4))
    plex, V = SynthPlex(mesh), VectorSpace(mesh, 3)
    ps = patch_dofs_from_points(plex, V, macro_star_points(plex, "vertices")[0], V.boundary_nodes())
    assert hist(ps) == {93: 6, 189: 18, 294: 2, 399: 18, 609: 54, 1275: 27}
    cols = greedy_colouring(ps, V.ndofs)
    assert cols.max() + 1 == 8
    # the reference's literal MacroStar reaches around the link edges in 3-D: 2175 dofs inside
    ps = patch_dofs_from_points(plex, V, macro_star_points(plex, "all")[0], V.boundary_nodes())
    assert ps.sizes.max() == 2175


@pytest.mark.parametrize("dim,M,k,kind", [(2, 3, 2, "macro"), (2, 3, 2, "star"), (3, 2, 3, "macro"), (3, 1, 3, "macro-all")])
def test_pcpatch_dofs_vectorised_equals_literal(dim, M, k, kind):
    mesh = alfeld_split(kuhn_mesh(dim, M))
    plex, V = SynthPlex(mesh), VectorSpace(mesh, k)
    H = {"macro": lambda: macro_star_points(plex, "vertices")[0], "macro-all": lambda: macro_star_points(plex, "all")[0],
         "star": lambda: star_points(plex)[0]}[kind]()
    bc = V.boundary_nodes()
    ps = patch_dofs_from_points(plex, V, H, bc)
    sets = [H.indices[H.indptr[i]:H.indptr[i + 1]] for i in range(H.shape[0])]
    off, dofs = pcpatch.patch_dofs(plex, V, sets, bc)
    assert np.array_equal(off, ps.offsets)
    assert np.array_equal(dofs, ps.dofs)
    # colouring: host definition == literal definition, and it is a proper colouring
    cols = greedy_colouring(ps, V.ndofs)
    assert np.array_equal(cols, pcpatch.greedy_colouring(ps.offsets, ps.dofs, ps.order, V.ndofs))
    for c in range(cols.max() + 1):
        d = np.concatenate([ps.patch(p) for p in np.flatnonzero(cols == c)] or [np.empty(0, int)])
        assert np.unique(d).size == d.size


def test_sort_order_semantics():
    assert parse_sort_order("0+:1-|1+") == [[(0, 1), (1, -1)], [(1, 1)]]
    assert parse_sort_order("None") is None and parse_sort_order("") is None and parse_sort_order(None) is None
    coords = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    # same as sorted(enumerate(coords), key=lambda z: (z[1][0], -z[1][1]))  (relaxation.py:104-107)
    want = [i for i, _ in sorted(enumerate(coords), key=lambda z: (z[1][0], -z[1][1]))]
    assert iteration_order(coords, "0+:1-").tolist() == want
    two = iteration_order(coords, "0+:1-|1+", literal=False)
    assert two.size == 8 and two[:4].tolist() == want
    # the reference's key functions all see the last sweep's keys (closure over the loop variable,
    # relaxation.py:96-107; pinned by tests/test_reference_code.py): both sweeps sort by "1+"
    lit = iteration_order(coords, "0+:1-|1+")
    assert lit[:4].tolist() == lit[4:].tolist() == iteration_order(coords, "1+").tolist()
    plex = SynthPlex(alfeld_split(kuhn_mesh(2, 2)))
    patches, order = MacroStar()(FakePC(plex, {"pc_patch_construction_MacroStar_sort_order": "0+:1-"}))
    assert sorted(order.tolist()) == list(range(len(patches)))


@pytest.mark.parametrize("dim,N,k,bary", [(2, 2, 2, True), (2, 2, 2, False), (3, 1, 3, True)])
def test_transfer_cell_patches(dim, N, k, bary):
    levels = build_hierarchy(dim, N, 1, bary)
    fine = levels[1]
    V = VectorSpace(fine.mesh, k)
    H = coarse_cell_points(levels, 1, bary)
    cls = CoarseCellMacroPatches if bary else CoarseCellPatches
    patches, order = cls()(FakePC(fine.plex, ctx=Ctx(levels, 1)))
    assert (points_to_csr(patches, fine.plex.npoints) != H).nnz == 0
    cb = fix_coarse_boundaries(fine.plex, V, 1)
    assert np.array_equal(cb, fix_coarse_boundaries_loop(fine.plex, V, 1))
    ps = patch_dofs_from_points(fine.plex, V, H, cb)
    want = {(2, True): 38, (2, False): 6, (3, True): 390}[(dim, bary)]
    assert set(ps.sizes.tolist()) == {want}
    assert ps.npatch == levels[0].macro.nc
    # cell patches are dof-disjoint and, with the coarse boundary, cover the space
    assert np.unique(ps.dofs).size == ps.dofs.size
    assert ps.dofs.size + cb.size * V.bs == V.ndofs
