"""BASELINE.json configs[2] — backward-facing step on a Gmsh mesh (examples/bfs2d/bfs2d.py): the MSH 2.2
reader, uniform refinement of a general triangle mesh with boundary markers, patch index sets on an
unstructured barycentric mesh (varying vertex valence), tagged Dirichlet conditions with a natural outflow,
and the velocity-block multigrid of the CPU oracle on it.  Host side only; the CUDA library is mesh-agnostic
and is run on this configuration by tests/test_gpu_bfs.py."""
import os

import numpy as np
import pytest

from alfi_b200.patches import greedy_colouring, patch_dofs_from_points, points_to_csr
from alfi_b200.relaxation import MacroStar, iteration_order, macro_star_points
from alfi_b200.synth.fem import VectorSpace
from alfi_b200.synth.gmsh import INFLOW, NOSLIP, OUTFLOW, read_msh, step_mesh, write_msh
from alfi_b200.synth.hierarchy import build_hierarchy_from
from alfi_b200.synth.mesh import alfeld_split, kuhn_mesh, refine_uniform
from alfi_b200.synth.plex import SynthPlex
from alfi_b200.transfer import cell_patch_set
from oracle import hotpath as hp
from oracle import pcpatch
from tests.test_patches import FakePC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "step_n2.msh")
REFERENCE_MSH = os.path.join(os.environ.get("ALFI_REFERENCE", "/root/reference"), "examples", "bfs2d", "coarse09.msh")


def areas(m):
    X = m.coords[m.cells]
    a, b = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]
    return 0.5 * np.abs(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])


def boundary_edge_mask(m):
    return np.bincount(m.cell_edges.ravel(), minlength=m.ne) == 1


def check_step_domain(m):
    assert abs(areas(m).sum() - 19.0) < 1e-10 and areas(m).min() > 0          # [0,10]x[0,2] minus [0,1]x[0,1]
    assert m.nv - m.ne + m.nc == 1                                            # simply connected
    assert np.array_equal(boundary_edge_mask(m), m.facet_tag > 0)             # every boundary edge is tagged
    mid = m.coords[m.edges].mean(axis=1)
    assert np.allclose(mid[m.facet_tag == INFLOW][:, 0], 0.0) and (mid[m.facet_tag == INFLOW][:, 1] > 1).all()
    assert np.allclose(mid[m.facet_tag == OUTFLOW][:, 0], 10.0)
    assert set(np.unique(m.facet_tag).tolist()) == {0, INFLOW, NOSLIP, OUTFLOW}


def test_msh_fixture_round_trip(tmp_path):
    """The committed fixture is what `write_msh(step_mesh(2))` produces, and reading it gives the mesh back."""
    m = step_mesh(2)
    check_step_domain(m)
    r = read_msh(FIXTURE)
    assert np.array_equal(r.cells, m.cells) and np.allclose(r.coords, m.coords, atol=1e-15)
    assert np.array_equal(r.facet_tag, m.facet_tag)
    p = tmp_path / "again.msh"
    write_msh(str(p), r)
    assert open(p).read() == open(FIXTURE).read()


@pytest.mark.skipif(not os.path.exists(REFERENCE_MSH), reason="reference tree not present (GPU box)")
def test_reads_the_reference_mesh():
    """examples/bfs2d/coarse09.msh, the mesh BASELINE.json configs[2] is quoted on: 2979 vertices."""
    m = read_msh(REFERENCE_MSH)
    assert m.nv == 2979 and m.nc == 5685
    check_step_domain(m)
    valence = np.bincount(m.edges.ravel())
    assert valence.min() >= 3 and valence.max() == 8 and np.bincount(valence).argmax() == 6
    fine, c2f = refine_uniform(m)
    check_step_domain(fine)
    assert fine.nc == 4 * m.nc and np.allclose(areas(fine)[c2f].sum(axis=1), areas(m))


def test_unsupported_files_are_rejected(tmp_path):
    p = tmp_path / "bad.msh"
    p.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError, match="MSH 2"):
        read_msh(str(p))
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n")
    with pytest.raises(ValueError, match="triangles"):
        read_msh(str(p))


def test_refinement_matches_the_kuhn_hierarchy():
    """On a Kuhn mesh the general red refinement is the Kuhn mesh with twice as many cells per side."""
    m = kuhn_mesh(2, 3)
    fine, c2f = refine_uniform(m)
    ref = kuhn_mesh(2, 6)
    assert (fine.nv, fine.ne, fine.nc) == (ref.nv, ref.ne, ref.nc)
    key = lambda mesh: np.sort(np.round(mesh.coords[mesh.cells].mean(axis=1) * 1e6).astype(np.int64) @ np.array([1, 10 ** 8]))  # noqa: E731
    assert np.array_equal(key(fine), key(ref))
    assert np.allclose(areas(fine)[c2f].sum(axis=1), areas(m))
    cent = fine.coords[fine.cells].mean(axis=1)
    X = m.coords[m.cells][np.repeat(np.arange(m.nc), 4)]
    T = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))
    lam = np.linalg.solve(T, (cent[c2f.ravel()] - X[:, 0])[:, :, None])[:, :, 0]
    assert (lam.min(axis=1) > 0).all() and (lam.sum(axis=1) < 1).all()          # children lie in their parent


def test_tags_survive_refinement_and_alfeld_split():
    m = step_mesh(1)
    levels = build_hierarchy_from(m, 2, True)
    for lev in levels:
        check_step_domain(lev.macro)
        split = lev.mesh
        assert abs(areas(split).sum() - 19.0) < 1e-10
        assert np.array_equal(boundary_edge_mask(split), split.facet_tag > 0)
        n = [int((lev.macro.facet_tag == t).sum()) for t in (INFLOW, NOSLIP, OUTFLOW)]
        assert n == [2 ** lev.index * c for c in (1, 21, 2)]
        assert [int((split.facet_tag == t).sum()) for t in (INFLOW, NOSLIP, OUTFLOW)] == n
    V = VectorSpace(levels[1].mesh, 2)
    bc = V.tagged_boundary_nodes((INFLOW, NOSLIP))
    x = V.node_coords[bc]
    on_wall = ((np.abs(x[:, 1] - 2) < 1e-12) | (np.abs(x[:, 0]) < 1e-12) | ((x[:, 0] <= 1 + 1e-12) & (np.abs(x[:, 1] - 1) < 1e-12))
               | ((np.abs(x[:, 0] - 1) < 1e-12) & (x[:, 1] <= 1 + 1e-12)) | (np.abs(x[:, 1]) < 1e-12))
    assert on_wall.all()
    out = V.tagged_boundary_nodes((OUTFLOW,))
    free = np.setdiff1d(out, bc)
    assert free.size == out.size - 2 and np.allclose(V.node_coords[free][:, 0], 10.0)   # outflow stays natural


def test_macro_star_patches_on_an_unstructured_mesh():
    """Vectorised builders == per-entity callbacks == literal PCPATCH loops on a mesh with valences 2..8;
    interior macro stars of valence v have 10 v + 2 dofs (SV k=2: 62 for v = 6)."""
    macro = refine_uniform(step_mesh(1, seed=3))[0]
    mesh = alfeld_split(macro)
    plex, V = SynthPlex(mesh), VectorSpace(mesh, 2)
    H, ents = macro_star_points(plex, "vertices")
    ms = MacroStar()
    patches, order = ms(FakePC(plex, {"pc_patch_construction_MacroStar_sort_order": "0+:1-"}))
    assert ms.entities == list(ents) and (points_to_csr(patches, plex.npoints) != H).nnz == 0
    coords = np.array([plex.point_coords(p) for p in ents])
    assert np.array_equal(order, iteration_order(coords, "0+:1-"))
    assert (np.diff(coords[order][:, 0]) >= 0).all()                           # swept downstream (bfs2d.py:32)
    bc = V.tagged_boundary_nodes((INFLOW, NOSLIP))
    ps = patch_dofs_from_points(plex, V, H, bc_nodes=bc, order=order)
    sets = [H.indices[H.indptr[i]:H.indptr[i + 1]] for i in range(H.shape[0])]
    off, dofs = pcpatch.patch_dofs(plex, V, sets, bc)
    assert np.array_equal(off, ps.offsets) and np.array_equal(dofs, ps.dofs)
    cols = greedy_colouring(ps, V.ndofs)
    assert np.array_equal(cols, pcpatch.greedy_colouring(ps.offsets, ps.dofs, ps.order, V.ndofs))
    valence = np.bincount(macro.edges.ravel(), minlength=macro.nv)
    interior = ~np.isin(np.arange(macro.nv), np.unique(macro.edges[macro.facet_tag > 0]))
    vid = np.asarray(ents) - plex.vStart
    for p in np.flatnonzero(interior[vid]):
        assert ps.sizes[p] == 10 * valence[vid[p]] + 2, (p, valence[vid[p]], ps.sizes[p])
    assert len(set(valence[interior].tolist())) >= 3                           # the mesh really is irregular


def test_cell_patches_have_38_dofs():
    """Transfer patches are per coarse macro cell, independent of the vertex valence."""
    hier = build_hierarchy_from(step_mesh(1, seed=3), 1, True)
    V = VectorSpace(hier[1].mesh, 2)
    cps, cb = cell_patch_set(hier, 1, V, True)
    assert cps.npatch == hier[0].macro.nc and set(cps.sizes.tolist()) == {38}
    d = np.concatenate([cps.patch(p) for p in range(cps.npatch)])
    assert np.unique(d).size == d.size                                         # disjoint: one colour


@pytest.fixture(scope="module")
def bfs(problems):
    prob = problems("bfs2d-sv-k2-tiny")
    return prob, [hp.level_from_host(l) for l in prob.levels]


def test_problem_shapes(bfs):
    prob, lv = bfs
    fine = prob.finest
    assert prob.config.nu == 1.0 / prob.config.re                              # char_length 1 (alfi/problem.py:43)
    assert fine.patches.npatch == prob.levels[1].level.macro.nv
    x = fine.V.node_coords[fine.bc_nodes]
    assert not np.any(np.abs(x[:, 0] - 10.0) < 1e-12) or np.all(np.abs(x[np.abs(x[:, 0] - 10.0) < 1e-12][:, 1] % 2) < 1e-12)
    order = fine.patches.order
    assert sorted(order.tolist()) == list(range(fine.patches.npatch))
    assert fine.patches.blocks is not None and (fine.patches.blocks >= 0).any()


def test_cycle_contracts_and_transfers_are_adjoint(bfs):
    prob, lv = bfs
    L, Lc = lv[1], lv[0]
    rng = np.random.default_rng(11)
    c = rng.standard_normal(Lc.n)
    c[Lc.bc_dofs] = 0
    f = rng.standard_normal(L.n)
    f[L.bc_dofs] = 0
    lhs, rhs = f @ hp.prolong(L, c), hp.restrict(L, f, Lc.bc_dofs) @ c
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), 1.0)
    b = rng.standard_normal(L.n)
    b[L.bc_dofs] = 0
    x = hp.fcycle(lv, b, prob.config.m)
    assert np.linalg.norm(b - L.A @ x) < 0.2 * np.linalg.norm(b)


def test_condensed_lists_on_the_unstructured_mesh(bfs):
    """Macro-cell blocks are shared by 3 vertex patches each, whatever the valences; the condensed apply
    (host restatement of the CUDA kernels, tests/condense_host_shim.cpp) equals the dense patch solves."""
    from tests.test_condense_host import Host, rel
    import ctypes as C
    import subprocess
    out = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libcondense_host_shim.so")
    src = os.path.join(ROOT, "tests", "condense_host_shim.cpp")
    hdr = os.path.join(ROOT, "alfi_b200", "csrc", "condense_host.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.dirname(hdr), src, "-o", so])
    lib = C.CDLL(so)
    from tests.test_condense_host import _f64p, _i32p, _i64p
    lib.ch_create.restype = C.c_void_p
    lib.ch_create.argtypes = [C.c_int, C.c_int, _i32p, _i32p, C.c_int, _i64p, _i32p, C.c_int, _i32p, _i32p, C.c_int,
                              _i32p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.ch_destroy.argtypes = [C.c_void_p]
    lib.ch_stats.argtypes = [C.c_void_p, _i64p]
    lib.ch_factor.argtypes = [C.c_void_p, _f64p]
    lib.ch_apply.argtypes = [C.c_void_p, _f64p, _f64p]
    lib.ch_check_disjoint.argtypes = [C.c_void_p]
    prob, lv = bfs
    ld = prob.finest
    ps = ld.patches
    host = Host(lib, ld, ps, ps.blocks, True)
    assert host.h, host.err
    st = host.stats()
    assert st["shared"] == 1 and st["ndist"] == ld.level.macro.nc and st["nblocks"] == 3 * st["ndist"]
    assert lib.ch_check_disjoint(host.h) == 0 and host.factor(ld.A.vals) == 0
    mats = hp.patch_matrices(ld.A.to_csr(), ps.offsets, ps.dofs)
    x = np.random.default_rng(12).standard_normal(ld.V.ndofs)
    want = np.zeros_like(x)
    for p in ps.order:
        I = ps.patch(p)
        if I.size:
            want[I] += np.linalg.solve(mats[p], x[I])
    kappa = max(np.linalg.cond(M) for M in mats if M.size)
    assert rel(host.apply(x), want) <= 1e-11 * max(1.0, kappa * np.finfo(float).eps / 1e-12)
    host.close()
