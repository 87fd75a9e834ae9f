"""[P1+FacetBubble]^3 (BASELINE configs[3]) and its flux-preserving BubbleTransfer (alfi/bubble.py):
matrix form (product path) against the literal per-cell restatement of the reference's C kernels,
plus the property the transfer exists for — the flux across every coarse facet is preserved."""
import os

import numpy as np
import pytest

from alfi_b200.bubble import BubbleTransfer, bubble_transfer_matrix
from alfi_b200.synth.fem import P1FBElement, VectorSpace, assemble_velocity_block
from alfi_b200.synth.hierarchy import build_hierarchy, prolongation_matrix
from alfi_b200.synth.mesh import kuhn_mesh
from oracle.bubble import LiteralBubbleTransfer


def facet_flux(V, u):
    """∫_F u.n over every face F of the mesh for a P1FB function (n = the face's fixed unit normal).
    On F only the three vertex functions and F's bubble are non-zero: ∫λ_v = |F|/3, ∫b_F = 0.45|F|."""
    m = V.mesh
    u = u.reshape(V.nnodes, 3)
    X = m.coords[m.faces]
    nA = 0.5 * np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0])            # area * normal
    uv = u[V.vertex_nodes[:, 0]][m.faces].sum(axis=1)                     # sum of the 3 vertex values
    uf = u[V.face_nodes[:, 0]]
    return np.einsum("fa,fa->f", (1.0 / 3.0 - 0.15) * uv + 0.45 * uf, nA)


def test_element_is_nodal_and_contains_p1():
    el = P1FBElement()
    assert np.allclose(el.tabulate(el.nodes_ref), np.eye(8), atol=1e-14)
    x = np.random.default_rng(0).random((7, 3)) * 0.3
    f = lambda p: 1 + 2 * p[:, 0] - 3 * p[:, 1] + 0.5 * p[:, 2]          # noqa: E731
    assert np.allclose(el.tabulate(x) @ f(el.nodes_ref), f(x), atol=1e-14)
    M = 2
    V = VectorSpace(kuhn_mesh(3, M), 1, "p1fb")
    assert V.nnodes == (M + 1) ** 3 + 12 * M ** 3 + 6 * M ** 2             # SURVEY §8d node-count formula
    A = assemble_velocity_block(V, 1.0, 1.0, divform="pkp0").to_csr()
    assert abs(A - A.T).max() < 1e-13
    xc = V.node_coords
    rot = np.stack([-xc[:, 1], xc[:, 0], 0 * xc[:, 0]], 1).ravel()
    assert abs(A @ rot).max() < 1e-12                                     # rigid rotation: no strain, no divergence


@pytest.fixture(scope="module")
def pair():
    lev = build_hierarchy(3, 1, 1, False)
    Vc, Vf = VectorSpace(lev[0].mesh, 1, "p1fb"), VectorSpace(lev[1].mesh, 1, "p1fb")
    return lev, Vc, Vf


def test_matrix_form_equals_literal_kernels(pair):
    lev, Vc, Vf = pair
    P = bubble_transfer_matrix(Vc, Vf, lev[0].c2f)
    lit = LiteralBubbleTransfer(Vc, Vf, lev[0].c2f)
    c = np.random.default_rng(1).standard_normal(Vc.ndofs)
    assert np.abs(P @ c - lit.prolong(c)).max() < 1e-13
    bt = BubbleTransfer(Vc, Vf, lev[0].c2f)
    fine, coarse = np.empty(Vf.ndofs), np.empty(Vc.ndofs)
    bt.prolong(c, fine)
    f = np.random.default_rng(2).standard_normal(Vf.ndofs)
    bt.restrict(f, coarse)
    assert abs(f @ fine - coarse @ c) < 1e-12 * abs(f @ fine)            # restrict = prolong^T (bubble.py:204-231)


def test_flux_across_coarse_facets_is_preserved(pair):
    """bubble.py:1-6, 247-250: the standard prolongation loses 37.5 % of a bubble's flux across the
    coarse facet, the corrected one none."""
    lev, Vc, Vf = pair
    mc, mf = Vc.mesh, Vf.mesh
    P = bubble_transfer_matrix(Vc, Vf, lev[0].c2f)
    Pstd = prolongation_matrix(Vc, Vf, lev[0].c2f)
    rng = np.random.default_rng(3)
    c = rng.standard_normal(Vc.ndofs)
    # fine faces lying on a coarse face: all three vertices on the coarse face's plane and inside it
    fc = facet_flux(Vc, c)
    ff = facet_flux(Vf, P @ c)
    ff_std = facet_flux(Vf, (np.kron(Pstd.toarray(), np.eye(3))) @ c)
    cent = mf.coords[mf.faces].mean(axis=1)
    worst, worst_std = 0.0, 0.0
    for F in range(mc.nf):
        X = mc.coords[mc.faces[F]]
        n = np.cross(X[1] - X[0], X[2] - X[0])
        on_plane = np.abs((mf.coords[mf.faces] - X[0]) @ n).max(axis=1) < 1e-12
        # barycentric test of the fine centroid inside the coarse triangle
        T = np.stack([X[1] - X[0], X[2] - X[0]], axis=1)
        lam = np.linalg.lstsq(T, (cent - X[0]).T, rcond=None)[0].T
        inside = on_plane & (lam.min(axis=1) > -1e-12) & (lam.sum(axis=1) < 1 + 1e-12)
        assert inside.sum() == 4
        # orientation: fine normals may be flipped w.r.t. the coarse one
        Xf = mf.coords[mf.faces[inside]]
        nf = np.cross(Xf[:, 1] - Xf[:, 0], Xf[:, 2] - Xf[:, 0])
        sign = np.sign(nf @ n)
        worst = max(worst, abs((sign * ff[inside]).sum() - fc[F]))
        worst_std = max(worst_std, abs((sign * ff_std[inside]).sum() - fc[F]))
    assert worst < 1e-13
    assert worst_std > 1e-3                                               # the uncorrected transfer does not


# ---------------------------------------------------------------- the reference's own C kernels (oracle/_ref)
def _ref_kernels():
    from oracle import build_ref
    return build_ref.load()


def _ptr(a):
    import ctypes as C
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.skipif(_ref_kernels() is None, reason="oracle/_ref not built and no reference tree to build it from")
def test_kernel_tables_equal_the_compiled_reference_kernels():
    """split / splitadj / combine / combineadj / count of alfi/bubble.py:57-185, compiled by oracle/build_ref.py
    from the reference's own source strings, against the tables oracle/bubble.py restates them with."""
    from oracle.bubble import A_COMB, A_SPLIT, B_COMB, B_SPLIT
    lib = _ref_kernels()
    rng = np.random.default_rng(4)
    for _ in range(5):
        both = rng.standard_normal((8, 3))
        p1, fb = np.zeros((4, 3)), np.zeros((4, 3))
        lib.split(_ptr(p1), _ptr(fb), _ptr(both))
        assert np.array_equal(p1, A_SPLIT.T @ both) and np.allclose(fb, B_SPLIT.T @ both, rtol=0, atol=2e-16 * 8)
        p1, fb = rng.standard_normal((4, 3)), rng.standard_normal((4, 3))
        out = np.zeros((8, 3))
        lib.splitadj(_ptr(p1), _ptr(fb), _ptr(out))
        assert np.allclose(out, A_SPLIT @ p1 + B_SPLIT @ fb, rtol=0, atol=1e-15)
        out = np.zeros((8, 3))
        lib.combine(_ptr(p1), _ptr(fb), _ptr(out))
        assert np.allclose(out, A_COMB.T @ p1 + B_COMB.T @ fb, rtol=0, atol=1e-15)
        both = rng.standard_normal((8, 3))
        p1, fb = np.zeros((4, 3)), np.zeros((4, 3))
        lib.combineadj(_ptr(both), _ptr(p1), _ptr(fb))
        assert np.allclose(p1, A_COMB @ both, rtol=0, atol=1e-15) and np.allclose(fb, B_COMB @ both, rtol=0, atol=1e-15)
    both, fb, p1 = np.zeros((8, 3)), np.zeros((4, 3)), np.zeros((4, 3))
    lib.count(_ptr(both), _ptr(fb), _ptr(p1))
    assert (both == 1).all() and (fb == 1).all() and (p1 == 1).all()


@pytest.mark.skipif(_ref_kernels() is None, reason="oracle/_ref not built and no reference tree to build it from")
def test_matrix_form_equals_the_sequence_run_with_the_reference_kernels(pair):
    """bubble.py:233-265 (prolong) with the cell loops executing the compiled reference kernels — the par_loops
    of the reference, INC access = scatter-add — equals the dof-level CSR the library is given."""
    lib = _ref_kernels()
    lev, Vc, Vf = pair
    lit = LiteralBubbleTransfer(Vc, Vf, lev[0].c2f)

    def par_loop_split(V, both):
        p1, fb = np.zeros_like(both), np.zeros_like(both)
        for c in range(V.mesh.nc):
            nodes = V.cell_nodes[c]
            a, b = np.zeros((4, 3)), np.zeros((4, 3))
            lib.split(_ptr(a), _ptr(b), _ptr(np.ascontiguousarray(both[nodes])))
            p1[nodes[:4]] += a
            fb[nodes[4:]] += b
        return p1, fb

    def par_loop_combine(V, p1, fb):
        both = np.zeros_like(p1)
        for c in range(V.mesh.nc):
            nodes = V.cell_nodes[c]
            out = np.zeros((8, 3))
            lib.combine(_ptr(np.ascontiguousarray(p1[nodes[:4]])), _ptr(np.ascontiguousarray(fb[nodes[4:]])), _ptr(out))
            both[nodes] += out
        return both

    def counts(V):
        both = np.zeros((V.nnodes, 3))
        for c in range(V.mesh.nc):
            nodes = V.cell_nodes[c]
            a, b, p = np.zeros((8, 3)), np.zeros((4, 3)), np.zeros((4, 3))
            lib.count(_ptr(a), _ptr(b), _ptr(p))
            both[nodes] += a
        return both[:, 0]

    assert np.array_equal(counts(Vc), lit.cnt["c"]) and np.array_equal(counts(Vf), lit.cnt["f"])
    c = np.random.default_rng(5).standard_normal(Vc.ndofs)
    coarse = c.reshape(Vc.nnodes, 3)
    p1c, fbc = par_loop_split(Vc, coarse)
    cv, cf = lit.cnt["c"], lit.cnt["f"]
    p1c[Vc.vertex_nodes[:, 0]] /= cv[Vc.vertex_nodes[:, 0], None]
    fbc[Vc.face_nodes[:, 0]] /= cv[Vc.face_nodes[:, 0], None]
    fbc = lit.scale_normal(Vc, fbc)
    fine = par_loop_combine(Vf, lit.point_prolong("p1", p1c), lit.point_prolong("fb", fbc)) / cf[:, None]
    P = bubble_transfer_matrix(Vc, Vf, lev[0].c2f)
    assert np.abs(P @ c - fine.ravel()).max() < 1e-13
    assert np.abs(lit.prolong(c) - fine.ravel()).max() < 1e-13


@pytest.mark.skipif(_ref_kernels() is None or not os.path.exists(os.path.join(os.environ.get("ALFI_REFERENCE", "/root/reference"), "alfi", "bubble.py")),
                    reason="needs the reference tree (build container)")
def test_reference_bubble_transfer_methods_equal_the_matrix_form(pair):
    """alfi/bubble.py:204-265 executed verbatim (oracle/refshim_bubble.py: par_loops run the reference's compiled
    kernels) == the dof-level CSR handed to the library, and its restrict == the transpose."""
    from oracle.refshim_bubble import Harness
    lev, Vc, Vf = pair
    h = Harness(Vc, Vf, lev[0].c2f)
    P = bubble_transfer_matrix(Vc, Vf, lev[0].c2f)
    rng = np.random.default_rng(6)
    c, f = rng.standard_normal(Vc.ndofs), rng.standard_normal(Vf.ndofs)
    with h.reference_class() as BubbleTransfer:
        coarse, fine = h.full(Vc, c), h.full(Vf)
        BubbleTransfer.prolong(h.me, coarse, fine)
        assert np.abs(fine.dat.data.ravel() - P @ c).max() < 1e-13
        fine2, coarse2 = h.full(Vf, f), h.full(Vc)
        BubbleTransfer.restrict(h.me, fine2, coarse2)
        assert np.abs(coarse2.dat.data.ravel() - P.T @ f).max() < 1e-12
