"""[P1+FacetBubble]^3 (BASELINE configs[3]) and its flux-preserving BubbleTransfer (alfi/bubble.py):
matrix form (product path) against the literal per-cell restatement of the reference's C kernels,
plus the property the transfer exists for — the flux across every coarse facet is preserved."""
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
