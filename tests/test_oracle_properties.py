"""First-principles checks of the CPU oracle (the reference pins nothing — SURVEY §4, §8c(v))."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import hotpath as hp

NAMES = ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny", "ldc3d-pkp0-tiny", "bfs2d-sv-k2-tiny"]


@pytest.fixture(scope="module")
def olevels(problems):
    cache = {}

    def get(name, mode="inverse", **kw):
        key = (name, mode, tuple(sorted(kw.items())))
        if key not in cache:
            prob = problems(name, **kw)
            cache[key] = (prob, [hp.level_from_host(l, mode) for l in prob.levels])
        return cache[key]
    return get


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", NAMES)
def test_additive_schwarz_equals_dense_formula(olevels, name):
    """y = sum_i R_i^T (R_i A R_i^T)^-1 R_i x computed densely, then y[bc] = x[bc]."""
    prob, lv = olevels(name, gamma=10.0, nu=0.2)
    L = lv[1]
    A = L.A.toarray()
    x = np.random.default_rng(1).standard_normal(L.n)
    want = np.zeros(L.n)
    for p in L.order:
        I = L.dofs[L.offsets[p]:L.offsets[p + 1]]
        if I.size == 0:
            continue
        R = sp.csr_matrix((np.ones(I.size), (np.arange(I.size), I)), shape=(I.size, L.n)).toarray()
        want += R.T @ np.linalg.solve(R @ A @ R.T, R @ x)
    want[L.bc_dofs] = x[L.bc_dofs]
    got = hp.smoother_apply(x, L.offsets, L.dofs, L.order, L.factors, L.bc_dofs)
    assert rel(got, want) < 1e-12
    _, lu = olevels(name, "lu", gamma=10.0, nu=0.2)
    got_lu = hp.smoother_apply(x, L.offsets, L.dofs, L.order, lu[1].factors, L.bc_dofs)
    assert rel(got_lu, want) < 1e-12


@pytest.mark.parametrize("name", NAMES)
def test_restrict_is_prolong_transpose(olevels, name):
    """The block solve only reads and writes cell-interior dofs, A0 and D are symmetric, so the
    sequence of transfer.py:261-275 is the exact adjoint of transfer.py:246-259 on vectors that
    vanish on the Dirichlet boundary."""
    prob, lv = olevels(name, gamma=10.0, nu=0.2)
    L, Lc = lv[1], lv[0]
    rng = np.random.default_rng(2)
    c = rng.standard_normal(Lc.n)
    c[Lc.bc_dofs] = 0
    f = rng.standard_normal(L.n)
    f[L.bc_dofs] = 0
    lhs = f @ hp.prolong(L, c)
    rhs = hp.restrict(L, f, Lc.bc_dofs) @ c
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)
    lhs = f @ hp.prolong(L, c, robust=False)
    rhs = hp.restrict(L, f, Lc.bc_dofs, robust=False) @ c
    assert abs(lhs - rhs) <= 1e-13 * max(abs(lhs), 1.0)


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc3d-sv-k3-tiny", "bfs2d-sv-k2-tiny"])
def test_schoeberl_prolongation_property(olevels, name):
    """Defining property of the robust transfer (transfer.py:246-259): the correction t solves
    A0 t = gamma D rhs on every coarse-cell interior with zero trace on the coarse facets; for
    these gamma-dominated problems that removes most of the divergence energy the plain
    prolongation creates, so the A0-energy must not increase."""
    prob, lv = olevels(name)
    L, Lc = lv[1], lv[0]
    A0 = prob.levels[1].A0.to_csr()
    c = np.random.default_rng(3).standard_normal(Lc.n)
    c[Lc.bc_dofs] = 0
    plain = hp.prolong(L, c, robust=False)
    robust = hp.prolong(L, c, robust=True)
    assert robust @ (A0 @ robust) <= (plain @ (A0 @ plain)) * (1 + 1e-12)
    rhs = L.P @ c
    b = L.D @ rhs
    b[L.cb_dofs] = 0
    t = hp._block_solve(L, b)
    interior = L.c_dofs
    assert rel((A0 @ t)[interior], b[interior]) < 1e-9
    # traces on coarse facets are untouched by the correction
    assert np.array_equal(t[L.cb_dofs], np.zeros(L.cb_dofs.size))


@pytest.mark.parametrize("m", [1, 3, 6])
def test_fgmres_minimises_the_residual(m):
    """x_m - x0 in span(Z) minimises ||b - A x|| : compare with a dense least-squares solve."""
    rng = np.random.default_rng(4)
    n = 40
    A = rng.standard_normal((n, n)) + 6 * np.eye(n)
    Minv = np.linalg.inv(A + 0.5 * rng.standard_normal((n, n)))
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    x = hp.fgmres(lambda v: A @ v, lambda v: Minv @ v, b, x0, m)
    # rebuild Z by running the same Arnoldi process in exact terms: Krylov space of A Minv on r0
    r0 = b - A @ x0
    K = [r0]
    for _ in range(m - 1):
        K.append(A @ (Minv @ K[-1]))
    Z = np.stack([Minv @ k for k in K], axis=1)
    y = np.linalg.lstsq(A @ Z, r0, rcond=None)[0]
    assert rel(x, x0 + Z @ y) < 1e-9
    assert np.linalg.norm(b - A @ x) <= np.linalg.norm(r0) * (1 + 1e-12)


def test_fgmres_zero_rhs_and_happy_breakdown():
    A = np.diag([1.0, 2.0, 3.0])
    x = hp.fgmres(lambda v: A @ v, lambda v: v, np.zeros(3), np.zeros(3), 4)
    assert np.array_equal(x, np.zeros(3))
    b = np.array([1.0, 0.0, 0.0])                       # Krylov space has dimension 1
    x = hp.fgmres(lambda v: A @ v, lambda v: v, b, np.zeros(3), 3)
    assert np.allclose(A @ x, b)


@pytest.mark.parametrize("name", NAMES)
def test_fcycle_is_the_petsc_full_cycle(olevels, name):
    """PCMG full (Appendix A.5) written out by hand for the 2- and 3-level cases."""
    prob, lv = olevels(name, gamma=10.0, nu=0.2)
    m = prob.config.m
    b = np.random.default_rng(5).standard_normal(lv[-1].n)
    b[lv[-1].bc_dofs] = 0
    got = hp.fcycle(lv, b, m)
    nl = len(lv)
    bs = {nl - 1: b}
    for l in range(nl - 1, 0, -1):
        bs[l - 1] = hp.restrict(lv[l], bs[l], lv[l - 1].bc_dofs)

    def V(l, bl, xl):
        if l == 0:
            return hp.coarse_solve(lv[0], bl)
        xl = hp.smooth(lv[l], bl, xl, m)
        rc = hp.restrict(lv[l], bl - lv[l].A @ xl, lv[l - 1].bc_dofs)
        xl = xl + hp.prolong(lv[l], V(l - 1, rc, np.zeros_like(rc)))
        return hp.smooth(lv[l], bl, xl, m)
    x = V(0, bs[0], None)
    for l in range(1, nl):
        x = V(l, bs[l], hp.prolong(lv[l], x))
    assert rel(got, x) < 1e-12
    # and it is a contraction for this SPD-dominated problem
    assert np.linalg.norm(b - lv[-1].A @ got) < 0.5 * np.linalg.norm(b)
