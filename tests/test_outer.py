"""The pieces that bracket fieldsplit_0 in alfi's outer solver (alfi/solver.py:15-38, 405-421,
463-474): Schur-complement fieldsplit application, DGMassInv, the B / B^T products and the outer FGMRES.

CPU: the oracle restatement (oracle/outer.py) against dense linear algebra and against the host stand-in.
GPU: csrc/outer.cu through the C-ABI against the oracle — per application, per linear solve, and as the linear solver
of the Newton continuation (identical iteration counts)."""
import dataclasses

import numpy as np
import pytest
import scipy.sparse as sp

from alfi_b200.synth.fem import assemble_divergence
from alfi_b200.synth.outer import ContinuationSolver, fgmres_outer
from alfi_b200.synth.problem import CONFIGS
from oracle import outer as oo

SMALL2D = dataclasses.replace(CONFIGS["ldc2d-sv-k2"], N=4)


def test_oracle_fgmres_equals_the_host_stand_in_and_restarts():
    rng = np.random.default_rng(3)
    A = 0.6 * rng.standard_normal((80, 80)) + 9 * np.eye(80)
    b = rng.standard_normal(80)
    Aop, Mop = (lambda v: A @ v), (lambda v: v / 9.0)
    x1, its1, h1 = oo.fgmres(Aop, Mop, b, 1e-12, 0.0, maxit=200, restart=12)
    x2, its2, h2 = fgmres_outer(Aop, Mop, b, 1e-12, 0.0, maxit=200, restart=12)
    assert its1 == its2 and its1 > 12
    assert np.allclose(h1, h2, rtol=1e-6, atol=1e-9 * h1[0]) and np.linalg.norm(x1 - x2) <= 1e-11 * np.linalg.norm(x2)
    assert np.linalg.norm(A @ x1 - b) <= 1e-10 * np.linalg.norm(b)
    assert oo.fgmres(Aop, Mop, np.zeros(80), 1e-8, 1e-10)[1] == 0            # zero right-hand side: no iteration


def test_oracle_schur_apply_is_the_full_block_factorisation():
    """With exact velocity solves P^-1 = [I -A^-1 B^T; 0 I] [A^-1 0; 0 S^-1] [I 0; -B A^-1 I] (PCFIELDSPLIT full)."""
    rng = np.random.default_rng(5)
    nu_d, npd = 30, 9
    A = rng.standard_normal((nu_d, nu_d)) + 8 * np.eye(nu_d)
    B = sp.csr_matrix(rng.standard_normal((npd, nu_d)))
    Minv = sp.diags(rng.uniform(1, 2, npd)).tocsr()
    bc = np.array([0, 7])
    A[bc, :] = 0.0
    A[:, bc] = 0.0
    A[bc, bc] = 1.0
    Bz = B.toarray().copy()
    Bz[:, bc] = 0.0
    Ai = np.linalg.inv(A)
    nu, gamma = 0.3, 100.0
    Si = -(nu + gamma) * Minv.toarray()
    I, Z = np.eye(nu_d), np.zeros((nu_d, npd))
    P = np.block([[I, -Ai @ Bz.T], [Z.T, np.eye(npd)]]) @ np.block([[Ai, Z], [Z.T, Si]]) @ np.block([[I, Z], [-Bz @ Ai, np.eye(npd)]])
    r = rng.standard_normal(nu_d + npd)
    y = oo.schur_apply(lambda v: Ai @ v, B, Minv, bc, nu, gamma, r, remove_constant=False)
    assert np.linalg.norm(y - P @ r) <= 1e-12 * np.linalg.norm(P @ r)
    yc = oo.schur_apply(lambda v: Ai @ v, B, Minv, bc, nu, gamma, r, remove_constant=True)
    assert abs(yc[nu_d:].mean()) <= 1e-13 * np.abs(yc[nu_d:]).max()
    J = np.block([[A, Bz.T], [Bz, np.zeros((npd, npd))]])
    z = rng.standard_normal(nu_d + npd)
    assert np.linalg.norm(oo.jacobian_apply(sp.csr_matrix(A), B, bc, z) - J @ z) <= 1e-13 * np.linalg.norm(J @ z)


@pytest.fixture(scope="module")
def host_run():
    from oracle.backend import OracleBackend
    s = ContinuationSolver(SMALL2D, OracleBackend(SMALL2D.m))
    return s, [s.solve(re) for re in (10, 100)]


@pytest.mark.parametrize("outer", ["schur", "device"])
def test_continuation_through_the_outer_backend_interface(host_run, outer):
    """The `outer` switch of the continuation stand-in hands the Schur application / the whole linear solve to the
    backend; with the oracle behind that interface nothing may change."""
    s0, i0 = host_run
    s = ContinuationSolver(SMALL2D, oo.OracleOuterBackend(SMALL2D.m), outer=outer)
    infos = [s.solve(re) for re in (10, 100)]
    assert [i["nonlinear_iter"] for i in infos] == [i["nonlinear_iter"] for i in i0]
    assert [i["linear_iter"] for i in infos] == [i["linear_iter"] for i in i0]
    assert np.linalg.norm(s.u - s0.u) <= 1e-8 * np.linalg.norm(s0.u)        # the solves stop at ksp_rtol 1e-9
    assert np.linalg.norm(s.p - s0.p) <= 1e-8 * np.linalg.norm(s0.p)


# ---- GPU ------------------------------------------------------------------------------------------------------------
def _pair(problems, name, gamma, nu):
    from alfi_b200.multigrid import DeviceBackend, level_input_from_synth
    prob = problems(name, gamma=gamma, nu=nu)
    cfg = prob.config
    levels = [level_input_from_synth(l) for l in prob.levels]
    B, Minv = assemble_divergence(prob.finest.V, cfg.k - 1 if cfg.discretisation == "sv" else 0)
    dev, ora = DeviceBackend(cfg.m, deterministic=True), oo.OracleOuterBackend(cfg.m)
    for be in (dev, ora):
        be.setup(levels)
        be.setup_outer(B, Minv, prob.finest.bc_dofs)
    return prob, dev, ora, B


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc3d-sv-k3-tiny", "ldc2d-pkp0-tiny"])
def test_schur_and_jacobian_apply_equal_oracle(problems, name):
    gamma, nu = 10.0, 0.2
    prob, dev, ora, B = _pair(problems, name, gamma, nu)
    rng = np.random.default_rng(20261017)
    n = B.shape[0] + B.shape[1]
    r = rng.standard_normal(n)
    r[prob.finest.bc_dofs] = 0.0
    y, yo = dev.schur_apply(nu, gamma, r), ora.schur_apply(nu, gamma, r)
    assert np.linalg.norm(y - yo) <= 1e-11 * np.linalg.norm(yo)              # tolerance of north-star condition 2
    assert abs(y[B.shape[1]:].mean()) <= 1e-12 * np.abs(y[B.shape[1]:]).max()
    z = rng.standard_normal(n)
    j, jo = dev.jacobian_apply(z), ora.jacobian_apply(z)
    assert np.linalg.norm(j - jo) <= 1e-13 * np.linalg.norm(jo)


@pytest.mark.gpu
@pytest.mark.parametrize("name,restart", [("ldc2d-sv-k2-tiny", 30), ("ldc3d-sv-k3-tiny", 30), ("ldc2d-sv-k2-tiny", 3)])
def test_outer_solve_equals_oracle(problems, name, restart):
    gamma, nu = 10.0, 0.2
    prob, dev, ora, B = _pair(problems, name, gamma, nu)
    rng = np.random.default_rng(7)
    rhs = rng.standard_normal(B.shape[0] + B.shape[1])
    rhs[prob.finest.bc_dofs] = 0.0
    rhs[B.shape[1]:] -= rhs[B.shape[1]:].mean()                              # compatible pressure right-hand side
    x, its, hist = dev.outer_solve(nu, gamma, rhs, 1e-9, 1e-12, 200, restart)
    xo, itso, histo = ora.outer_solve(nu, gamma, rhs, 1e-9, 1e-12, 200, restart)
    assert its == itso and its >= 2 and (restart > 3 or its > restart)
    assert np.allclose(hist, histo, rtol=1e-5, atol=1e-9 * hist[0])
    assert np.linalg.norm(x - xo) <= 1e-8 * np.linalg.norm(xo)
    res = rhs - ora.jacobian_apply(x)
    assert np.linalg.norm(res) <= 2e-9 * np.linalg.norm(rhs)


@pytest.mark.gpu
def test_continuation_with_the_linear_solve_on_the_device(host_run):
    from alfi_b200.multigrid import DeviceBackend
    s0, i0 = host_run
    for outer in ("schur", "device"):
        s = ContinuationSolver(SMALL2D, DeviceBackend(SMALL2D.m, deterministic=True), outer=outer)
        infos = [s.solve(re) for re in (10, 100)]
        assert [i["nonlinear_iter"] for i in infos] == [i["nonlinear_iter"] for i in i0], outer
        for a, b in zip(i0, infos):
            assert abs(a["linear_iter"] - b["linear_iter"]) <= a["nonlinear_iter"], (outer, a, b)   # +-1 per Newton step
        assert np.linalg.norm(s.u - s0.u) <= 1e-8 * np.linalg.norm(s0.u)
        assert np.linalg.norm(s.p - s0.p) <= 1e-8 * np.linalg.norm(s0.p)


@pytest.mark.gpu
def test_al_fieldsplit_pc_equals_oracle(problems):
    """The coarsest-grained drop-in: the whole outer fieldsplit PC as one python PC (alfi_b200.ALFieldsplitPC)."""
    import alfi_b200
    from alfi_b200.synth.fakepetsc import FakePC, FakeVec, SynthAdapter
    gamma, nu = 10.0, 0.2
    prob, _, ora, B = _pair(problems, "ldc3d-sv-k3-tiny", gamma, nu)
    pc = FakePC(prob.finest.level.plex, attrs={"alfi_b200_adapter": SynthAdapter(prob, deterministic=True)})
    p = alfi_b200.ALFieldsplitPC()
    p.setUp(pc)
    n = B.shape[0] + B.shape[1]
    r = np.random.default_rng(11).standard_normal(n)
    r[prob.finest.bc_dofs] = 0.0
    y = FakeVec(n)
    p.apply(pc, FakeVec(r), y)
    want = ora.schur_apply(nu, gamma, r)
    assert np.linalg.norm(y.array - want) <= 1e-11 * np.linalg.norm(want)
    p.setUp(pc)                                                              # second PCSetUp = update()
    p.apply(pc, FakeVec(r), y)
    assert np.linalg.norm(y.array - want) <= 1e-11 * np.linalg.norm(want)
