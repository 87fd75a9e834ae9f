"""GPU parity on BASELINE.json configs[2] (backward-facing step, Gmsh-style unstructured mesh).

First run on a B200 in round 2 (9 passed, profiles/r2_gpu_tests.txt).  The CUDA library is mesh-agnostic; the
CPU side of this configuration is covered by tests/test_bfs.py, tests/test_golden.py and
tests/test_oracle_properties.py.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-11
EPS = np.finfo(np.float64).eps


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _vec(lv, seed):
    x = np.random.default_rng(20261017 + seed).standard_normal(lv.n)
    x[lv.bc_dofs] = 0.0
    return x


@pytest.mark.parametrize("regime", ["mild", "prod"])
@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("condense", [True, False])
def test_bfs_parity(problems, regime, deterministic, condense):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    from oracle import hotpath as hp
    prob = problems("bfs2d-sv-k2-tiny", gamma=10.0, nu=0.2) if regime == "mild" else problems("bfs2d-sv-k2-tiny")
    mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m,
                         deterministic=deterministic, condense=condense)
    olv = [hp.level_from_host(l) for l in prob.levels]
    L, Lc = olv[1], olv[0]
    kappa = max(np.linalg.cond(M) for M in hp.patch_matrices(L.A, L.offsets, L.dofs) if M.size)
    tol = TOL * max(1.0, kappa * EPS / 1e-12)
    if regime == "mild":
        assert tol == TOL
    assert np.array_equal(mg.ctx.colours(1, prob.finest.patches.npatch), prob.finest.patches.colours)
    assert mg.ctx.patch_storage_form(1) == (2 if condense else 0)
    x = _vec(L, 1)
    assert rel(mg.ctx.spmv(1, x, np.empty_like(x)), L.A @ x) <= TOL
    y = mg.ctx.smoother_apply(1, x, np.empty_like(x))
    assert rel(y, hp.smoother_apply(x, L.offsets, L.dofs, L.order, L.factors, L.bc_dofs)) <= tol
    if deterministic:
        assert np.array_equal(y, mg.ctx.smoother_apply(1, x, np.empty_like(x)))
    c, f = _vec(Lc, 2), _vec(L, 3)
    assert rel(mg.ctx.prolong(1, c, np.empty(L.n)), hp.prolong(L, c)) <= tol
    assert rel(mg.ctx.restrict(1, f, np.empty(Lc.n)), hp.restrict(L, f, Lc.bc_dofs)) <= tol
    b = _vec(L, 4)
    assert rel(mg.apply(b, np.empty_like(b)), hp.fcycle(olv, b, prob.config.m)) <= 100 * tol
    mg.ctx.close()


def test_bfs_golden(problems):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bfs2d-sv-k2-tiny.npz"))
    prob = problems("bfs2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m)
    x, c = g["l1_x"], g["l1_c"]
    assert rel(mg.ctx.spmv(1, x, np.empty_like(x)), g["l1_spmv"]) <= TOL
    assert rel(mg.ctx.smoother_apply(1, x, np.empty_like(x)), g["l1_apply"]) <= TOL
    assert rel(mg.ctx.prolong(1, c, np.empty_like(x)), g["l1_prolong"]) <= TOL
    assert rel(mg.ctx.restrict(1, x, np.empty_like(c)), g["l1_restrict"]) <= TOL
    assert rel(mg.apply(g["b"], np.empty_like(g["b"])), g["fcycle"]) <= TOL
    mg.ctx.close()
