"""GPU parity on the reference's LITERAL 3-D MacroStar (alfi/relaxation.py:168-177: 2 175-dof interior patches whose
separators reach 735 dofs and whose macro-cell blocks are cut differently by different patches, so they are not shared
between patches): index sets and colouring bit-exact, smoother application / FGMRES smoother / F-cycle against the CPU
oracle, condensed form against dense inverses.  The benchmark's default sets are the open macro stars
(`macro_expand="vertices"`, an extension); this file is the device coverage of what the reference itself does."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _vec(n, bc, seed):
    x = np.random.default_rng(20261017 + seed).standard_normal(n)
    x[bc] = 0.0
    return x


@pytest.mark.parametrize("name,kw", [("ldc3d-sv-k3-tiny-literal", dict(gamma=10.0, nu=0.2)), ("ldc3d-sv-k3-tiny-literal", {}),
                                     ("ldc3d-sv-k3-small-literal", {})])
@pytest.mark.parametrize("deterministic", [False, True])
def test_literal_macrostar_against_the_oracle(problems, name, kw, deterministic):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    from oracle import hotpath as hp
    prob = problems(name, **kw)
    fine = prob.finest
    L = len(prob.levels) - 1
    ps = fine.patches
    if "small" in name:
        assert int(ps.sizes.max()) == 2175 and ps.colours.max() + 1 >= 20       # the finding of DESIGN §1
    levels = [level_input_from_synth(l) for l in prob.levels]
    olv = [hp.level_from_host(l) for l in prob.levels]
    mats = hp.patch_matrices(olv[L].A, ps.offsets, ps.dofs)
    kappa = max(np.linalg.cond(M) for M in mats if M.size)
    tol = 1e-11 * max(1.0, kappa * EPS / 1e-12)
    n = fine.ndofs
    x, b = _vec(n, fine.bc_dofs, 1), _vec(n, fine.bc_dofs, 2)
    want_apply = hp.smoother_apply(x, olv[L].offsets, olv[L].dofs, olv[L].order, olv[L].factors, olv[L].bc_dofs)
    want_smooth = hp.smooth(olv[L], b, np.zeros(n), prob.config.m)
    want_cycle = hp.fcycle(olv, b, prob.config.m)
    got = {}
    for condense in (True, False):
        mg = DeviceMultigrid(levels, prob.config.m, deterministic=deterministic, condense=condense)
        assert np.array_equal(mg.ctx.colours(L, ps.npatch), ps.colours)
        assert mg.ctx.patch_storage_form(L) == (1 if condense else 0)         # condensed, blocks per (patch, block)
        y = mg.ctx.smoother_apply(L, x, np.empty(n)).copy()
        assert rel(y, want_apply) <= tol, (condense, rel(y, want_apply), kappa)
        s = mg.ctx.smooth(L, prob.config.m, b, np.zeros(n)).copy()
        assert rel(s, want_smooth) <= 100 * tol, (condense, rel(s, want_smooth))
        z = mg.apply(b, np.empty(n)).copy()
        assert rel(z, want_cycle) <= 100 * tol, (condense, rel(z, want_cycle))
        p = int(np.argmax(ps.sizes))
        X = mg.ctx.patch_inverse(L, p, int(ps.sizes[p]))
        back = np.linalg.norm(X @ mats[p] - np.eye(mats[p].shape[0])) / (np.linalg.norm(X) * np.linalg.norm(mats[p]))
        assert back < (1e-9 if condense else 100 * EPS), (condense, back)
        got[condense] = y
        mg.ctx.close()
    assert rel(got[True], got[False]) <= tol
