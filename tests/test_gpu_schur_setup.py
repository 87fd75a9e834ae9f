"""GPU: ALFIB_SCHUR_SETUP=1 — X_SS of the condensed patch inverses (and of the condensed coarse inverse) from the
Schur complement formed with solves, instead of the S x S cut of the pivoted inverse of the whole patch
(CPU statement: tests/test_condense_host.py::test_schur_setup_*).  Round-2 preparation; written
without a GPU."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name,kw", [("ldc2d-sv-k2-tiny", {}), ("ldc3d-sv-k3-tiny", {}), ("ldc2d-sv-k2", {}),
                                     ("ldc3d-sv-k3-small", {}), ("bfs2d-sv-k2-tiny", {})])
@pytest.mark.parametrize("shared", ["1", "0"])
def test_schur_setup_equals_full_inverse_setup(problems, monkeypatch, name, kw, shared):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    prob = problems(name, **kw)
    levels = [level_input_from_synth(l) for l in prob.levels]
    n = prob.finest.ndofs
    L = len(levels) - 1
    rng = np.random.default_rng(4)
    x = rng.standard_normal(n)
    b = rng.standard_normal(n)
    b[prob.finest.bc_dofs] = 0
    monkeypatch.setenv("ALFIB_CONDENSE_SHARED", shared)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ALFIB_SCHUR_SETUP", flag)
        t0 = time.time()
        mg = DeviceMultigrid(levels, prob.config.m, deterministic=False)
        mg.ctx.synchronize()
        t1 = time.time()
        mg.update_operators(levels)                    # the per-Newton-step setup once more, warm
        mg.ctx.synchronize()
        t2 = time.time()
        assert mg.ctx.patch_storage_form(L) == (2 if shared == "1" else 1)
        y = mg.ctx.smoother_apply(L, x, np.empty(n)).copy()
        z = mg.apply(b, np.empty(n)).copy()
        ps = prob.finest.patches
        p = int(np.argmax(ps.sizes))
        X = mg.ctx.patch_inverse(L, p, int(ps.sizes[p]))
        out[flag] = (y, z, X, t2 - t1)
        mg.ctx.close()
    from oracle import hotpath as hp
    lv = hp.level_from_host(prob.finest)
    mats = hp.patch_matrices(lv.A, lv.offsets, lv.dofs)
    M = mats[p]
    kappa = np.linalg.cond(M)
    tol = max(1e-10, 100 * kappa * np.finfo(float).eps)
    assert rel(out["1"][0], out["0"][0]) <= tol, (rel(out["1"][0], out["0"][0]), kappa)
    assert rel(out["1"][1], out["0"][1]) <= max(1e-9, tol)
    for flag in ("0", "1"):
        X = out[flag][2]
        back = np.linalg.norm(X @ M - np.eye(M.shape[0])) / (np.linalg.norm(X) * np.linalg.norm(M))
        assert back < 1e-9, (flag, back)
    print("%s shared=%s: per-Newton-step setup %.3f s (full inverse) -> %.3f s (Schur), smoother rel diff %.1e, kappa %.1e"
          % (name, shared, out["0"][3], out["1"][3], rel(out["1"][0], out["0"][0]), kappa))
