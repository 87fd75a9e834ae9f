"""oracle/c kernels (CPU baseline) against the numpy oracle."""
import numpy as np

from oracle import cport
from oracle import hotpath as hp


def test_c_kernels_match_numpy(problems):
    prob = problems("ldc2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l, "inverse") for l in prob.levels]
    L = lv[1]
    ck = cport.CLevel(L, prob.levels[1].patches.colours)
    x = np.random.default_rng(0).standard_normal(L.n)
    assert np.allclose(ck.spmv(x), L.A @ x, rtol=0, atol=1e-12 * np.abs(L.A @ x).max())
    want = hp.smoother_apply(x, L.offsets, L.dofs, L.order, L.factors, L.bc_dofs)
    assert np.linalg.norm(ck.smoother_apply(x) - want) <= 1e-13 * np.linalg.norm(want)
    b = x.copy()
    b[L.bc_dofs] = 0
    x0 = hp.fcycle(lv, b, prob.config.m)
    cport.accelerate(lv, [None if l.patches is None else l.patches.colours for l in prob.levels])
    x1 = hp.fcycle(lv, b, prob.config.m)
    assert np.linalg.norm(x1 - x0) <= 1e-11 * np.linalg.norm(x0)
