"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU part: the host index-set builders reproduce the stored integer data exactly and the oracle
reproduces its stored float outputs.  GPU part: the CUDA path reproduces the stored vectors."""
import os

import numpy as np
import pytest

from oracle import hotpath as hp

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny", "ldc3d-pkp0-tiny"]
# the backward-facing step (BASELINE configs[2]) joined after the last GPU run of round 1: its CUDA check
# lives in tests/test_gpu_bfs.py until it has been run once on a B200
CPU_NAMES = NAMES + ["bfs2d-sv-k2-tiny"]
TOL = 1e-11


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", CPU_NAMES)
def test_index_sets_bit_exact(problems, name):
    g = load(name)
    prob = problems(name, gamma=10.0, nu=0.2)
    for l, ld in enumerate(prob.levels):
        if ld.patches is None:
            continue
        for key, val in (("offsets", ld.patches.offsets), ("dofs", ld.patches.dofs), ("order", ld.patches.order),
                         ("colours", ld.patches.colours), ("cell_offsets", ld.cell_patches.offsets),
                         ("cell_dofs", ld.cell_patches.dofs), ("cb_dofs", ld.cb_dofs)):
            assert np.array_equal(g["l%d_%s" % (l, key)], val), (l, key)


@pytest.mark.parametrize("name", CPU_NAMES)
def test_oracle_reproduces_golden(problems, name):
    g = load(name)
    prob = problems(name, gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in prob.levels]
    for l, L in enumerate(lv):
        if L.offsets is None:
            continue
        x, c = g["l%d_x" % l], g["l%d_c" % l]
        assert rel(L.A @ x, g["l%d_spmv" % l]) <= 1e-13
        assert rel(hp.smoother_apply(x, L.offsets, L.dofs, L.order, L.factors, L.bc_dofs), g["l%d_apply" % l]) <= TOL
        assert rel(hp.prolong(L, c), g["l%d_prolong" % l]) <= TOL
        assert rel(hp.restrict(L, x, lv[l - 1].bc_dofs), g["l%d_restrict" % l]) <= TOL
        assert rel(hp.smooth(L, x, np.zeros(L.n), prob.config.m), g["l%d_smooth" % l]) <= TOL
    assert rel(hp.fcycle(lv, g["b"], prob.config.m), g["fcycle"]) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_golden(problems, name):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    g = load(name)
    prob = problems(name, gamma=10.0, nu=0.2)
    mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m)
    for l, ld in enumerate(prob.levels):
        if ld.patches is None:
            continue
        x, c = g["l%d_x" % l], g["l%d_c" % l]
        n, nc = ld.ndofs, prob.levels[l - 1].ndofs
        assert rel(mg.ctx.spmv(l, x, np.empty(n)), g["l%d_spmv" % l]) <= TOL
        assert rel(mg.ctx.smoother_apply(l, x, np.empty(n)), g["l%d_apply" % l]) <= TOL
        assert rel(mg.ctx.prolong(l, c, np.empty(n)), g["l%d_prolong" % l]) <= TOL
        assert rel(mg.ctx.restrict(l, x, np.empty(nc)), g["l%d_restrict" % l]) <= TOL
        assert rel(mg.ctx.smooth(l, prob.config.m, x, np.zeros(n)), g["l%d_smooth" % l]) <= TOL
        assert np.array_equal(mg.ctx.colours(l, ld.patches.npatch), g["l%d_colours" % l])
    assert rel(mg.apply(g["b"], np.empty_like(g["b"])), g["fcycle"]) <= TOL
