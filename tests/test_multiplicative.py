"""Multiplicative patch composition (`--patch-composition multiplicative`, alfi/solver.py:306-308,322-335: sequential
sweep in the problem's relaxation direction, symmetrised): the stage schedule that lets the device run the sequential
sweep of PCApply_PATCH, on the CPU; the CUDA path against the sequential oracle, on the GPU."""
import copy
import dataclasses
import json
import os

import numpy as np
import pytest

from alfi_b200.patches import sweep_stages
from alfi_b200.synth.problem import CONFIGS, build_problem
from oracle import hotpath as hp

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny", "ldc3d-pkp0-tiny"]


def mult_problem(name, **kw):
    base = CONFIGS[name]
    cfg = dataclasses.replace(base, name=name + "-mult", composition="multiplicative", sort_order=base.sort_order or "0+:1-")
    return build_problem(cfg, **kw)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", CASES)
def test_stage_schedule_reproduces_the_sequential_sweep(name):
    prob = mult_problem(name, gamma=10.0, nu=0.2)
    fine = prob.finest
    ps = fine.patches
    stages = ps.stages
    assert stages is not None and ps.symmetrise and stages.shape == ps.order.shape
    A = fine.A.to_csr()
    # coupled visits lie in increasing stages; visits of one stage are uncoupled
    import scipy.sparse as sp
    bs = fine.V.bs
    An = sp.csr_matrix((np.ones(fine.A.colidx.size), fine.A.colidx, fine.A.rowptr), shape=(fine.V.nnodes,) * 2)
    pat = sp.kron(An, np.ones((bs, bs)), format="csr")         # structural pattern of the BAIJ operator
    for k, p in enumerate(ps.order):
        Ip = ps.patch(p)
        reads = np.unique(pat[Ip].indices)                     # dofs the residual of visit k reads
        for k2 in range(k):
            if np.intersect1d(reads, ps.patch(ps.order[k2])).size:
                assert stages[k2] < stages[k], (k2, k)
    lv = hp.level_from_host(fine)
    x = np.random.default_rng(5).standard_normal(fine.ndofs)
    for symmetric in (False, True):
        want = hp.smoother_apply_multiplicative(A, x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs, symmetric)
        # stage by stage, ONE residual per stage (what the device does)
        y = np.zeros_like(x)
        order_of = [np.flatnonzero(stages == s) for s in range(stages.max() + 1)]
        for seq in ([order_of] + ([order_of[::-1]] if symmetric else [])):
            for visits in seq:
                r = x - A @ y
                for k in visits:
                    I = ps.patch(ps.order[k])
                    y[I] += lv.factors[ps.order[k]][1] @ r[I]
        y[lv.bc_dofs] = x[lv.bc_dofs]
        assert rel(y, want) <= 1e-13
    # a sweep is not the additive sum
    assert rel(want, hp.smoother_apply(x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)) > 1e-3


def test_reference_multiplicative_dictionaries_are_accepted():
    import alfi_b200
    params = json.load(open(os.path.join(HERE, "golden", "reference_parameters.json")))
    for key in ("ldc2d-pkp0-star-multiplicative", "ldc3d-sv-k3-multiplicative"):
        fs0 = copy.deepcopy(params[key]["outer"]["fieldsplit_0"])
        assert fs0["mg_levels"]["patch_pc_patch_local_type"] == "multiplicative"
        got = alfi_b200.pc.fieldsplit0_config(fs0)
        assert got["local_type"] == "multiplicative" and got["symmetrise_sweep"] is True
        assert got["construct"] in ("alfi.Star", "alfi.MacroStar") and got["sort_order"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_sweep_equals_the_sequential_oracle(name):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    prob = mult_problem(name, gamma=10.0, nu=0.2)
    fine = prob.finest
    L = len(prob.levels) - 1
    levels = [level_input_from_synth(l) for l in prob.levels]
    assert levels[L].patch_stages is not None and levels[L].symmetrise_sweep
    mg = DeviceMultigrid(levels, prob.config.m)
    assert mg.ctx.patch_storage_form(L) == 0                   # dense inverses under a sweep
    olv = [hp.level_from_host(l) for l in prob.levels]
    n = fine.ndofs
    x = np.random.default_rng(6).standard_normal(n)
    lv = olv[L]
    want = hp.smoother_apply_multiplicative(lv.A, x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs, True)
    got = mg.ctx.smoother_apply(L, x, np.empty(n)).copy()
    assert rel(got, want) <= 1e-11, rel(got, want)
    # forward sweep only, then back to additive: the switches really switch
    mg.ctx.set_sweep_stages(L, levels[L].patch_stages, False)
    want_f = hp.smoother_apply_multiplicative(lv.A, x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs, False)
    assert rel(mg.ctx.smoother_apply(L, x, np.empty(n)), want_f) <= 1e-11
    mg.ctx.set_sweep_stages(L, None)
    want_a = hp.smoother_apply(x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)
    assert rel(mg.ctx.smoother_apply(L, x, np.empty(n)), want_a) <= 1e-11
    # FGMRES(m) smoother and the F-cycle with the symmetrised sweep on every level
    mg.ctx.set_sweep_stages(L, levels[L].patch_stages, True)

    def mult_smooth(lvl, b, x0):
        o = olv[lvl]
        return hp.fgmres(lambda v: o.A @ v, lambda v: hp.smoother_apply_multiplicative(
            o.A, v, o.offsets, o.dofs, o.order, o.factors, o.bc_dofs, True), b, x0, prob.config.m)
    b = np.random.default_rng(7).standard_normal(n)
    b[fine.bc_dofs] = 0.0
    assert rel(mg.ctx.smooth(L, prob.config.m, b, np.zeros(n)), mult_smooth(L, b, np.zeros(n))) <= 1e-10
    # a wrong schedule (two coupled visits in one stage) is an error, never a wrong answer
    bad = levels[L].patch_stages.copy()
    bad[:] = 0
    with pytest.raises(RuntimeError):
        mg.ctx.set_sweep_stages(L, bad, True)
    mg.ctx.close()


@pytest.mark.gpu
def test_patchpc_with_the_reference_multiplicative_dictionary(problems):
    """`alfi_b200.PatchPC` under the mg_levels options of the reference's multiplicative run (Star python constructor
    with the relaxation direction as sort order, symmetrise_sweep)."""
    import alfi_b200
    from alfi_b200.synth.fakepetsc import FakePC, FakeVec, SynthAdapter
    params = json.load(open(os.path.join(HERE, "golden", "reference_parameters.json")))
    lvopts = dict(params["ldc2d-pkp0-star-multiplicative"]["outer"]["fieldsplit_0"]["mg_levels"])
    prob = problems("ldc2d-pkp0-tiny", gamma=10.0, nu=0.2)
    ld = prob.levels[2]
    pc = FakePC(ld.level.plex, options=lvopts, attrs={"alfi_b200_adapter": SynthAdapter(prob, 2)})
    p = alfi_b200.PatchPC()
    p.setUp(pc)
    assert p.local_type == "multiplicative" and p.symmetrise
    A = ld.A.to_csr()
    ps = p.patches
    facs = hp.factor_patches(hp.patch_matrices(A, ps.offsets, ps.dofs))
    x = np.random.default_rng(8).standard_normal(ld.ndofs)
    want = hp.smoother_apply_multiplicative(A, x, ps.offsets, ps.dofs, ps.order, facs, ld.bc_dofs, True)
    y = FakeVec(ld.ndofs)
    p.apply(pc, FakeVec(x), y)
    assert rel(y.array, want) <= 1e-11
