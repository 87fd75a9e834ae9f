"""The python-PC plugins against a fake petsc4py protocol.

CPU part: option handling and the index sets a PatchPC hands to the library, with the CUDA
context replaced by a recorder.  GPU part: PatchPC / VelocityMGPC applied through the fake PC
objects equal the oracle."""
import numpy as np
import pytest

import alfi_b200
from alfi_b200 import pc as pcmod
from alfi_b200.synth.fakepetsc import FakePC, FakeVec, SynthAdapter
from oracle import hotpath as hp

# the mg_levels dictionary of alfi/solver.py:313-328 + 339-342 + 655-659, as PETSc would expose it
# under the level's prefix
LEVEL_OPTS = {
    "patch_pc_patch_save_operators": True,
    "patch_pc_patch_partition_of_unity": False,
    "patch_pc_patch_local_type": "additive",
    "patch_pc_patch_statistics": False,
    "patch_pc_patch_symmetrise_sweep": False,
    "patch_pc_patch_precompute_element_tensors": True,
    "patch_sub_ksp_type": "preonly",
    "patch_sub_pc_type": "lu",
    "patch_pc_patch_construct_type": "python",
    "patch_pc_patch_construct_python_type": "alfi.MacroStar",
    "patch_pc_patch_construction_MacroStar_sort_order": "0+:1-",
    "patch_pc_patch_construction_MacroStar_expand": "vertices",
    "patch_pc_patch_sub_mat_type": "seqaij",
    "patch_sub_pc_factor_mat_solver_type": "petsc",
}


class Recorder:
    def __init__(self, *a, **k):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append((name, a, k))
            return 0
        return f


def make_pc(problems, name, level, opts, **kw):
    prob = problems(name, gamma=10.0, nu=0.2)
    ad = SynthAdapter(prob, level, **kw)
    return prob, FakePC(prob.levels[level].level.plex, options=dict(opts), attrs={"alfi_b200_adapter": ad})


def test_drop_in_names_exist():
    for name in ("Star", "MacroStar", "CoarseCellPatches", "CoarseCellMacroPatches", "SVSchoeberlTransfer",
                 "PkP0SchoeberlTransfer", "NullTransfer", "PatchPC", "VelocityMGPC"):
        assert hasattr(alfi_b200, name)
    for meth in ("initialize", "update", "apply", "applyTranspose"):
        assert callable(getattr(alfi_b200.PatchPC, meth))


def test_patchpc_builds_the_reference_index_sets(problems, monkeypatch):
    monkeypatch.setattr(pcmod, "Context", Recorder)
    prob, pc = make_pc(problems, "ldc2d-sv-k2-tiny", 1, LEVEL_OPTS)
    p = alfi_b200.PatchPC()
    p.initialize(pc)
    ps, ref = p.patches, prob.levels[1].patches
    assert np.array_equal(ps.offsets, ref.offsets) and np.array_equal(ps.dofs, ref.dofs)
    # iteration set follows the sort order "0+:1-" (relaxation.py:88-108)
    plex = prob.levels[1].level.plex
    ents = np.flatnonzero(plex.labels["MacroVertices"][plex.vStart:plex.vEnd] == 1) + plex.vStart
    coords = np.array([plex.point_coords(e) for e in ents])
    want = [i for i, _ in sorted(enumerate(coords), key=lambda z: (z[1][0], -z[1][1]))]
    assert ps.order.tolist() == want
    names = [c[0] for c in p.ctx.calls]
    assert names == ["level_create", "set_bsr_pattern", "set_bc", "set_patches", "set_bsr_values", "factor"]
    p.update(pc)
    assert [c[0] for c in p.ctx.calls][-2:] == ["set_bsr_values", "factor"]
    with pytest.raises(NotImplementedError):
        p.applyTranspose(pc, None, None)


def test_patchpc_builtin_star_and_rejections(problems, monkeypatch):
    monkeypatch.setattr(pcmod, "Context", Recorder)
    opts = dict(LEVEL_OPTS)
    opts.update({"patch_pc_patch_construct_type": "star", "patch_pc_patch_construct_dim": 0})
    prob, pc = make_pc(problems, "ldc2d-pkp0-tiny", 2, opts)
    p = alfi_b200.PatchPC()
    p.initialize(pc)
    assert np.array_equal(p.patches.dofs, prob.levels[2].patches.dofs)
    for key, val in (("patch_pc_patch_partition_of_unity", True), ("patch_pc_patch_local_type", "symmetric_multiplicative"),
                     ("patch_sub_pc_type", "ilu")):
        bad = dict(opts)
        bad[key] = val
        _, pc2 = make_pc(problems, "ldc2d-pkp0-tiny", 2, bad)
        with pytest.raises(NotImplementedError):
            alfi_b200.PatchPC().initialize(pc2)
    # multiplicative composition: the sweep stages are handed over between the patches and the values
    mult = dict(opts)
    mult.update({"patch_pc_patch_local_type": "multiplicative", "patch_pc_patch_symmetrise_sweep": True})
    _, pc3 = make_pc(problems, "ldc2d-pkp0-tiny", 2, mult)
    pm = alfi_b200.PatchPC()
    pm.initialize(pc3)
    names = [c[0] for c in pm.ctx.calls]
    assert names == ["level_create", "set_bsr_pattern", "set_bc", "set_patches", "set_sweep_stages", "set_bsr_values", "factor"]
    assert pm.symmetrise and pm.stages.size == pm.patches.order.size


def test_transfer_rebuild_logic():
    """rebuild-on-parameter-change of transfer.py:173-184, 238-244."""
    class Const:
        def __init__(self, v):
            self.v = v

        def __float__(self):
            return float(self.v)

    class Backend:
        def __init__(self):
            self.updates, self.calls = 0, []

        def transfer_update(self, level, a0, d):
            self.updates += 1

        def prolong(self, level, c, f):
            self.calls.append(("prolong", level))

        def restrict(self, level, f, c):
            self.calls.append(("restrict", level))

        def level_sizes(self):
            return {0: 4, 1: 10}

    nu, gamma, be = Const(1.0), Const(1e4), Backend()
    t = alfi_b200.SVSchoeberlTransfer((nu, gamma), 2, "bary", backend=be, values_for=lambda l, n, g: (None, None))
    assert t.patch_constructor is alfi_b200.CoarseCellMacroPatches
    fine, coarse = np.zeros(10), np.zeros(4)
    t.prolong(coarse, fine)
    t.restrict(fine, coarse)
    assert be.updates == 1 and be.calls == [("prolong", 1), ("restrict", 1)]
    nu.v = 0.5                                     # new Reynolds number
    t.prolong(coarse, fine)
    assert be.updates == 2
    t.prolong(coarse, fine)
    assert be.updates == 2
    t.force_rebuild()
    t.restrict(fine, coarse)
    assert be.updates == 3
    dst = np.zeros(3)
    alfi_b200.NullTransfer().inject(None, dst)
    assert np.isnan(dst).all()


@pytest.mark.gpu
def test_patchpc_apply_equals_oracle(problems):
    prob, pc = make_pc(problems, "ldc2d-sv-k2-tiny", 1, {k: v for k, v in LEVEL_OPTS.items() if "sort_order" not in k})
    p = alfi_b200.PatchPC()
    p.setUp(pc)
    lv = hp.level_from_host(prob.levels[1])
    x = FakeVec(np.random.default_rng(0).standard_normal(lv.n))
    y = FakeVec(lv.n)
    p.apply(pc, x, y)
    want = hp.smoother_apply(x.array, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)
    assert np.linalg.norm(y.array - want) <= 1e-11 * np.linalg.norm(want)
    p.setUp(pc)                                   # second PCSetUp = update()
    p.apply(pc, x, y)
    assert np.linalg.norm(y.array - want) <= 1e-11 * np.linalg.norm(want)


@pytest.mark.gpu
def test_velocity_mg_pc_equals_oracle(problems):
    prob = problems("ldc3d-sv-k3-tiny", gamma=10.0, nu=0.2)
    ad = SynthAdapter(prob)
    pc = FakePC(prob.finest.level.plex, attrs={"alfi_b200_adapter": ad})
    p = alfi_b200.VelocityMGPC()
    p.setUp(pc)
    olv = [hp.level_from_host(l) for l in prob.levels]
    b = np.random.default_rng(1).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0
    x = FakeVec(prob.finest.ndofs)
    p.apply(pc, FakeVec(b), x)
    want = hp.fcycle(olv, b, prob.config.m)
    assert np.linalg.norm(x.array - want) <= 1e-11 * np.linalg.norm(want)
    p.setUp(pc)
    p.apply(pc, FakeVec(b), x)
    assert np.linalg.norm(x.array - want) <= 1e-11 * np.linalg.norm(want)
