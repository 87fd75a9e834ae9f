"""Rows T3, T4, T5: the robust transfer's composition, executed by the REFERENCE'S OWN CODE.

oracle/refshim_transfer.py runs `AutoSchoeberlTransfer.restrict_or_prolong` (alfi/transfer.py:194-275) from the
reference tree with its Firedrake calls replaced by stand-ins over the synthetic problem (forms -> the level's
assembled parts, LinearSolver -> PCPATCH semantics with the patch constructor the reference's `patchparams`
name, prolong/restrict -> P_H).  The oracle's `prolong` / `restrict` — and through them the CUDA path, which is
tested against the oracle — must be that sequence.  Needs the reference tree (build container); the committed
fixture tests/golden/reference_transfer.npz carries its outputs to where the tree is absent.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # script mode: regenerate the fixture

from oracle import hotpath as hp  # noqa: E402
from oracle import refshim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "reference_transfer.npz")
CASES = [("ldc2d-sv-k2-tiny", "SVSchoeberlTransfer", "bary"), ("ldc3d-sv-k3-tiny", "SVSchoeberlTransfer", "bary"),
         ("ldc2d-pkp0-tiny", "PkP0SchoeberlTransfer", "uniform"), ("bfs2d-sv-k2-tiny", "SVSchoeberlTransfer", "bary")]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def vectors(lv, l, seed):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal(lv[l - 1].n)
    c[lv[l - 1].bc_dofs] = 0
    f = rng.standard_normal(lv[l].n)
    f[lv[l].bc_dofs] = 0
    return c, f


def run_reference(prob, kind, hierarchy):
    """{key: array}: what the reference's prolong / restrict write for seeded inputs on every level pair."""
    from oracle.refshim_transfer import Harness
    lv = [hp.level_from_host(l) for l in prob.levels]
    h = Harness(prob)
    out = {}
    with h.transfer(kind, hierarchy) as (tr, nu, gamma):
        for l in range(1, len(lv)):
            c, f = vectors(lv, l, 100 + l)
            fine = h.function(l)
            tr.prolong(h.function(l - 1, c), fine)
            out["l%d_prolong" % l] = fine.dat.data.reshape(-1).copy()
            coarse = h.function(l - 1)
            tr.restrict(h.function(l, f), coarse)
            out["l%d_restrict" % l] = coarse.dat.data.reshape(-1).copy()
    return out, h


@pytest.mark.parametrize("name,kind,hierarchy", CASES)
def test_oracle_transfers_are_the_reference_sequence(problems, name, kind, hierarchy):
    """Stored outputs of the reference's code == oracle (the Dirichlet rows are zeroed by the caller)."""
    g = np.load(FIXTURE)
    prob = problems(name, gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in prob.levels]
    for l in range(1, len(lv)):
        c, f = vectors(lv, l, 100 + l)
        ref = g["%s/l%d_prolong" % (name, l)].copy()
        ref[lv[l].bc_dofs] = 0
        assert rel(hp.prolong(lv[l], c), ref) < 1e-12
        ref = g["%s/l%d_restrict" % (name, l)].copy()
        ref[lv[l - 1].bc_dofs] = 0
        assert rel(hp.restrict(lv[l], f, lv[l - 1].bc_dofs), ref) < 1e-12


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,kind,hierarchy", CASES)
def test_fixture_is_what_the_reference_code_produces_now(problems, name, kind, hierarchy):
    g = np.load(FIXTURE)
    out, h = run_reference(problems(name, gamma=10.0, nu=0.2), kind, hierarchy)
    for key, val in out.items():
        assert rel(val, g[name + "/" + key]) < 1e-13, key
    nl = len(problems(name, gamma=10.0, nu=0.2).levels) - 1
    # first call on a level: one operator assembly + one patch setup; afterwards only the right-hand sides
    assert h.assemblies["matrix"] == nl and h.patch_setups == nl and h.assemblies["vector"] == 2 * nl


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_rebuild_only_when_nu_or_gamma_change(problems):
    """transfer.py:173-184, 238-244 (row T5), and the device-side class follows the same rule."""
    from oracle.refshim_transfer import Harness
    prob = problems("ldc2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in prob.levels]
    h = Harness(prob)
    c, f = vectors(lv, 1, 7)
    with h.transfer("SVSchoeberlTransfer", "bary") as (tr, nu, gamma):
        fine = h.function(1)
        tr.prolong(h.function(0, c), fine)
        first = fine.dat.data.copy()
        tr.prolong(h.function(0, c), fine)
        tr.restrict(h.function(1, f), h.function(0))
        assert h.assemblies["matrix"] == 1 and h.patch_setups == 1           # nothing changed: no rebuild
        assert np.array_equal(fine.dat.data, first)
        nu.assign(0.05)                                                       # new Reynolds number
        tr.prolong(h.function(0, c), fine)
        assert h.assemblies["matrix"] == 2 and h.patch_setups == 2
        import dataclasses
        from alfi_b200.synth.problem import assemble_transfer
        ld = dataclasses.replace(prob.levels[1])
        assemble_transfer(prob.config, ld, 0.05, 10.0)
        L = hp.level_from_host(ld)
        want = hp.prolong(L, c)
        got = fine.dat.data.reshape(-1).copy()
        got[L.bc_dofs] = 0
        assert rel(got, want) < 1e-12
        tr.prolong(h.function(0, c), fine)
        assert h.assemblies["matrix"] == 2                                    # and stays built
    assemble_transfer(prob.config, prob.levels[1], 0.2, 10.0)                # leave the cached problem as it was


if __name__ == "__main__":                       # python tests/test_reference_transfer.py : regenerate the fixture
    from alfi_b200.synth.problem import build_problem
    blob = {}
    for name, kind, hierarchy in CASES:
        out, _ = run_reference(build_problem(name, gamma=10.0, nu=0.2), kind, hierarchy)
        blob.update({name + "/" + k: v for k, v in out.items()})
    np.savez_compressed(FIXTURE, **blob)
    print(FIXTURE, os.path.getsize(FIXTURE))
