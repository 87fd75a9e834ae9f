"""The fine-grained drop-in of INTEGRATION.md §1, end to end: the reference's OWN `fieldsplit_0` dictionary
(`get_parameters()` executed from the reference tree, fixture tests/golden/reference_parameters.json) with nothing
changed but the string `pc_python_type`, interpreted by a PETSc stand-in (tests/minipetsc.py) that instantiates
`alfi_b200.PatchPC` per level and calls `alfi_b200.SVSchoeberlTransfer` for the transfers — against the CPU oracle
and against the coarse-grained `alfi_b200.VelocityMGPC` (the whole cycle on the device)."""
import copy
import json
import os

import numpy as np
import pytest

import alfi_b200
from alfi_b200.multigrid import level_input_from_synth
from alfi_b200.synth.fakepetsc import FakePC, FakeVec, SynthAdapter
from alfi_b200.synth.problem import assemble_transfer
from alfi_b200.transfer import device_transfer_backend
from oracle import hotpath as hp

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = json.load(open(os.path.join(HERE, "golden", "reference_parameters.json")))


def reference_fieldsplit0(key):
    fs0 = copy.deepcopy(PARAMS[key]["outer"]["fieldsplit_0"])
    assert fs0["mg_levels"]["pc_python_type"] == "firedrake.PatchPC"
    fs0["mg_levels"]["pc_python_type"] = "alfi_b200.PatchPC"            # THE change (solver.py:319)
    return fs0


def test_reference_dictionary_is_consumed_unchanged_on_the_cpu_side():
    fs0 = reference_fieldsplit0("ldc3d-sv-k3")
    cfgd = alfi_b200.pc.fieldsplit0_config(fs0)
    assert cfgd["smoothing"] == 10 and cfgd["construct"] == "alfi.MacroStar" and cfgd["sort_order"] == "0+:1-"


@pytest.mark.gpu
@pytest.mark.parametrize("key,name", [("ldc2d-sv-k2", "ldc2d-sv-k2-tiny"), ("ldc3d-sv-k3", "ldc3d-sv-k3-tiny-literal")])
def test_fine_grained_drop_in_equals_oracle_and_device_cycle(problems, key, name):
    from tests.minipetsc import MiniFieldsplit0
    prob = problems(name, gamma=10.0, nu=0.2)
    fs0 = reference_fieldsplit0(key)
    assert int(fs0["mg_levels"]["ksp_max_it"]) == prob.config.m
    levels = [level_input_from_synth(l) for l in prob.levels]
    nu, gamma = prob.nu, prob.gamma
    params = [nu, gamma]

    def values_for(level, nu_, gamma_):                       # what transfer.py:238-244 re-assembles
        a0, d = assemble_transfer(prob.config, prob.levels[level], nu_, gamma_)
        return a0.vals, d.vals
    transfer = alfi_b200.SVSchoeberlTransfer(params, prob.config.dim, "bary",
                                             **device_transfer_backend(levels, values_for=values_for))
    mini = MiniFieldsplit0(fs0, [l.A.to_csr() for l in prob.levels], [l.bc_dofs for l in prob.levels],
                           [l.level.plex for l in prob.levels],
                           [SynthAdapter(prob, i) for i in range(len(prob.levels))], transfer)
    b = np.random.default_rng(3).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0.0
    got = mini.apply(b)
    want = hp.fcycle([hp.level_from_host(l) for l in prob.levels], b, prob.config.m)
    assert np.linalg.norm(got - want) <= 1e-10 * np.linalg.norm(want)
    # the coarse-grained plugin: the same dictionary read by fieldsplit0_config, the whole cycle on the device
    pc = FakePC(prob.finest.level.plex, attrs={"alfi_b200_adapter": SynthAdapter(prob)})
    whole = alfi_b200.VelocityMGPC()
    whole.setUp(pc)
    x = FakeVec(prob.finest.ndofs)
    whole.apply(pc, FakeVec(b), x)
    assert np.linalg.norm(x.array - got) <= 1e-10 * np.linalg.norm(got)
    mini.setUp()                                              # second PCSetUp (next Newton step): PatchPC.update
    assert np.linalg.norm(mini.apply(b) - want) <= 1e-10 * np.linalg.norm(want)
