"""GPU: the coarse level held as one condensed patch (round-2 preparation; written without a GPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc3d-sv-k3-tiny", "ldc3d-sv-k3-small"])
@pytest.mark.parametrize("deterministic", [False, True])
def test_condensed_coarse_solve(problems, monkeypatch, name, deterministic):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    from oracle import hotpath as hp
    prob = problems(name, gamma=10.0, nu=0.2)
    lv0 = hp.level_from_host(prob.levels[0])
    b = np.random.default_rng(0).standard_normal(lv0.n)
    b[lv0.bc_dofs] = 0
    want = hp.coarse_solve(lv0, b)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("ALFIB_COARSE_CONDENSED", flag)
        mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, deterministic=deterministic)
        assert mg.ctx.patch_storage_form(0) == (2 if flag == "1" else 0)
        x = mg.ctx.coarse_solve(b, np.empty_like(b))
        assert rel(x, want) <= 1e-11, (flag, rel(x, want))
        rhs = np.random.default_rng(1).standard_normal(prob.finest.ndofs)
        rhs[prob.finest.bc_dofs] = 0
        out[flag] = mg.apply(rhs, np.empty_like(rhs)).copy()
        mg.ctx.close()
    assert rel(out["1"], out["0"]) <= 1e-11
