"""GPU: ALFIB_FUSE_INDEX=1 folds sep_rhs_kernel / slot_sum_kernel into the source fetch of the X_SS / [D|-Wf] tile
ops (csrc/condense.cu) — same summation order, so a deterministic application must be bitwise unchanged.  Round-2
preparation; written without a GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc3d-sv-k3-tiny", "ldc3d-sv-k3-small", "bfs2d-sv-k2-tiny"])
@pytest.mark.parametrize("shared", ["1", "0"])
def test_fused_index_kernels_are_bitwise_neutral(problems, monkeypatch, name, shared):
    from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth
    prob = problems(name, gamma=10.0, nu=0.2)
    levels = [level_input_from_synth(l) for l in prob.levels]
    monkeypatch.setenv("ALFIB_CONDENSE_SHARED", shared)
    monkeypatch.setenv("ALFIB_FUSE_INDEX", "0")
    mg = DeviceMultigrid(levels, prob.config.m, deterministic=True)
    n = prob.finest.ndofs
    L = len(levels) - 1
    x = np.random.default_rng(8).standard_normal(n)
    y0 = mg.ctx.smoother_apply(L, x, np.empty(n)).copy()
    l0 = mg.ctx.launches
    mg.ctx.smoother_apply(L, x, np.empty(n))
    per_apply = mg.ctx.launches - l0
    monkeypatch.setenv("ALFIB_FUSE_INDEX", "1")
    y1 = mg.ctx.smoother_apply(L, x, np.empty(n)).copy()
    l1 = mg.ctx.launches
    mg.ctx.smoother_apply(L, x, np.empty(n))
    fused = mg.ctx.launches - l1
    assert np.array_equal(y0, y1)
    assert fused == per_apply - (2 if shared == "1" else 1), (per_apply, fused)
    mg.ctx.close()
