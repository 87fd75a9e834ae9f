"""GPU, one rank: the distributed-vector code path (alfib_level_set_halo) with a single owner — no peers, no
ghosts, but the halo branches of SpMV / patch apply / FGMRES / transfers, the local-numbering hand-over and the
dof-level local P_H all run.  The multi-rank check is scripts/dist_check_halo.py (needs N GPUs).  Round-2
preparation; written without a GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny", "ldc3d-pkp0-tiny"])
def test_single_rank_halo_path_equals_oracle(problems, name):
    from alfi_b200.multigrid import DistributedMultigrid, level_input_from_synth
    from oracle import hotpath as hp
    prob = problems(name, gamma=10.0, nu=0.2)
    levels = [level_input_from_synth(l) for l in prob.levels]
    mg = DistributedMultigrid(levels, prob.config.m, 0, 1, None, deterministic=True)
    olv = [hp.level_from_host(l) for l in prob.levels]
    n = prob.finest.ndofs
    b = np.random.default_rng(3).standard_normal(n)
    b[prob.finest.bc_dofs] = 0
    loc = mg.local_dofs
    assert mg.n_owned == n and np.array_equal(np.sort(loc), np.arange(n))
    L = len(levels) - 1
    lv = olv[L]
    y = mg.ctx.smoother_apply(L, mg.scatter(b), np.empty(n))
    want = hp.smoother_apply(b, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)
    assert rel(y, want[loc]) <= 1e-11
    x = mg.apply(mg.scatter(b), np.empty(n))
    want = hp.fcycle(olv, b, prob.config.m)
    assert rel(x, want[loc]) <= 1e-9, rel(x, want[loc])
    x2 = x.copy()
    for _ in range(3):                                  # CUDA-graph replay from the third application on
        x2 = mg.apply(mg.scatter(b), np.empty(n))
    assert rel(x2, x) <= 1e-13
    mg.ctx.close()
