"""Newton continuation around the velocity-block PC (north-star condition 3).

CPU: the stand-in outer solver (alfi_b200/synth/outer.py) with the oracle backend converges like
the reference is documented to (few Krylov iterations per Newton step, exactly divergence-free
Scott-Vogelius velocity).  GPU: with the CUDA library as `fieldsplit_0` the Krylov iteration
counts per Newton step match the oracle's within +-1 and the final velocity / pressure agree to
1e-8 relative."""
import dataclasses

import numpy as np
import pytest

from alfi_b200.synth.outer import ContinuationSolver, fgmres_outer
from alfi_b200.synth.problem import CONFIGS
from oracle.backend import OracleBackend

SMALL2D = dataclasses.replace(CONFIGS["ldc2d-sv-k2"], N=4)
RES = (1, 10, 100)


def test_outer_fgmres_solves_a_dense_system():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((60, 60)) + 8 * np.eye(60)
    b = rng.standard_normal(60)
    x, its, hist = fgmres_outer(lambda v: A @ v, lambda v: v / 8.0, b, 1e-12, 0.0, maxit=200, restart=20)
    assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b)
    assert its > 20                      # went through a restart
    assert all(h1 <= h0 * (1 + 1e-12) for h0, h1 in zip(hist, hist[1:]))


@pytest.fixture(scope="module")
def oracle_run():
    s = ContinuationSolver(SMALL2D, OracleBackend(SMALL2D.m))
    infos = [s.solve(re) for re in RES]
    return s, infos


def test_continuation_with_oracle_backend(oracle_run):
    s, infos = oracle_run
    for info in infos:
        assert info["nonlinear_iter"] <= 5
        assert info["linear_iter"] / max(info["nonlinear_iter"], 1) <= 10      # Reynolds-robust
        assert info["residual"] <= max(1e-8, 1e-9 * info["residual0"])         # snes_atol / snes_rtol
    # Scott-Vogelius on the barycentric mesh: the discrete velocity is exactly divergence free
    assert np.linalg.norm(s.B @ s.u.ravel()) <= 1e-12
    assert abs(s.p.mean()) <= 1e-12


@pytest.mark.gpu
def test_iteration_counts_match_on_gpu(oracle_run):
    from alfi_b200.multigrid import DeviceBackend
    so, io = oracle_run
    sd = ContinuationSolver(SMALL2D, DeviceBackend(SMALL2D.m, deterministic=True))
    idev = [sd.solve(re) for re in RES]
    for a, b in zip(io, idev):
        assert a["nonlinear_iter"] == b["nonlinear_iter"], (a, b)
        assert abs(a["linear_iter"] - b["linear_iter"]) <= a["nonlinear_iter"], (a, b)    # +-1 per Newton step
    assert np.linalg.norm(sd.u - so.u) <= 1e-8 * np.linalg.norm(so.u)
    assert np.linalg.norm(sd.p - so.p) <= 1e-8 * max(np.linalg.norm(so.p), 1e-300)


@pytest.mark.gpu
def test_iteration_counts_match_on_gpu_3d():
    from alfi_b200.multigrid import DeviceBackend
    cfg = CONFIGS["ldc3d-sv-k3-tiny"]
    so = ContinuationSolver(cfg, OracleBackend(cfg.m))
    sd = ContinuationSolver(cfg, DeviceBackend(cfg.m, deterministic=True))
    for re in (1, 10):
        a, b = so.solve(re), sd.solve(re)
        assert a["nonlinear_iter"] == b["nonlinear_iter"], (a, b)
        assert abs(a["linear_iter"] - b["linear_iter"]) <= a["nonlinear_iter"], (a, b)
    assert np.linalg.norm(sd.u - so.u) <= 1e-8 * np.linalg.norm(so.u)


FIXTURES = {"ldc3d-sv-k3-small": "continuation_3d_small.npz", "ldc3d-sv-k3-small-burman": "continuation_3d_small_burman.npz"}


def _cont3d(name="ldc3d-sv-k3-small"):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("cont3d", os.path.join(root, "scripts", "cont3d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, os.path.join(root, "tests", "golden", FIXTURES[name])


@pytest.mark.parametrize("name,min_steps", [("ldc3d-sv-k3-small", 30), ("ldc3d-sv-k3-small-burman", 30)])
def test_3d_fixture_follows_the_reference_ladder(name, min_steps):
    """tests/golden/continuation_3d_small*.npz: the CPU oracle with LU patch solves on the reference's Reynolds ladder
    (examples/iters.py:33-37), written by `scripts/cont3d.py oracle` (hours of CPU time; a prefix of the ladder), without
    stabilisation and with the reference's Burman stabilisation (generate_submission:69-87)."""
    mod, path = _cont3d(name)
    ref = np.load(path)
    n = len(ref["re"])
    assert str(ref["config"]) == name and n >= min_steps
    assert [float(r) for r in ref["re"]] == [float(r) for r in mod.LADDER[:n]]
    assert (ref["nonlinear_iter"] <= 5).all() and (ref["residual"] <= 1e-5).all() and (ref["residual"][3:] <= 1e-8).all()
    assert ref["u"].shape == (7957, 3)
    if "u_polished" in ref.files:                     # one more Newton step: the state moves by (stopping tolerance) x |J^-1|
        assert float(ref["polish_re"]) == float(ref["re"][-1])
        # (with Burman the extra solve also moves the wind inside the stabilisation from the previous Reynolds number's
        #  state to this one's — solver.py:270-271 — so it is a slightly different discrete problem: 9e-4 at Re 5000)
        bound = 1e-2 if name.endswith("burman") else 1e-4
        assert np.linalg.norm(ref["u_polished"] - ref["u"]) <= bound * np.linalg.norm(ref["u"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,outer", [("ldc3d-sv-k3-small", "host"), ("ldc3d-sv-k3-small", "device"),
                                        ("ldc3d-sv-k3-small-burman", "host")])
def test_3d_continuation_against_the_lu_fixture(name, outer):
    """North-star condition 3 on the 3-D Scott-Vogelius k = 3 family: the first Reynolds numbers of the ladder here (the
    whole ladder: bench.py `continuation.three_d` / `three_d_burman`, scripts/cont3d.py) — identical Newton counts, Krylov
    counts within +-1 per Newton step of the oracle that SOLVES with LU factors where the device applies explicit
    (condensed; dense with Burman's patch corrections) inverses."""
    mod, path = _cont3d(name)
    out = mod.compare_with_fixture(name, path, outer, log=lambda *a: None, max_steps=6)
    assert out["newton_counts_equal"] and out["krylov_counts_within_1_per_newton_step"], out
    assert out["re_max"] == 400.0


def test_state_verdict_of_the_3d_comparison():
    """scripts/cont3d.py state_verdict: the bar is the north star's 1e-8 unless the CPU-vs-CPU floor of the fixture is
    above a third of it; either the states as stopped or the polished states must meet it."""
    mod, _ = _cont3d()
    v = mod.state_verdict(8.3e-9, 9.3e-9, 1.03e-8, 1.23e-8, 2.6e-11)          # unstabilised, Re 2900 (call 22)
    assert v["state_bar"] == 1e-8 and v["state_ok_as_stopped"] and not v["state_ok_polished"] and v["state_ok"]
    v = mod.state_verdict(1.15e-7, 1.21e-7, 1.72e-8, 2.48e-8, 1.0e-8)         # Burman, Re 5000 (call 22)
    assert abs(v["state_bar"] - 3e-8) < 1e-20 and not v["state_ok_as_stopped"] and v["state_ok_polished"] and v["state_ok"]
    v = mod.state_verdict(1e-6, 1e-6, 1e-7, 1e-7, 1e-8)
    assert not v["state_ok"]
    v = mod.state_verdict(1e-6, 1e-6, None, None, 0.0)
    assert not v["state_ok"] and not v["state_ok_polished"]


def test_free_running_newton_count_of_a_residual_history():
    """scripts/cont3d.py natural_newton_count: the device run follows the oracle's Newton counts and records what its own
    stopping test would have done."""
    mod, _ = _cont3d()
    assert mod.natural_newton_count([1e-3, 1e-6, 5e-9], 1e-8) == (2, None)                 # stops where the history ends
    assert mod.natural_newton_count([1e-3, 1e-6, 0.987e-8, 4e-11], 1e-8) == (2, 0.987e-8)   # would have stopped a step early
    assert mod.natural_newton_count([1e-3, 1e-6, 1.004e-8], 1e-8) == (3, 1.004e-8)          # would have gone on
    assert abs(0.987e-8 / 1e-8 - 1.0) <= mod.KNIFE_EDGE and abs(1.004e-8 / 1e-8 - 1.0) <= mod.KNIFE_EDGE
