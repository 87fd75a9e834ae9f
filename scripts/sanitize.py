"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize.py
Covers factor (all phases incl. DMMA far update: n > 64), apply (atomic + coloured), SpMV, transfers,
FGMRES, coarse solve, graph capture."""
import sys

import numpy as np

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

for name, det in (("ldc3d-sv-k3-tiny", True), ("ldc2d-pkp0-tiny", False), ("ldc2d-sv-k2-tiny", False)):
    prob = build_problem(name, gamma=10.0, nu=0.2)
    mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, deterministic=det)
    b = np.random.default_rng(0).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0
    for _ in range(3):
        x = mg.apply(b, np.empty_like(b))
    print(name, "ok", float(np.linalg.norm(x)))
    mg.ctx.close()
