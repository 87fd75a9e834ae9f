"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize.py [section ...]
Sections (default: all):
  cycle    factor (all phases incl. DMMA far update: n > 64; Schur-complement setup of the condensed inverses), apply
           (TMA tile ops, atomic + coloured), SpMV, transfers, FGMRES, condensed coarse inverse, graph capture + replay
  burman   patch operators with corrections (dense inverses, alfib_level_set_patch_corrections)
  sweeps   multiplicative / symmetrised composition (stage schedule)
  outer    Schur-complement fieldsplit, B / B^T, outer FGMRES (csrc/outer.cu)
Every section prints "<name> ok <norm>"; the numbers are not compared here (the GPU tests do that), the point is the
sanitizer's error summary."""
import dataclasses
import sys

import numpy as np

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceBackend, DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import CONFIGS, build_problem  # noqa: E402

SECTIONS = sys.argv[1:] or ["cycle", "burman", "sweeps", "outer"]


def cycles(prob, det, tag, reps=3):
    mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, deterministic=det)
    b = np.random.default_rng(0).standard_normal(prob.finest.ndofs)
    b[prob.finest.bc_dofs] = 0
    for _ in range(reps):                       # the third application replays the captured graph
        x = mg.apply(b, np.empty_like(b))
    print(tag, "ok", float(np.linalg.norm(x)), flush=True)
    mg.ctx.close()


if "cycle" in SECTIONS:
    for name, det in (("ldc3d-sv-k3-tiny", True), ("ldc2d-pkp0-tiny", False), ("ldc2d-sv-k2-tiny", False)):
        cycles(build_problem(name, gamma=10.0, nu=0.2), det, name)

if "burman" in SECTIONS:
    for name in ("ldc2d-sv-k2-tiny-burman", "ldc3d-sv-k3-tiny-burman"):
        cycles(build_problem(name, gamma=10.0, nu=0.2), True, name, reps=2)

if "sweeps" in SECTIONS:
    for name in ("ldc2d-sv-k2-tiny", "ldc3d-pkp0-tiny"):
        base = CONFIGS[name]
        cfg = dataclasses.replace(base, name=name + "-mult", composition="multiplicative", sort_order=base.sort_order or "0+:1-")
        cycles(build_problem(cfg, gamma=10.0, nu=0.2), False, cfg.name, reps=2)

if "outer" in SECTIONS:
    from alfi_b200.synth.outer import assemble_divergence
    for name in ("ldc2d-sv-k2-tiny",):
        gamma, nu = 10.0, 0.2
        prob = build_problem(name, gamma=gamma, nu=nu)
        cfg = prob.config
        B, Minv = assemble_divergence(prob.finest.V, cfg.k - 1 if cfg.discretisation == "sv" else 0)
        dev = DeviceBackend(cfg.m, deterministic=True)
        dev.setup([level_input_from_synth(l) for l in prob.levels])
        dev.setup_outer(B, Minv, prob.finest.bc_dofs)
        rhs = np.random.default_rng(7).standard_normal(B.shape[0] + B.shape[1])
        rhs[prob.finest.bc_dofs] = 0.0
        rhs[B.shape[1]:] -= rhs[B.shape[1]:].mean()
        y = dev.schur_apply(nu, gamma, rhs)
        x, its, _ = dev.outer_solve(nu, gamma, rhs, 1e-9, 1e-12, 200, 30)
        print(name, "outer ok", float(np.linalg.norm(y)), its, float(np.linalg.norm(x)), flush=True)
