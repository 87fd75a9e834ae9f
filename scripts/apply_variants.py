"""Finest-level condensed PCApply_PATCH: tile-op variants on one problem, one process, CUDA events.

    python scripts/apply_variants.py [config] [reps]
Variants are selected by the run-time switches of csrc/condense.cu (read at every launch): v2 (LDG stream),
v2 + fused index kernels, TMA (tile_tma.cuh), TMA + fused index kernels.  Every variant's result is compared with
v2's (relative 2-norm); GB/s are against the algorithmic bytes of the application."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d-sv-k3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
prob = build_problem(name)
L = len(prob.levels) - 1
n = prob.finest.ndofs
torch.manual_seed(1)
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, condense=True)
mg.ctx.synchronize()
stream = torch.cuda.ExternalStream(mg.ctx.stream)
nbytes = mg.ctx.patch_apply_bytes(L)
variants = [("v2", {"ALFIB_TILE_TMA": "0"}),
            ("tma 8w x 8KB", {"ALFIB_TILE_TMA": "1", "ALFIB_TILE_TMA_CFG": "0"}),
            ("tma 16w x 4KB", {"ALFIB_TILE_TMA": "1", "ALFIB_TILE_TMA_CFG": "1"}),
            ("tma 16w x 4KB + RED", {"ALFIB_TILE_TMA": "1", "ALFIB_TILE_TMA_CFG": "2"}),
            ("tma 8w x 8KB + RED", {"ALFIB_TILE_TMA": "1", "ALFIB_TILE_TMA_CFG": "3"})]
ref = None
rows = []
for label, env in variants:
    for k in ("ALFIB_FUSE_INDEX", "ALFIB_TILE_TMA", "ALFIB_TILE_TMA_CFG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for lvl in range(1, L + 1):        # every level once (smaller op lists, boundary-only patches)
        xl = torch.randn(prob.levels[lvl].ndofs, dtype=torch.float64, device="cuda")
        yl = torch.empty_like(xl)
        mg.ctx.smoother_apply(lvl, xl, yl)
    mg.ctx.smoother_apply(L, x, y)
    mg.ctx.synchronize()
    if ref is None:
        ref = y.clone()
    err = float((y - ref).norm() / ref.norm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        mg.ctx.smoother_apply(L, x, y)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    row = {"variant": label, "ms": ms, "GBs": nbytes / ms / 1e6, "frac_of_6457": nbytes / ms / 1e6 / 6457.4,
           "rel_diff_vs_v2": err}
    rows.append(row)
    print(json.dumps(row), flush=True)
print(json.dumps({"config": name, "apply_bytes": nbytes, "rows": rows}))
