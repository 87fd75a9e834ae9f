"""Small driver for ncu: set up a config and run a few smoother applications / SpMVs / cycles.

    ncu --set full --clock-control none --import-source on -k regex:patch_apply -s 8 -c 3 \
        -o gpurun_out/prof_apply python scripts/profile_apply.py ldc3d-sv-k3-half apply 6
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d-sv-k3-half"
what = sys.argv[2] if len(sys.argv) > 2 else "apply"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
det = int(sys.argv[4]) if len(sys.argv) > 4 else 0
prob = build_problem(name)
mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, deterministic=bool(det))
L = len(prob.levels) - 1
n = prob.finest.ndofs
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
torch.cuda.synchronize()
for _ in range(reps):
    if what == "apply":
        mg.ctx.smoother_apply(L, x, y)
    elif what == "spmv":
        mg.ctx.spmv(L, x, y)
    elif what == "cycle":
        mg.apply(x, y)
    elif what == "factor":
        mg.ctx.factor(L)
mg.ctx.synchronize()
print("done", name, what, reps, float(y.abs().max()))
