"""Time the individual hot-path operations with CUDA events on the library stream.

    python scripts/kernel_bench.py [config] [reps] [condense=1]
Prints ms per call and achieved GB/s against the algorithmic bytes of SURVEY §8d."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d-sv-k3-half"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
condense = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
prob = build_problem(name)
t0 = time.time()
mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, condense=condense)
mg.ctx.synchronize()
print("setup %.2fs" % (time.time() - t0))
stream = torch.cuda.ExternalStream(mg.ctx.stream)
L = len(prob.levels) - 1
fine = prob.finest
n, nc = fine.ndofs, prob.levels[L - 1].ndofs
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
xc = torch.randn(nc, dtype=torch.float64, device="cuda")
yc = torch.empty_like(xc)
x0 = torch.randn(prob.levels[0].ndofs, dtype=torch.float64, device="cuda")
y0 = torch.empty_like(x0)
bs = fine.V.bs
sizes = fine.patches.sizes.astype(float)
csz = fine.cell_patches.sizes.astype(float)
cell_bytes = mg.ctx.patch_apply_bytes(L, 1) - 16 * n        # factors + indices of the transfer's cell patches
bytes_ = {
    "apply": mg.ctx.patch_apply_bytes(L),
    "spmv": fine.A.nnzb * (8 * bs * bs + 4) + 4 * (fine.V.nnodes + 1) + 16 * n,
    "prolong": 12 * fine.P.nnz + 12 * fine.A.nnzb * bs * bs + cell_bytes + 8 * nc + 24 * n,
    "restrict": 12 * fine.P.nnz + 12 * fine.A.nnzb * bs * bs + cell_bytes + 8 * nc + 24 * n,
    "coarse": 8.0 * prob.levels[0].ndofs ** 2,
    "smooth": None, "cycle": None, "factor": None,
}
ops = {
    "apply": lambda: mg.ctx.smoother_apply(L, x, y),
    "spmv": lambda: mg.ctx.spmv(L, x, y),
    "prolong": lambda: mg.ctx.prolong(L, xc, y),
    "restrict": lambda: mg.ctx.restrict(L, x, yc),
    "coarse": lambda: mg.ctx.coarse_solve(x0, y0),
    "smooth": lambda: mg.ctx.smooth(L, prob.config.m, x, y),
    "cycle": lambda: mg.apply(x, y),
    "factor": lambda: mg.ctx.factor(L),
}
out = {}
for k, fn in ops.items():
    r = reps if k not in ("factor", "cycle") else max(2, reps // 10)
    fn()
    mg.ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(r):
        fn()
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / r
    gbs = bytes_[k] / ms / 1e6 if bytes_[k] else None
    out[k] = {"ms": ms, "GB/s": gbs}
    print("%-9s %10.4f ms  %s" % (k, ms, "" if gbs is None else "%8.1f GB/s (%.0f%% of 6457)" % (gbs, 100 * gbs / 6457.4)), flush=True)
print(json.dumps({"config": name, "condensed": condense, "apply_bytes": bytes_["apply"],
                  "factor_bytes": mg.ctx.patch_storage_bytes(L), "ops": out}))
