"""Condensed patch apply: shared vs per-instance blocks x tile op v2 vs v1, same problem, one process.

    python scripts/variant_bench.py [config] [reps]
Times the finest-level PCApply_PATCH, one FGMRES(m) smoother call and the whole F-cycle with CUDA
events on the library stream, and prints the stored / algorithmic bytes of each variant."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d-sv-k3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
prob = build_problem(name)
L = len(prob.levels) - 1
n = prob.finest.ndofs
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
out = []
for shared in ("1", "0"):
    os.environ["ALFIB_CONDENSE_SHARED"] = shared
    t0 = time.time()
    mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, condense=True)
    mg.ctx.synchronize()
    setup = time.time() - t0
    stream = torch.cuda.ExternalStream(mg.ctx.stream)
    nbytes = mg.ctx.patch_apply_bytes(L)
    for v1 in ("0", "1"):
        os.environ["ALFIB_TILE_V1"] = v1
        row = {"shared": int(shared), "tile_v1": int(v1), "setup_s": setup, "apply_bytes": nbytes,
               "storage_bytes": mg.ctx.patch_storage_bytes(L)}
        ops = {"apply": (lambda: mg.ctx.smoother_apply(L, x, y), reps),
               "smooth": (lambda: mg.ctx.smooth(L, prob.config.m, x, y), max(2, reps // 10))}
        for k, (fn, r) in ops.items():
            fn()
            mg.ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(r):
                fn()
            e1.record(stream)
            e1.synchronize()
            row[k + "_ms"] = e0.elapsed_time(e1) / r
        row["apply_GBs"] = nbytes / row["apply_ms"] / 1e6
        out.append(row)
        print(json.dumps(row), flush=True)
    mg.ctx.close()
    del mg
    torch.cuda.empty_cache()
print(json.dumps({"config": name, "variants": out}))
