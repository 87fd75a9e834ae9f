"""Cycle timing of the small BASELINE configs, eager launches vs CUDA graph replay."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

for name in sys.argv[1:] or ["ldc2d-sv-k2", "ldc2d-pkp0"]:
    prob = build_problem(name)
    for graph in (0, 1):
        mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m)
        mg.ctx.set_option(5, graph)
        n = prob.finest.ndofs
        b = torch.randn(n, dtype=torch.float64, device="cuda")
        x = torch.empty_like(b)
        for _ in range(4):
            mg.apply(b, x)
        mg.ctx.synchronize()
        l0 = mg.ctx.launches
        stream = torch.cuda.ExternalStream(mg.ctx.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(20):
            mg.apply(b, x)
        e1.record(stream)
        e1.synchronize()
        wall = (time.perf_counter() - t0) / 20 * 1e3
        print("%-14s graph=%d  %8.3f ms/cycle (device)  %8.3f ms (wall)  %7.2f MDoF/s  launches/cycle %d  levels %d  dofs %d"
              % (name, graph, e0.elapsed_time(e1) / 20, wall, n / (e0.elapsed_time(e1) / 20) / 1e3,
                 (mg.ctx.launches - l0) // 20, len(prob.levels), n), flush=True)
        xr = x.clone()
        mg.ctx.close()
    # graph and eager results must agree
