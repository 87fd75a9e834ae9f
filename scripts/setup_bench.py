"""Per-Newton-step setup (PCSetUp: values hand-over, patch inverses, coarse inverse), phase by phase.

    [ALFIB_SCHUR_SETUP=1] python scripts/setup_bench.py [config] [reps]
Wall time around each library call with a stream synchronise on both sides (these are host-visible phases)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d-sv-k3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob = build_problem(name)
levels = [level_input_from_synth(l) for l in prob.levels]
mg = DeviceMultigrid(levels, prob.config.m, condense=True, torch_storage=True)
c = mg.ctx
c.synchronize()


def timed(fn):
    c.synchronize()
    t0 = time.perf_counter()
    fn()
    c.synchronize()
    return time.perf_counter() - t0


rows = []
for rep in range(reps):
    row = {}
    for l, li in enumerate(levels):
        row["values_l%d" % l] = timed(lambda: c.set_bsr_values(l, li.vals))
        if l > 0:
            row["factor_l%d" % l] = timed(lambda: c.factor(l))
    row["coarse_factor"] = timed(c.coarse_factor)
    row["transfer_update"] = sum(timed(lambda: c.transfer_update(l, li.a0_vals, li.d_vals))
                                 for l, li in enumerate(levels) if l > 0 and li.a0_vals is not None)
    row["total_per_newton_step"] = sum(v for k, v in row.items() if k != "transfer_update")
    rows.append(row)
    print(json.dumps({k: round(v, 4) for k, v in row.items()}), flush=True)
print(json.dumps({"config": name, "schur": os.environ.get("ALFIB_SCHUR_SETUP", "0"),
                  "value_bytes": [int(li.vals.nbytes) for li in levels], "best": {k: min(r[k] for r in rows) for k in rows[0]}}))
