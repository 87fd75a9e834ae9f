"""Debug aid: condensed apply on the clustered edge case, repeated; reports where runs differ."""
import sys

import numpy as np

sys.path.insert(0, ".")
from alfi_b200.lib import Context  # noqa: E402
from tests.condense_cases import clustered_problem, dense_reference  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
case = clustered_problem(bs, seed=bs)
n = case["n_nodes"] * bs
x = np.random.default_rng(11).standard_normal(n)
npatch = len(case["patches"])
ctx = Context(deterministic=True)
ctx.level_create(0, case["n_nodes"], bs)
ctx.set_bsr_pattern(0, case["rowptr"], case["colidx"])
ctx.set_bsr_values(0, case["vals"])
ctx.set_bc(0, np.empty(0, np.int32))
ctx.set_patches(0, case["offsets"], case["dofs"], np.arange(npatch, dtype=np.int32), None)
ctx.set_patch_blocks(0, case["blocks"])
ctx.factor(0)
cols = ctx.colours(0, npatch)
print("colours", cols.tolist(), "sizes", case["sizes"].tolist())
ys = [ctx.smoother_apply(0, x, np.empty(n)).copy() for _ in range(reps)]
ref = dense_reference(case, range(npatch), x)
print("rel err", np.linalg.norm(ys[0] - ref) / np.linalg.norm(ref))
for k in range(1, reps):
    d = np.flatnonzero(ys[k] != ys[0])
    print("run", k, "differs at", d.size, "dofs", d[:20].tolist(), "max abs diff", np.abs(ys[k] - ys[0]).max())
    for g in d[:6]:
        where = []
        for p, I in enumerate(case["patches"]):
            pos = np.flatnonzero(I == g)
            if pos.size:
                where.append((p, int(case["blocks"][case["offsets"][p] + pos[0]])))
        print("   dof", g, "in (patch, block label)", where)
ctx.close()
