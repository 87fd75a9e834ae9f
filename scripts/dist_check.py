"""Multi-GPU parity check: torchrun --nproc-per-node N scripts/dist_check.py [config]

Every rank runs the sharded cycle; rank 0 compares with the CPU oracle of the unsharded problem."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from alfi_b200.dist import bootstrap_unique_id  # noqa: E402
from alfi_b200.multigrid import DeviceMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d-sv-k3-tiny"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prob = build_problem(name, gamma=10.0, nu=0.2)
mg = DeviceMultigrid([level_input_from_synth(l) for l in prob.levels], prob.config.m, device=local, deterministic=True,
                     rank=rank, nranks=world, unique_id=bootstrap_unique_id(rank),
                     peer_memory=bool(int(os.environ.get("ALFIB_PEER", "0"))))
n = prob.finest.ndofs
b = np.random.default_rng(1).standard_normal(n)
b[prob.finest.bc_dofs] = 0
L = len(prob.levels) - 1
y = mg.ctx.smoother_apply(L, b, np.empty(n))
z = mg.ctx.spmv(L, b, np.empty(n))
p = mg.ctx.prolong(L, np.ones(prob.levels[L - 1].ndofs), np.empty(n))
x = mg.apply(b, np.empty(n))
# all ranks must hold identical (replicated) results
t = torch.from_numpy(np.concatenate([x, y, z, p])).cuda()
t0 = t.clone()
dist.broadcast(t0, src=0)
same = bool(torch.equal(t, t0))
if rank == 0:
    from oracle import hotpath as hp
    olv = [hp.level_from_host(l) for l in prob.levels]
    rel = lambda a, c: np.linalg.norm(a - c) / np.linalg.norm(c)
    lv = olv[L]
    c1 = np.ones(olv[L - 1].n)
    print("world %d %s: apply %.2e spmv %.2e prolong %.2e cycle %.2e" % (
        world, name, rel(y, hp.smoother_apply(b, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)),
        rel(z, lv.A @ b), rel(p, hp.prolong(lv, c1)), rel(x, hp.fcycle(olv, b, prob.config.m))), flush=True)
flags = [None] * world
dist.all_gather_object(flags, same)
if rank == 0:
    print("replicated results identical on all ranks:", all(flags), flush=True)
dist.barrier()
dist.destroy_process_group()
