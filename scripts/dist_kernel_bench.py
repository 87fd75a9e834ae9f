"""Per-operation timing of the DISTRIBUTED-VECTOR path on N GPUs (CUDA events on the library stream, max over ranks):

    [ALFIB_PEER=1] python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P scripts/dist_kernel_bench.py [config] [reps]

SpMV (= owner->ghost update + owned rows), PCApply_PATCH (= update + patches + ghost->owner sum), one FGMRES(m)
smoother call on every level >= 1, and the cycle; the difference to the single-GPU kernel times divided by the
number of exchanges is the cost of one exchange."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from alfi_b200.dist import bootstrap_unique_id  # noqa: E402
from alfi_b200.multigrid import DistributedMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args else "ldc3d-sv-k3"
reps = int(args[1]) if len(args) > 1 else 50
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prob = build_problem(name)
levels = [level_input_from_synth(l) for l in prob.levels]
peer = bool(int(os.environ.get("ALFIB_PEER", "0")))
mg = DistributedMultigrid(levels, prob.config.m, rank, world, bootstrap_unique_id(rank), device=local, torch_storage=True,
                          peer_memory=peer)
c = mg.ctx
stream = torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local))
out = {"config": name, "world": world, "peer_memory": peer, "mbox_off": os.environ.get("ALFIB_MBOX_OFF", "0"), "levels": {}}


def timed(fn, r):
    for _ in range(3):
        fn()
    c.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(r):
        fn()
    e1.record(stream)
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / r], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


for l in range(1, len(levels)):
    ll = mg.local[l]
    x = torch.randn(ll.n_local, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    row = {"n_owned": int(ll.n_owned), "n_ghost": int(ll.n_local - ll.n_owned),
           "spmv_ms": timed(lambda: c.spmv(l, x, y), reps),
           "apply_ms": timed(lambda: c.smoother_apply(l, x, y), reps),
           "smooth_ms": timed(lambda: c.smooth(l, prob.config.m, x, y), max(3, reps // 10))}
    out["levels"][l] = row
    if rank == 0:
        print("level", l, json.dumps(row), flush=True)
ll = mg.local[-1]
b = torch.randn(ll.n_local, dtype=torch.float64, device="cuda")
xx = torch.empty_like(b)
out["cycle_ms"] = timed(lambda: mg.apply(b, xx), max(3, reps // 10))
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
