"""Multi-GPU parity check of the DISTRIBUTED-VECTOR path (alfib_level_set_halo):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/dist_check_halo.py [config] [--time]

Every rank holds only its LocalLevels; the smoother application, SpMV, FGMRES smoother, prolongation,
restriction and the whole cycle are compared, step by step, with the CPU oracle of the unsharded problem
(rank 0 prints one line per step, so the first failing step is visible).  ALFIB_PEER=1: exchanges over NVLink
peer memory instead of NCCL send/recv.  --time also prints ms per cycle
(CUDA events, max over ranks) next to the replicated-vector design on the same problem.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from alfi_b200.dist import bootstrap_unique_id  # noqa: E402
from alfi_b200.multigrid import DeviceMultigrid, DistributedMultigrid, level_input_from_synth  # noqa: E402
from alfi_b200.synth.problem import build_problem  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args else "ldc3d-sv-k3-tiny"
timed = "--time" in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
small = name.endswith("tiny") or name.endswith("small")
prob = build_problem(name, gamma=10.0, nu=0.2) if small else build_problem(name)
levels = [level_input_from_synth(l) for l in prob.levels]
mg = DistributedMultigrid(levels, prob.config.m, rank, world, bootstrap_unique_id(rank), device=local,
                          deterministic=bool(int(os.environ.get("ALFIB_DET", "0"))), torch_storage=not small,
                          peer_memory=bool(int(os.environ.get("ALFIB_PEER", "0"))))
c = mg.ctx
L = len(levels) - 1
n = prob.finest.ndofs
b = np.random.default_rng(1).standard_normal(n)
b[prob.finest.bc_dofs] = 0


def fine_local(v):
    return mg.scatter(v)


def to_global(v_local, level=L):
    """owned parts of a level's local vector -> global vector on every rank"""
    ll = mg.local[level]
    out = torch.zeros(mg.layouts[level].ndofs, dtype=torch.float64, device="cuda")
    out[torch.from_numpy(ll.local_dofs[:ll.n_owned]).cuda()] = torch.from_numpy(v_local[:ll.n_owned]).cuda()
    dist.all_reduce(out)
    return out.cpu().numpy()


nl = mg.local[L].n_local
res = {}
res["apply"] = to_global(c.smoother_apply(L, fine_local(b), np.empty(nl)))
res["spmv"] = to_global(c.spmv(L, fine_local(b), np.empty(nl)))
x0 = np.zeros(nl)
res["smooth"] = to_global(c.smooth(L, prob.config.m, fine_local(b), x0))
nc = prob.levels[L - 1].ndofs
cvec = np.random.default_rng(2).standard_normal(nc)
cvec[prob.levels[L - 1].bc_dofs] = 0
if L - 1 == 0:
    cl = cvec.copy()                                   # level 0 is replicated
    ncl = nc
else:
    cl = np.ascontiguousarray(cvec[mg.local[L - 1].local_dofs])
    ncl = cl.size
res["prolong"] = to_global(c.prolong(L, cl, np.empty(nl)))
rc = c.restrict(L, fine_local(b), np.empty(ncl))
res["restrict"] = rc.copy() if L - 1 == 0 else to_global(rc, L - 1)
res["cycle"] = to_global(mg.apply(fine_local(b), np.empty(nl)))
res["cycle_graph"] = to_global(mg.apply(fine_local(b), np.empty(nl)))
for _ in range(2):
    res["cycle_graph"] = to_global(mg.apply(fine_local(b), np.empty(nl)))      # third call on: CUDA-graph replay

if rank == 0:
    if n <= 400000:
        from oracle import hotpath as hp
        olv = [hp.level_from_host(l) for l in prob.levels]
        lv = olv[L]
        want = {"apply": hp.smoother_apply(b, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs), "spmv": lv.A @ b,
                "smooth": hp.smooth(lv, b, np.zeros(n), prob.config.m), "prolong": hp.prolong(lv, cvec),
                "restrict": hp.restrict(lv, b, olv[L - 1].bc_dofs), "cycle": hp.fcycle(olv, b, prob.config.m)}
        want["cycle_graph"] = want["cycle"]
        for k, v in res.items():
            print("world %d %s: %-12s rel diff vs serial oracle %.2e" % (world, name, k, np.linalg.norm(v - want[k]) / np.linalg.norm(want[k])),
                  flush=True)
    else:
        print("world %d %s: too large for the CPU oracle here; graph replay vs eager %.2e" % (
            world, name, np.linalg.norm(res["cycle_graph"] - res["cycle"]) / np.linalg.norm(res["cycle"])), flush=True)

if timed:
    def time_cycles(apply, bvec, xvec, stream_handle, reps=5):
        stream = torch.cuda.ExternalStream(stream_handle, device=torch.device("cuda", local))
        for _ in range(3):
            apply(bvec, xvec)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            apply(bvec, xvec)
        e1.record(stream)
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])
    bd = torch.from_numpy(fine_local(b)).cuda()
    xd = torch.empty_like(bd)
    t_halo = time_cycles(mg.apply, bd, xd, c.stream)
    c.profile(True)
    c.profile_reset()
    mg.apply(bd, xd)
    prof = c.profile_get(-1)
    c.profile(False)
    lay = mg.layouts[L]
    mg.ctx.close()
    del mg
    torch.cuda.empty_cache()
    rep = DeviceMultigrid(levels, prob.config.m, device=local, rank=rank, nranks=world, unique_id=bootstrap_unique_id(rank),
                          torch_storage=not small)
    bg = torch.from_numpy(b).cuda()
    xg = torch.empty_like(bg)
    t_rep = time_cycles(rep.apply, bg, xg, rep.ctx.stream)
    if rank == 0:
        print("world %d %s: ms per cycle  distributed vectors %.2f | replicated vectors %.2f | ghosts/owned on rank 0: %d/%d, "
              "exchange bytes per rank (max) %d" % (world, name, t_halo, t_rep, lay.ranks[0].ghost.size, lay.ranks[0].n_owned,
                                                     lay.exchange_bytes()[0]), flush=True)
        print("  events of one eager cycle (rank 0):", {k: (round(v[0], 3), v[1]) for k, v in prof.items()}, flush=True)
dist.barrier()
dist.destroy_process_group()
