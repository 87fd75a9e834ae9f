"""Multi-GPU parity check of RANK-LOCALLY generated problems (alfi_b200/synth/bricks.py -> DistributedMultigrid.from_local):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        scripts/dist_check_bricks.py [config, default ldc3d-sv-k3-wtiny2]        (ALFIB_PEER=1: NVLink peer-memory transport)

The number of ranks must equal the config's rank grid.  Every rank builds only its brick; rank 0 additionally builds
the global box problem and its serial CPU oracle, and the distributed cycle is compared with it through the nodes'
lattice keys."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from alfi_b200.dist import bootstrap_unique_id  # noqa: E402
from alfi_b200.multigrid import DistributedMultigrid  # noqa: E402
from alfi_b200.synth.bricks import build_rank_local, node_keys  # noqa: E402
from alfi_b200.synth.problem import CONFIGS, build_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ldc3d-sv-k3-wtiny2"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = CONFIGS[name]
assert int(np.prod(cfg.shape)) == world, "run with as many ranks as the config's rank grid"
rlp = build_rank_local(cfg, rank, nu=0.2, gamma=10.0)
mg = DistributedMultigrid.from_local(rlp, cfg.m, bootstrap_unique_id(rank), device=local,
                                     peer_memory=bool(int(os.environ.get("ALFIB_PEER", "0"))))
L = len(rlp.local) - 1
ll = rlp.local[L]
bs = ll.bs
# a right-hand side that is a function of the node key, so that every rank fills its entries without the global vector
key = rlp.keys[L]
bl = (np.sin(0.37 * (key % 1000003))[:, None] + 0.1 * np.arange(bs)[None, :]).ravel()
bl[ll.bc_dofs] = 0.0
xl = mg.apply(bl, np.empty(ll.n_local))
for _ in range(2):
    xg = mg.apply(bl, np.empty(ll.n_local))                       # CUDA-graph replay from the third call on
pieces = [None] * world
dist.all_gather_object(pieces, (key[:ll.n_owned // bs], xl[:ll.n_owned], xg[:ll.n_owned]))
if rank == 0:
    from oracle import hotpath as hp
    glob = build_problem(cfg, gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in glob.levels]
    _, gkey = node_keys(glob.finest.V.node_coords, cfg.N * 2 ** L, cfg.length, cfg.shape)
    b = (np.sin(0.37 * (gkey % 1000003))[:, None] + 0.1 * np.arange(bs)[None, :]).ravel()
    b[lv[-1].bc_dofs] = 0.0
    want = hp.fcycle(lv, b, cfg.m)
    order = np.argsort(gkey)
    got, gotg = np.full(b.size, np.nan), np.full(b.size, np.nan)
    for k, v, vg in pieces:
        pos = order[np.searchsorted(gkey[order], k)]
        idx = (pos[:, None] * bs + np.arange(bs)[None, :]).ravel()
        got[idx], gotg[idx] = v, vg
    print("world %d %s (rank-local generation): cycle rel diff vs serial oracle %.2e, graph replay %.2e, owned sets cover %s" % (
        world, name, np.linalg.norm(got - want) / np.linalg.norm(want), np.linalg.norm(gotg - want) / np.linalg.norm(want),
        bool(np.isfinite(got).all())), flush=True)
dist.barrier()
dist.destroy_process_group()
