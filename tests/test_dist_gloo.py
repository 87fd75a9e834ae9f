"""N > 1 host logic on CPU: world_size-2 gloo run of the patch sharding.

Each rank applies only the patches `partition_patches` gives it (numpy oracle as the local
solver), the contributions are summed with a gloo all_reduce — the same exchange step the
library performs with ncclAllReduce — and the result must equal the unsharded application."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alfi_b200.dist import partition_patches, shard_patch_arrays


def test_partition_is_balanced_and_complete(problems):
    prob = problems("ldc3d-sv-k3-tiny", gamma=10.0, nu=0.2)
    ps = prob.levels[1].patches
    for nranks in (1, 2, 4):
        owner = partition_patches(ps.offsets, ps.dofs, nranks)
        assert owner.min() >= 0 and owner.max() < nranks
        cost = np.bincount(owner, weights=ps.sizes.astype(float) ** 2, minlength=nranks)
        assert cost.max() <= 2.0 * cost.sum() / nranks + ps.sizes.max() ** 2
        seen = []
        for r in range(nranks):
            off, dofs, order, cols, mine = shard_patch_arrays(ps.offsets, ps.dofs, ps.order, ps.colours, owner, r)
            seen.append(mine)
            for k, p in enumerate(mine):
                assert np.array_equal(dofs[off[k]:off[k + 1]], ps.patch(p))
            assert sorted(order.tolist()) == list(range(mine.size))
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(ps.npatch))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alfi_b200.synth.problem import build_problem
    from oracle import hotpath as hp
    prob = build_problem("ldc2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    lv = hp.level_from_host(prob.levels[1])
    ps = prob.levels[1].patches
    x = np.random.default_rng(0).standard_normal(lv.n)
    owner = partition_patches(ps.offsets, ps.dofs, world)
    off, dofs, order, cols, mine = shard_patch_arrays(ps.offsets, ps.dofs, ps.order, ps.colours, owner, rank)
    factors = [lv.factors[p] for p in mine]
    y = hp.smoother_apply(x, off, dofs, order, factors, np.empty(0, np.int64))      # local patches only
    t = torch.from_numpy(y)
    dist.all_reduce(t)                                                               # the exchange step
    y = t.numpy()
    y[lv.bc_dofs] = x[lv.bc_dofs]
    want = hp.smoother_apply(x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)
    out[rank] = float(np.linalg.norm(y - want) / np.linalg.norm(want))
    # row-sharded SpMV + broadcast of owned rows
    n_nodes = prob.levels[1].V.nnodes
    bs = prob.levels[1].V.bs
    start = [n_nodes * r // world * bs for r in range(world + 1)]
    z = np.zeros(lv.n)
    z[start[rank]:start[rank + 1]] = (lv.A @ x)[start[rank]:start[rank + 1]]
    tz = torch.from_numpy(z)
    for r in range(world):
        dist.broadcast(tz[start[r]:start[r + 1]], src=r)
    out[world + rank] = float(np.linalg.norm(tz.numpy() - lv.A @ x))
    dist.destroy_process_group()


def test_sharded_apply_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    for r in range(world):
        assert res[r] <= 1e-13, res
        assert res[world + r] == 0.0, res


def test_condensed_cost_balances_condensed_bytes(problems):
    """With condensed inverses the partition balances what is streamed per application, not n^2."""
    from alfi_b200.dist import condensed_cost
    prob = problems("ldc3d-sv-k3-tiny", gamma=10.0, nu=0.2)
    ps = prob.levels[1].patches
    cost = condensed_cost(ps.offsets, ps.blocks)
    assert cost.shape == (ps.npatch,) and (cost > 0).all()
    # interior patch: 195 separator dofs and 24 blocks of 45 dofs
    p = int(np.argmax(ps.sizes))
    assert cost[p] == 195.0 ** 2 + 24 * 3 * 45.0 ** 2
    for nranks in (2, 4):
        owner = partition_patches(ps.offsets, ps.dofs, nranks, cost)
        share = np.bincount(owner, weights=cost, minlength=nranks)
        assert share.max() <= 2.0 * cost.sum() / nranks + cost.max()
        assert np.array_equal(np.sort(np.unique(owner)), np.arange(nranks))
