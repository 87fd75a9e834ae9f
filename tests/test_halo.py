"""Owned/ghost layouts for distributed level vectors (alfi_b200/halo.py) — the N > 1 host logic of the next
multi-GPU step (owner->ghost broadcast, ghost->owner sum, neighbour exchanges only).

CPU: layout invariants, the two exchange steps against their global definitions, a level smoother executed
rank by rank on local arrays only (oracle/distributed.py) against the serial oracle, and the same exchange
steps with two real processes over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alfi_b200.dist import partition_patches
from alfi_b200.halo import build_layout
from oracle import distributed as od
from oracle import hotpath as hp


def layout_for(prob, level, nranks):
    ld = prob.levels[level]
    ps = ld.patches
    owner = partition_patches(ps.offsets, ps.dofs, nranks)
    return build_layout(ps.offsets, ps.dofs, ps.order, owner, ld.A.rowptr, ld.A.colidx, ld.V.bs, ld.V.ndofs), owner


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc3d-sv-k3-tiny", "ldc2d-pkp0-tiny", "bfs2d-sv-k2-tiny"])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_layout_invariants(problems, name, nranks):
    prob = problems(name, gamma=10.0, nu=0.2)
    level = len(prob.levels) - 1
    ld = prob.levels[level]
    ps = ld.patches
    lay, owner = layout_for(prob, level, nranks)
    n = ld.V.ndofs
    assert lay.owner.min() >= 0 and lay.owner.max() < lay.nranks
    assert np.array_equal(np.sort(np.concatenate([r.owned for r in lay.ranks])), np.arange(n))     # a partition
    A = ld.A.to_csr()
    for r in lay.ranks:
        assert np.array_equal(lay.owner[r.owned], np.full(r.n_owned, r.rank)) and (lay.owner[r.ghost] != r.rank).all()
        loc = set(r.local.tolist())
        assert np.array_equal(r.patches, ps.order[owner[ps.order] == r.rank])                        # iteration order kept
        for p in r.patches:
            assert set(ps.patch(p).tolist()) <= loc                                                  # patch gather is local
        cols = np.unique(A[r.owned].indices)
        assert set(cols.tolist()) <= loc                                                             # SpMV columns are local
        for peer, pos in r.recv.items():
            assert np.array_equal(r.ghost[pos], lay.ranks[peer].owned[lay.ranks[peer].send[r.rank]])
        assert sum(v.size for v in r.recv.values()) == r.ghost.size
    mx, tot = lay.exchange_bytes()
    assert tot == 8 * sum(r.ghost.size for r in lay.ranks) and (nranks > 1) == (tot > 0)


@pytest.mark.parametrize("nranks", [2, 4])
def test_exchange_steps_are_the_sf_pattern(problems, nranks):
    """update_ghosts = every copy equals the owner's value; reduce_ghosts = the owner gets the sum of all copies
    (PetscSF bcast / reduce(sum) of the reference's PCPATCH)."""
    prob = problems("ldc3d-sv-k3-tiny", gamma=10.0, nu=0.2)
    lay, _ = layout_for(prob, 1, nranks)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(lay.ndofs)
    locs = [np.concatenate([x[r.owned], np.full(r.ghost.size, np.nan)]) for r in lay.ranks]
    lay.update_ghosts(locs)
    for r in lay.ranks:
        assert np.array_equal(locs[r.rank], x[r.local])
    parts = [rng.standard_normal(r.n_local) for r in lay.ranks]
    total = np.zeros(lay.ndofs)
    for r, p in zip(lay.ranks, parts):
        np.add.at(total, r.local, p)
    locs = [p.copy() for p in parts]
    lay.reduce_ghosts(locs)
    assert np.allclose(lay.gather(locs), total, rtol=0, atol=1e-14)
    assert all((l[r.n_owned:] == 0).all() for r, l in zip(lay.ranks, locs))


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc3d-sv-k3-tiny", "bfs2d-sv-k2-tiny"])
@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_distributed_smoother_equals_serial(problems, name, nranks):
    """FGMRES(m) + patch smoother + SpMV on local arrays only == hp.smooth; 3 exchanges per Krylov iteration."""
    prob = problems(name, gamma=10.0, nu=0.2)
    ld = prob.levels[1]
    lv = hp.level_from_host(ld)
    lay, _ = layout_for(prob, 1, nranks)
    rng = np.random.default_rng(1)
    b, x0 = rng.standard_normal(lv.n), rng.standard_normal(lv.n)
    b[lv.bc_dofs] = 0
    x0[lv.bc_dofs] = 0
    m = prob.config.m
    want = hp.smooth(lv, b, x0, m)
    got, stats = od.smooth(lv, lay, b, x0, m)
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
    assert stats["reduce"] == m and stats["update"] == 2 * m + 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alfi_b200.synth.problem import build_problem
    prob = build_problem("ldc2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    ld = prob.levels[1]
    lv = hp.level_from_host(ld)
    ps = ld.patches
    owner = partition_patches(ps.offsets, ps.dofs, world)
    lay = build_layout(ps.offsets, ps.dofs, ps.order, owner, ld.A.rowptr, ld.A.colidx, ld.V.bs, lv.n)
    me = lay.ranks[rank]
    x = np.random.default_rng(3).standard_normal(lv.n)
    x[lv.bc_dofs] = 0

    def update(loc):                                   # owner -> ghost with point-to-point messages
        reqs, bufs = [], {}
        for peer in sorted(me.send):
            reqs.append(dist.isend(torch.from_numpy(loc[me.send[peer]].copy()), dst=peer))
        for peer in sorted(me.recv):
            bufs[peer] = torch.empty(me.recv[peer].size, dtype=torch.float64)
            reqs.append(dist.irecv(bufs[peer], src=peer))
        for q in reqs:
            q.wait()
        for peer, t in bufs.items():
            loc[me.n_owned + me.recv[peer]] = t.numpy()

    def reduce(loc):                                   # ghost -> owner sum, peers in ascending order
        reqs, bufs = [], {}
        for peer in sorted(me.recv):
            reqs.append(dist.isend(torch.from_numpy(loc[me.n_owned + me.recv[peer]].copy()), dst=peer))
        for peer in sorted(me.send):
            bufs[peer] = torch.empty(me.send[peer].size, dtype=torch.float64)
            reqs.append(dist.irecv(bufs[peer], src=peer))
        for q in reqs:
            q.wait()
        for peer in sorted(bufs):
            loc[me.send[peer]] += bufs[peer].numpy()
        loc[me.n_owned:] = 0.0

    loc = np.concatenate([x[me.owned], np.zeros(me.ghost.size)])
    update(loc)
    ok_update = bool(np.array_equal(loc, x[me.local]))
    g2l = np.full(lv.n, -1, dtype=np.int64)
    g2l[me.local] = np.arange(me.n_local)
    y = np.zeros(me.n_local)
    for p in me.patches:                               # my patches only, gathered from my local vector
        I = g2l[ps.patch(p)]
        if I.size:
            y[I] += hp._solve(lv.factors[p], loc[I])
    reduce(y)
    want = hp.smoother_apply(x, lv.offsets, lv.dofs, lv.order, lv.factors, np.empty(0, np.int64))
    err = float(np.linalg.norm(y[:me.n_owned] - want[me.owned]) / np.linalg.norm(want))
    out[rank] = (ok_update, err, int(me.ghost.size), int(lv.n))
    dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    for r in range(world):
        ok, err, nghost, n = res[r]
        assert ok and err <= 1e-13 and 0 < nghost < n, res


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny", "bfs2d-sv-k2-tiny"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_distributed_fcycle_equals_serial(problems, name, nranks):
    """The whole fieldsplit_0 application — F-cycle, FGMRES smoothers, robust prolongation / restriction with
    their cell-patch solves, coarse solve — executed rank by rank on owned/ghost arrays (level 0 replicated,
    intermediate levels with a transfer halo for P_H) equals oracle.hotpath.fcycle."""
    prob = problems(name, gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in prob.levels]
    b = np.random.default_rng(2).standard_normal(lv[-1].n)
    b[lv[-1].bc_dofs] = 0
    want = hp.fcycle(lv, b, prob.config.m)
    got, h = od.fcycle(prob, lv, b, prob.config.m, nranks)
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
    assert h.exchanges > 0
    for l in range(1, len(lv)):
        lay = h.layouts[l]
        cp = prob.levels[l].cell_patches
        assert lay.extra_owner.shape == (cp.npatch,)
        for q in range(cp.npatch):                                   # every cell patch is local to its owner
            assert set(cp.patch(q).tolist()) <= set(lay.ranks[lay.extra_owner[q]].local.tolist())
        if l > 1:
            halo = h.halos[l]
            assert all(np.array_equal(a.owned, b_.owned) for a, b_ in zip(halo.ranks, h.layouts[l - 1].ranks))


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc3d-sv-k3-tiny", "ldc2d-pkp0-tiny"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_rank_local_inputs_are_self_sufficient(problems, name, nranks):
    """alfi_b200.halo.local_level: every rank's LocalLevel (local numbering, operator rows of its local nodes,
    its patches / cell patches / Dirichlet lists / transfer columns) is enough to reproduce the global smoother
    application, SpMV and the transfer's operators — nothing global is consulted."""
    import scipy.sparse as sp
    from alfi_b200.halo import local_level, transfer_halo
    from alfi_b200.multigrid import level_input_from_synth
    prob = problems(name, gamma=10.0, nu=0.2)
    layouts, halos = od.build_hierarchy_layouts(prob, nranks)
    rng = np.random.default_rng(4)
    for l in range(1, len(prob.levels)):
        ld = prob.levels[l]
        li = level_input_from_synth(ld)
        lv = hp.level_from_host(ld)
        lay, halo = layouts[l], halos[l]
        locs = [local_level(li, lay, r, halo) for r in range(nranks)]
        data = [od.LocalRankData(ll) for ll in locs]
        x = rng.standard_normal(lv.n)
        x[lv.bc_dofs] = 0
        xs = lay.scatter(x)
        for ll, r in zip(locs, lay.ranks):
            assert np.array_equal(ll.local_dofs, r.local) and ll.n_owned == r.n_owned
            assert ll.rowptr.size == ll.n_local_nodes + 1 and ll.colidx.max() < ll.n_local_nodes
        got = lay.gather(od.local_spmv(lay, locs, data, xs))
        assert np.linalg.norm(got - lv.A @ x) <= 1e-13 * np.linalg.norm(lv.A @ x)
        want = hp.smoother_apply(x, lv.offsets, lv.dofs, lv.order, lv.factors, lv.bc_dofs)
        got = lay.gather(od.local_smoother_apply(lay, locs, data, xs))
        assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
        # transfer pieces: P rows of the owned dofs in the coarse numbering of this rank; cell patches local and owned once
        P = ld.P.tocsr() if ld.P_dof_level else sp.kron(ld.P, sp.identity(ld.V.bs), format="csr")
        c = rng.standard_normal(P.shape[1])
        for ll, r in zip(locs, lay.ranks):
            cl = c if ll.coarse_local is None else c[ll.coarse_local]
            assert np.allclose(ll.P @ cl, (P @ c)[r.owned], rtol=0, atol=1e-13)
            for q, (I, X) in zip(ll.cell_ids, data[ll.rank].cells):
                assert np.array_equal(ll.local_dofs[I], ld.cell_patches.patch(q))
                A0 = ld.A0.to_csr()
                G = ld.cell_patches.patch(q)
                assert np.allclose(X @ A0[G][:, G].toarray(), np.eye(G.size), atol=1e-9)
            assert np.array_equal(np.sort(ll.local_dofs[ll.cb_dofs]), np.intersect1d(ld.cb_dofs, ll.local_dofs))
        assert sorted(np.concatenate([ll.cell_ids for ll in locs]).tolist()) == list(range(ld.cell_patches.npatch))
        assert sorted(np.concatenate([ll.patch_ids for ll in locs]).tolist()) == sorted(set(ld.patches.order.tolist()))
