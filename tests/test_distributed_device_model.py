"""Host model of the library's distributed-vector path (csrc: alfib_level_set_halo, halo_update / halo_reduce in
comm.cu, the halo branches of patch_apply_sum, launch_bsr_spmv, fgmres_device, prolong_device, restrict_device).

Test infrastructure.  `alfi_b200.multigrid.DistributedMultigrid` is run once per rank against a recording stand-in
of `alfi_b200.lib.Context` that applies the argument checks of csrc/api.cu; `Lockstep` then executes, for all ranks
side by side and from the recorded hand-over data ALONE, the exact sequence of steps the CUDA code enqueues —
including which step refreshes which ghosts.  Entries the device code leaves undefined (ghost parts of vectors
that were only produced on the owned rows) are NaN here, so a gather that reads a ghost nobody refreshed poisons
the result.  The outcome must equal the serial oracle (oracle/hotpath.py fcycle).  What this does not cover: the
kernels themselves and NCCL — those need the GPU run (scripts/dist_check_halo.py).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from alfi_b200.multigrid import DistributedMultigrid, level_input_from_synth
from oracle import hotpath as hp

NAN = float("nan")


class RecordingContext:
    """Records what DistributedMultigrid hands over; checks like csrc/api.cu."""

    def __init__(self):
        self.levels = {}
        self.options = {}
        self.nlevels = None

    # -- plumbing
    def set_option(self, key, value):
        self.options[key] = value

    def comm_init(self, unique_id, rank, nranks):
        self.rank, self.nranks = rank, nranks

    def level_create(self, level, n_nodes, bs):
        assert level not in self.levels and n_nodes > 0 and bs in (2, 3)
        self.levels[level] = dict(n_nodes=n_nodes, bs=bs, n=n_nodes * bs, n_owned=n_nodes * bs, halo=None, thalo=None,
                                  has_transfer=False, ps={})

    def set_halo(self, level, n_owned, n_local, send, recv, which=0, peer_offsets=None):
        L = self.levels[level]
        assert 0 <= n_owned <= n_local and not L["has_transfer"]
        if which == 0:
            assert n_local == L["n"] and n_owned % L["bs"] == 0
        else:
            assert level >= 1 and n_owned == self.levels[level - 1]["n_owned"]
        peers = sorted(set(send) | set(recv))
        assert all(0 <= q < self.nranks and q != self.rank for q in peers)
        ghosts = []
        for q in peers:
            s, r = np.asarray(send.get(q, [])), np.asarray(recv.get(q, []))
            assert ((0 <= s) & (s < n_owned)).all() and np.unique(s).size == s.size
            assert ((n_owned <= r) & (r < n_local)).all()
            ghosts.append(r)
        allg = np.concatenate(ghosts) if ghosts else np.empty(0, np.int64)
        assert np.unique(allg).size == allg.size
        assert self.nranks == 1 or allg.size == n_local - n_owned
        H = dict(n_owned=n_owned, n_local=n_local, peers=peers,
                 send={q: np.asarray(send.get(q, []), dtype=np.int64) for q in peers},
                 recv={q: np.asarray(recv.get(q, []), dtype=np.int64) for q in peers})
        # the packed layout of csrc/api.cu: lists concatenated by ascending peer, gather form of the ghost->owner sum
        H["send_off"] = np.concatenate(([0], np.cumsum([H["send"][q].size for q in peers]))).astype(np.int64)
        H["recv_off"] = np.concatenate(([0], np.cumsum([H["recv"][q].size for q in peers]))).astype(np.int64)
        H["send_idx"] = np.concatenate([H["send"][q] for q in peers]) if peers else np.empty(0, np.int64)
        H["recv_idx"] = np.concatenate([H["recv"][q] for q in peers]) if peers else np.empty(0, np.int64)
        pos = np.argsort(H["send_idx"], kind="stable")
        dof_sorted = H["send_idx"][pos]
        starts = np.flatnonzero(np.concatenate(([True], dof_sorted[1:] != dof_sorted[:-1]))) if pos.size else np.empty(0, np.int64)
        H["red_ptr"] = np.concatenate((starts, [pos.size])).astype(np.int64)
        H["red_dof"], H["red_src"] = dof_sorted[starts], pos
        H["peer_off"] = None
        if peer_offsets is not None:
            H["peer_off"] = (np.array([peer_offsets[q][0] for q in peers], dtype=np.int64),
                             np.array([peer_offsets[q][1] for q in peers], dtype=np.int64))
        if which == 0:
            L["halo"], L["n_owned"] = H, n_owned
        else:
            L["thalo"] = H

    def set_bsr_pattern(self, level, rowptr, colidx):
        L = self.levels[level]
        rowptr, colidx = np.asarray(rowptr), np.asarray(colidx)
        assert rowptr.size == L["n_nodes"] + 1 and rowptr[0] == 0 and rowptr[-1] == colidx.size
        assert ((0 <= colidx) & (colidx < L["n_nodes"])).all()
        L["rowptr"], L["colidx"] = rowptr.astype(np.int64), colidx.astype(np.int64)

    def set_bc(self, level, bc):
        bc = np.asarray(bc, dtype=np.int64)
        assert ((0 <= bc) & (bc < self.levels[level]["n"])).all()
        self.levels[level]["bc"] = bc

    def set_patches(self, level, offsets, dofs, order=None, colours=None, which=0):
        L = self.levels[level]
        offsets, dofs = np.asarray(offsets, dtype=np.int64), np.asarray(dofs, dtype=np.int64)
        assert ((0 <= dofs) & (dofs < L["n"])).all()
        order = np.arange(offsets.size - 1) if order is None else np.asarray(order)
        L["ps"][which] = (offsets, dofs, order)

    def set_patch_blocks(self, level, blocks, which=0):
        assert np.asarray(blocks).size == self.levels[level]["ps"][which][1].size

    def set_transfer(self, level, P, cb, dof_level=False):
        L, Lc = self.levels[level], self.levels[level - 1]
        P = P.tocsr()
        if L["halo"] is not None:
            assert dof_level
            assert L["thalo"] is not None or Lc["halo"] is None
            assert P.shape == (L["n_owned"], L["thalo"]["n_local"] if L["thalo"] is not None else Lc["n"])
        L["P"], L["cb"] = P, np.asarray(cb, dtype=np.int64)
        L["has_transfer"] = True

    def _bsr(self, level, vals):
        L = self.levels[level]
        vals = np.asarray(vals)
        assert vals.shape == (L["colidx"].size, L["bs"], L["bs"])
        return sp.bsr_matrix((vals, L["colidx"], L["rowptr"]), shape=(L["n"], L["n"])).tocsr()

    def set_bsr_values(self, level, vals):
        self.levels[level]["A"] = self._bsr(level, vals)

    def factor(self, level):
        pass

    def coarse_factor(self):
        pass

    def transfer_update(self, level, a0, d):
        self.levels[level]["A0"], self.levels[level]["D"] = self._bsr(level, a0), self._bsr(level, d)

    def cycle_setup(self, nlevels, smoothing):
        self.nlevels, self.smoothing = nlevels, smoothing


class Lockstep:
    """All ranks' library instances advanced together; vectors are lists of per-rank local arrays."""

    def __init__(self, ctxs):
        self.c = ctxs
        self.R = len(ctxs)
        self.nl = ctxs[0].nlevels
        self.m = ctxs[0].smoothing
        self.stats = {"update": 0, "reduce": 0, "allreduce": 0}
        self.inv = {}

    def L(self, r, l):
        return self.c[r].levels[l]

    # ---- comm.cu
    @staticmethod
    def _segment(off, pos):
        p = 0
        while p + 1 < off.size - 1 and pos >= off[p + 1]:
            p += 1
        return p

    def _peer_update(self, Hs, xs):
        """peer_halo_update_kernel: every rank packs x[send_idx] into its slot; ghosts are pulled from the owners' slots"""
        slot = [x[H["send_idx"]].copy() for H, x in zip(Hs, xs)]
        for H, x in zip(Hs, xs):
            for i in range(H["recv_idx"].size):
                p = self._segment(H["recv_off"], i)
                x[H["recv_idx"][i]] = slot[H["peers"][p]][H["peer_off"][0][p] + (i - H["recv_off"][p])]

    def _peer_reduce(self, Hs, ys):
        """peer_halo_sum_kernel: every rank packs y[recv_idx]; owners gather-sum in ascending position = peer order"""
        slot = [y[H["recv_idx"]].copy() for H, y in zip(Hs, ys)]
        for H, y in zip(Hs, ys):
            for i, d in enumerate(H["red_dof"]):
                v = y[d]
                for k in range(H["red_ptr"][i], H["red_ptr"][i + 1]):
                    pos = H["red_src"][k]
                    p = self._segment(H["send_off"], pos)
                    v += slot[H["peers"][p]][H["peer_off"][1][p] + (pos - H["send_off"][p])]
                y[d] = v
            y[H["n_owned"]:] = 0.0

    def halo_update(self, l, key, xs):
        Hs = [self.L(r, l)[key] for r in range(self.R)]
        if Hs[0] is None:
            return
        self.stats["update"] += 1
        if Hs[0]["peer_off"] is not None:
            return self._peer_update(Hs, xs)
        for r, H in enumerate(Hs):
            for q in H["peers"]:
                xs[r][H["recv"][q]] = xs[q][Hs[q]["send"][r]]

    def halo_reduce(self, l, key, ys):
        Hs = [self.L(r, l)[key] for r in range(self.R)]
        if Hs[0] is None:
            return
        self.stats["reduce"] += 1
        if Hs[0]["peer_off"] is not None:
            return self._peer_reduce(Hs, ys)
        packed = {(r, q): ys[r][H["recv"][q]].copy() for r, H in enumerate(Hs) for q in H["peers"]}
        for r, H in enumerate(Hs):
            for q in H["peers"]:                                  # ascending peer order
                ys[r][H["send"][q]] += packed[(q, r)]
            ys[r][H["n_owned"]:] = 0.0

    def allreduce(self, parts):
        self.stats["allreduce"] += 1
        tot = sum(parts)
        return [np.array(tot, copy=True) for _ in parts]

    # ---- spmv.cu: ghosts of x refreshed, owned rows computed, the rest of y undefined
    def spmv(self, l, name, xs, bs_=None):
        if self.L(0, l)["halo"] is None:                          # replicated level: every rank the same full product
            return [(b - self.L(r, l)[name] @ x) if bs_ is not None else self.L(r, l)[name] @ x
                    for r, (x, b) in enumerate(zip(xs, bs_ or [None] * self.R))]
        self.halo_update(l, "halo", xs)
        out = []
        for r, x in enumerate(xs):
            L = self.L(r, l)
            no = L["n_owned"]
            y = np.full(L["n"], NAN)
            y[:no] = (L[name][:no] @ x)
            if bs_ is not None:
                y[:no] = bs_[r][:no] - y[:no]
            out.append(y)
        return out

    # ---- patch_apply.cu: patch_apply_sum with a halo
    def _inverses(self, r, l, which):
        key = (r, l, which)
        if key not in self.inv:
            L = self.L(r, l)
            A = L["A"] if which == 0 else L["A0"]
            off, dofs, order = L["ps"][which]
            self.inv[key] = [np.linalg.inv(A[dofs[off[p]:off[p + 1]]][:, dofs[off[p]:off[p + 1]]].toarray())
                             if off[p + 1] > off[p] else None for p in range(off.size - 1)]
        return self.inv[key]

    def patch_apply_sum(self, l, which, xs):
        self.halo_update(l, "halo", xs)
        ys = []
        for r, x in enumerate(xs):
            L = self.L(r, l)
            off, dofs, order = L["ps"][which]
            inv = self._inverses(r, l, which)
            y = np.zeros(L["n"])
            for p in order:
                I = dofs[off[p]:off[p + 1]]
                if I.size:
                    y[I] += inv[p] @ x[I]
            ys.append(y)
        self.halo_reduce(l, "halo", ys)
        return ys

    def smoother_apply(self, l, xs):
        ys = self.patch_apply_sum(l, 0, xs)
        for r, (x, y) in enumerate(zip(xs, ys)):
            bc = self.L(r, l)["bc"]
            y[bc] = x[bc]
        return ys

    # ---- krylov.cu
    def fgmres(self, l, bs_, xs, m):
        no = [self.L(r, l)["n_owned"] for r in range(self.R)]
        n = [self.L(r, l)["n"] for r in range(self.R)]
        w = self.spmv(l, "A", xs, bs_)
        beta = np.sqrt(self.allreduce([np.dot(w[r][:no[r]], w[r][:no[r]]) for r in range(self.R)])[0])
        V = [np.full((m + 1, n[r]), NAN) for r in range(self.R)]
        Z = [np.full((m, n[r]), NAN) for r in range(self.R)]
        for r in range(self.R):
            V[r][0, :no[r]] = w[r][:no[r]] / beta
        H = np.zeros((m + 1, m))
        for k in range(m):
            vk = [V[r][k] for r in range(self.R)]
            zk = self.smoother_apply(l, vk)
            for r in range(self.R):
                Z[r][k] = zk[r]
            zk = [Z[r][k] for r in range(self.R)]                 # the SpMV refreshes the ghosts of Z_k in place
            w = self.spmv(l, "A", zk)
            h = self.allreduce([V[r][:k + 1, :no[r]] @ w[r][:no[r]] for r in range(self.R)])[0]
            H[:k + 1, k] = h
            for r in range(self.R):
                w[r][:no[r]] -= V[r][:k + 1, :no[r]].T @ h
            nrm = np.sqrt(self.allreduce([np.dot(w[r][:no[r]], w[r][:no[r]]) for r in range(self.R)])[0])
            H[k + 1, k] = nrm
            for r in range(self.R):
                V[r][k + 1, :no[r]] = w[r][:no[r]] / nrm if nrm > 0 else 0.0
        e1 = np.zeros(m + 1)
        e1[0] = beta
        y = np.linalg.lstsq(H, e1, rcond=None)[0]
        for r in range(self.R):
            xs[r][:no[r]] += Z[r][:, :no[r]].T @ y
        return xs

    # ---- cycle.cu
    def cell_block_solve_refined(self, l, bs_):
        ys = self.patch_apply_sum(l, 1, bs_)
        for r in range(self.R):
            cb = self.L(r, l)["cb"]
            ys[r][cb] = bs_[r][cb]
        rr = self.spmv(l, "A0", ys, bs_)
        dy = self.patch_apply_sum(l, 1, rr)
        return [y + d for y, d in zip(ys, dy)]

    def prolong(self, l, coarse):
        src = coarse
        if self.L(0, l)["thalo"] is not None:
            src = []
            for r in range(self.R):
                T = self.L(r, l)["thalo"]
                tc = np.full(T["n_local"], NAN)
                tc[:T["n_owned"]] = coarse[r][:T["n_owned"]]
                src.append(tc)
            self.halo_update(l, "thalo", src)
        rhs = []
        for r in range(self.R):
            L = self.L(r, l)
            v = np.full(L["n"], NAN)
            v[:L["P"].shape[0]] = L["P"] @ src[r]
            rhs.append(v)
        t1 = self.spmv(l, "D", rhs)
        for r in range(self.R):
            t1[r][self.L(r, l)["cb"]] = 0.0
        t2 = self.cell_block_solve_refined(l, t1)
        fine = [a - b for a, b in zip(rhs, t2)]
        for r in range(self.R):
            fine[r][self.L(r, l)["bc"]] = 0.0
        return fine

    def restrict(self, l, fine):
        t1 = [f.copy() for f in fine]
        for r in range(self.R):
            t1[r][self.L(r, l)["cb"]] = 0.0
        t2 = self.cell_block_solve_refined(l, t1)
        t1 = self.spmv(l, "D", t2)
        t2 = [f - b for f, b in zip(fine, t1)]
        parts = [self.L(r, l)["P"].T @ t2[r][:self.L(r, l)["P"].shape[0]] for r in range(self.R)]
        if self.L(0, l)["thalo"] is not None:
            self.halo_reduce(l, "thalo", parts)
            coarse = []
            for r in range(self.R):
                Lc, T = self.L(r, l - 1), self.L(r, l)["thalo"]
                v = np.full(Lc["n"], NAN)
                v[:T["n_owned"]] = parts[r][:T["n_owned"]]
                coarse.append(v)
        elif self.L(0, l)["halo"] is not None and self.R > 1:
            coarse = self.allreduce(parts)
        else:
            coarse = parts
        for r in range(self.R):
            coarse[r][self.L(r, l - 1)["bc"]] = 0.0
        return coarse

    def coarse_solve(self, bs_):
        return [np.linalg.solve(self.L(r, 0)["A"].toarray(), b) for r, b in enumerate(bs_)]

    def vcycle(self, l, b, x):
        if l == 0:
            return self.coarse_solve(b)
        x = self.fgmres(l, b, x, self.m)
        res = self.spmv(l, "A", x, b)
        bc = self.restrict(l, res)
        xc = self.vcycle(l - 1, bc, [np.zeros_like(v) for v in bc])
        p = self.prolong(l, xc)
        x = [a + c for a, c in zip(x, p)]
        return self.fgmres(l, b, x, self.m)

    def cycle(self, b_locals):
        nl = self.nl
        bs_ = [None] * nl
        bs_[nl - 1] = b_locals
        for l in range(nl - 1, 0, -1):
            bs_[l - 1] = self.restrict(l, bs_[l])
        x = [np.zeros_like(v) for v in bs_[0]]
        for l in range(nl - 1):
            x = self.vcycle(l, bs_[l], x)
            x = self.prolong(l + 1, x)
        return self.vcycle(nl - 1, bs_[nl - 1], x)


class PeerRecordingContext(RecordingContext):
    """peer_memory=True without CUDA IPC: the handle exchange is a no-op here."""

    def comm_peer_handle(self):
        return b"\0" * 64

    def comm_peer_open(self, handles):
        assert len(handles) == 64 * self.nranks


def run_model(prob, nranks, b, peer=False):
    levels = [level_input_from_synth(l) for l in prob.levels]
    if peer:
        import torch.distributed as dist
        mgs = []
        orig = dist.all_gather_object
        dist.all_gather_object = lambda out, obj: out.__setitem__(slice(None), [obj] * len(out))
        try:
            mgs = [DistributedMultigrid(levels, prob.config.m, r, nranks, None, ctx=PeerRecordingContext(), peer_memory=True)
                   for r in range(nranks)]
        finally:
            dist.all_gather_object = orig
    else:
        mgs = [DistributedMultigrid(levels, prob.config.m, r, nranks, None, ctx=RecordingContext(), condense=True)
               for r in range(nranks)]
    model = Lockstep([m.ctx for m in mgs])
    locs = []
    for m in mgs:
        v = m.scatter(b)
        v[m.n_owned:] = NAN                                        # only the owned part is promised on entry
        locs.append(v)
    out = model.cycle(locs)
    x = np.full(b.size, NAN)
    for m, v in zip(mgs, out):
        x[m.local_dofs[:m.n_owned]] = v[:m.n_owned]
    return x, model, mgs


@pytest.mark.parametrize("name", ["ldc2d-sv-k2-tiny", "ldc2d-pkp0-tiny", "ldc3d-sv-k3-tiny", "ldc3d-pkp0-tiny",
                                  "bfs2d-sv-k2-tiny"])
@pytest.mark.parametrize("nranks,peer", [(1, False), (2, False), (3, False), (2, True), (4, True)])
def test_device_sequence_on_local_data_equals_serial_oracle(problems, name, nranks, peer):
    """peer: the exchanges in their packed peer-memory form (slots, peer offsets, gather form of the sum)."""
    prob = problems(name, gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in prob.levels]
    b = np.random.default_rng(5).standard_normal(lv[-1].n)
    b[lv[-1].bc_dofs] = 0
    want = hp.fcycle(lv, b, prob.config.m)
    got, model, mgs = run_model(prob, nranks, b, peer)
    assert np.isfinite(got).all(), "a gather read a ghost entry that no step had refreshed"
    assert np.linalg.norm(got - want) <= 1e-10 * np.linalg.norm(want)
    if nranks > 1:
        assert model.stats["update"] > 0 and model.stats["reduce"] > 0
        # owned sets partition the dofs
        owned = np.concatenate([m.local_dofs[:m.n_owned] for m in mgs])
        assert np.array_equal(np.sort(owned), np.arange(b.size))


def test_exchange_count_per_krylov_iteration(problems):
    """Two owner->ghost updates, one ghost->owner sum and two small all-reduces per FGMRES iteration."""
    prob = problems("ldc2d-sv-k2-tiny", gamma=10.0, nu=0.2)
    levels = [level_input_from_synth(l) for l in prob.levels]
    mgs = [DistributedMultigrid(levels, prob.config.m, r, 2, None, ctx=RecordingContext()) for r in range(2)]
    model = Lockstep([m.ctx for m in mgs])
    rng = np.random.default_rng(0)
    b = rng.standard_normal(levels[-1].n_nodes * levels[-1].bs)
    m = prob.config.m
    model.fgmres(1, [g.scatter(b) for g in mgs], [np.zeros(g.local_dofs.size) for g in mgs], m)
    assert model.stats == {"update": 2 * m + 1, "reduce": m, "allreduce": 2 * m + 1}


@pytest.mark.parametrize("name,shape", [("ldc2d-sv-k2-tiny", (2, 1)), ("ldc2d-sv-k2-tiny", (2, 2)), ("ldc3d-sv-k3-tiny", (2, 1, 1)),
                                        ("ldc2d-pkp0-tiny", (3, 1)), ("ldc3d-pkp0-tiny", (1, 2, 1))])
@pytest.mark.parametrize("peer", [False, True])
def test_rank_locally_generated_problem_through_the_device_sequence(name, shape, peer):
    """alfi_b200.synth.bricks.build_rank_local -> DistributedMultigrid.from_local -> the device step sequence on
    every rank's own data == the serial oracle of the globally generated box problem (matched through the nodes'
    lattice keys).  No rank ever sees a global level >= 1."""
    import dataclasses

    import torch.distributed as dist

    from alfi_b200.synth.bricks import build_rank_local, node_keys
    from alfi_b200.synth.problem import CONFIGS, build_problem
    cfg = dataclasses.replace(CONFIGS[name], shape=shape)
    nranks = int(np.prod(shape))
    glob = build_problem(cfg, gamma=10.0, nu=0.2)
    lv = [hp.level_from_host(l) for l in glob.levels]
    b = np.random.default_rng(9).standard_normal(lv[-1].n)
    b[lv[-1].bc_dofs] = 0
    want = hp.fcycle(lv, b, cfg.m)
    probs = [build_rank_local(cfg, r, nu=0.2, gamma=10.0) for r in range(nranks)]
    gathered = []
    orig = dist.all_gather_object

    def fake_all_gather(out, obj):
        gathered.append(obj)
        # the lock-step construction below calls every rank in turn: hand each caller the full list once known
        out[:] = [obj] * len(out) if isinstance(obj, bytes) or fake_all_gather.everyone is None else fake_all_gather.everyone
    fake_all_gather.everyone = None
    if peer:
        from alfi_b200.lib import Context
        fake_all_gather.everyone = [{l: (v[0], v[1].tolist(), v[2].tolist()) for l, v in
                                     {l: Context.halo_peer_list(p.local[l].send, p.local[l].recv)
                                      for l in range(1, len(p.local))}.items()} for p in probs]
    dist.all_gather_object = fake_all_gather
    try:
        mgs = [DistributedMultigrid.from_local(p, cfg.m, None, ctx=(PeerRecordingContext() if peer else RecordingContext()),
                                               peer_memory=peer) for p in probs]
    finally:
        dist.all_gather_object = orig
    model = Lockstep([m.ctx for m in mgs])
    L = len(lv) - 1
    bs = glob.finest.V.bs
    _, gkey = node_keys(glob.finest.V.node_coords, cfg.N * 2 ** L, cfg.length, shape)
    order = np.argsort(gkey)
    g_of, locs = [], []
    for p in probs:
        pos = np.searchsorted(gkey[order], p.keys[L])
        g = (order[pos][:, None] * bs + np.arange(bs)[None, :]).ravel()
        v = b[g].copy()
        v[p.local[L].n_owned:] = NAN
        g_of.append(g)
        locs.append(v)
    out = model.cycle(locs)
    got = np.full(b.size, NAN)
    for p, g, v in zip(probs, g_of, out):
        no = p.local[L].n_owned
        got[g[:no]] = v[:no]
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= 1e-10 * np.linalg.norm(want)
