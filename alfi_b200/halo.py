"""Owned / ghost layout of a level for distributed vectors (host side, numpy).

The library's multi-GPU path of round 1 replicates every level vector and pays an all-reduce of the whole
vector after each patch application and an all-gather after each SpMV.  This module is the index
machinery of the replacement SURVEY §8(e) asks for — the PetscSF pattern of the reference's PCPATCH / MatMult
on a vertex-partitioned DMPlex (alfi/solver.py:604-605,661-662: overlap VERTEX,1 or VERTEX,2):

* patches are the owned unit (`alfi_b200.dist.partition_patches`);
* every dof has exactly one owner — the rank owning the first patch (in partition order) that contains it;
  dofs outside every patch (Dirichlet dofs) follow the first dof they are coupled to in the operator pattern;
* a rank's *local* vector = its owned dofs followed by its ghosts: every other dof touched by its patches or
  by the operator rows of its owned dofs;
* owner -> ghost update ("broadcast", before SpMV / patch gather) and ghost -> owner sum ("reduce", after the
  patch scatter-add) move only the ghost entries, between the pairs of ranks that share them.

`Layout.exchange_bytes()` is the traffic model quoted in DESIGN §6.  oracle/distributed.py executes a level
smoother on these layouts rank by rank and checks it against the serial oracle; the gloo test does the same
with real processes.  The device side (NCCL send/recv or NVLink peer stores driven by these lists) is the next
step and is not part of the round-1 library.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = ["Layout", "RankLayout", "LocalLevel", "build_layout", "layout_from_owner", "transfer_halo", "local_level"]


@dataclass
class RankLayout:
    rank: int
    owned: np.ndarray            # global dofs owned by this rank, ascending
    ghost: np.ndarray            # global dofs held as ghosts, ascending
    patches: np.ndarray          # global patch ids owned by this rank (iteration order preserved)
    send: dict                   # peer -> positions in `owned` whose values the peer holds as ghosts
    recv: dict                   # peer -> positions in `ghost` owned by the peer (same order as the peer's send)

    @property
    def local(self):
        return np.concatenate([self.owned, self.ghost])

    @property
    def n_owned(self):
        return self.owned.size

    @property
    def n_local(self):
        return self.owned.size + self.ghost.size


@dataclass
class Layout:
    nranks: int
    ndofs: int
    owner: np.ndarray            # owner rank of every dof
    ranks: list
    extra_owner: np.ndarray | None = None   # owner rank of every extra set (cell patch), if any

    def exchange_bytes(self):
        """(max, total) bytes one owner->ghost update (equivalently one ghost->owner reduce) moves per rank."""
        per_rank = [8 * sum(v.size for v in r.recv.values()) for r in self.ranks]
        return max(per_rank), sum(per_rank)

    def neighbours(self):
        return [sorted(r.recv) for r in self.ranks]

    # ---- the two exchange steps, on a list of per-rank local arrays (reference implementation) ----
    def update_ghosts(self, locs):
        """owner -> ghost: every ghost entry takes its owner's value."""
        for r in self.ranks:
            for peer, pos in r.recv.items():
                src = self.ranks[peer]
                locs[r.rank][r.n_owned + pos] = locs[peer][src.send[r.rank]]

    def reduce_ghosts(self, locs):
        """ghost -> owner: ghost contributions are added to the owner's entry (peers in ascending rank order, so the
        sum is reproducible) and the ghost copies are cleared."""
        for r in self.ranks:
            for peer in sorted(r.send):
                src = self.ranks[peer]
                locs[r.rank][r.send[peer]] += locs[peer][src.n_owned + src.recv[r.rank]]
        for r in self.ranks:
            locs[r.rank][r.n_owned:] = 0.0

    def scatter(self, x):
        """Global vector -> per-rank local arrays with consistent ghosts."""
        return [x[r.local].copy() for r in self.ranks]

    def gather(self, locs):
        out = np.empty(self.ndofs)
        for r in self.ranks:
            out[r.owned] = locs[r.rank][:r.n_owned]
        return out


def layout_from_owner(owner, needs, patches=None) -> Layout:
    """Layout for a given dof ownership and, per rank, the set of dofs it must hold locally (`needs[r]`, any
    order, owned entries allowed): ghosts = needs minus owned; exchange lists follow."""
    owner = np.asarray(owner, dtype=np.int64)
    ndofs = owner.size
    nranks = len(needs)
    ranks = []
    for r in range(nranks):
        owned = np.flatnonzero(owner == r)
        need = np.unique(np.asarray(needs[r], dtype=np.int64)) if len(needs[r]) else np.empty(0, np.int64)
        ghost = need[owner[need] != r]
        mine = np.empty(0, np.int64) if patches is None else patches[r]
        ranks.append(RankLayout(r, owned, ghost, mine, {}, {}))
    pos_in_owned = np.empty(ndofs, dtype=np.int64)
    for r in ranks:
        pos_in_owned[r.owned] = np.arange(r.owned.size)
    for r in ranks:
        go = owner[r.ghost]
        for peer in np.unique(go):
            sel = np.flatnonzero(go == peer)
            r.recv[int(peer)] = sel
            ranks[int(peer)].send[r.rank] = pos_in_owned[r.ghost[sel]]
    return Layout(nranks, ndofs, owner, ranks)


def transfer_halo(P, fine: Layout, coarse_owner) -> Layout:
    """Layout on the COARSE level for the standard prolongation `P` (fine dofs x coarse dofs, scipy CSR): a rank
    needs the coarse dofs in the columns of its owned fine rows.  Owned sets are those of `coarse_owner`; prolong
    uses `update_ghosts` on it, restrict (P^T) ends with `reduce_ghosts` on it."""
    needs = [np.unique(P[r.owned].indices) for r in fine.ranks]
    return layout_from_owner(coarse_owner, needs)


def build_layout(patch_offsets, patch_dofs, patch_order, patch_owner, rowptr, colidx, bs, ndofs,
                 extra_sets=None, nranks=None) -> Layout:
    """`patch_owner[p]` = rank of patch p (alfi_b200.dist.partition_patches); `rowptr/colidx` = the level's block
    pattern (block rows of `bs` dofs).  `extra_sets` = (offsets, dofs) of further index sets that must be local
    to one rank each (the transfer's cell patches): a set goes to the owner of its first dof."""
    patch_offsets = np.asarray(patch_offsets, dtype=np.int64)
    patch_dofs = np.asarray(patch_dofs, dtype=np.int64)
    patch_owner = np.asarray(patch_owner)
    if nranks is None:
        nranks = int(patch_owner.max()) + 1 if patch_owner.size else 1
    npatch = patch_offsets.size - 1
    order = np.arange(npatch) if patch_order is None else np.asarray(patch_order)
    # ---- dof ownership: first containing patch in iteration order
    owner = np.full(ndofs, -1, dtype=np.int64)
    for p in order[::-1]:                                   # later writes win -> iterate backwards
        owner[patch_dofs[patch_offsets[p]:patch_offsets[p + 1]]] = patch_owner[p]
    nodes = rowptr.size - 1
    node_of = np.arange(ndofs) // bs
    # dofs in no patch: take the owner of the first coupled dof that has one (repeat until settled)
    for _ in range(8):
        missing = np.flatnonzero(owner < 0)
        if missing.size == 0:
            break
        node_owner = np.full(nodes, -1, dtype=np.int64)
        have = owner >= 0
        np.maximum.at(node_owner, node_of[have], owner[have])
        for g in missing:
            nb = colidx[rowptr[node_of[g]]:rowptr[node_of[g] + 1]]
            cand = node_owner[nb]
            cand = cand[cand >= 0]
            if cand.size:
                owner[g] = cand[0]
    owner[owner < 0] = 0
    # ---- per rank: what must be local (patch dofs + operator columns of owned rows + its extra sets)
    needs, mine_all = [], []
    extra_owner = None
    if extra_sets is not None:
        eoff, edofs = np.asarray(extra_sets[0], dtype=np.int64), np.asarray(extra_sets[1], dtype=np.int64)
        first = edofs[np.minimum(eoff[:-1], max(edofs.size - 1, 0))] if edofs.size else np.empty(0, np.int64)
        extra_owner = np.where(np.diff(eoff) > 0, owner[first], 0) if edofs.size else np.zeros(eoff.size - 1, np.int64)
    for r in range(nranks):
        owned = np.flatnonzero(owner == r)
        mine = order[patch_owner[order] == r]
        need = [patch_dofs[patch_offsets[p]:patch_offsets[p + 1]] for p in mine]
        own_nodes = np.unique(node_of[owned])
        cols = np.unique(np.concatenate([colidx[rowptr[n]:rowptr[n + 1]] for n in own_nodes])) if own_nodes.size else np.empty(0, np.int64)
        need.append((cols[:, None] * bs + np.arange(bs)[None, :]).ravel())
        if extra_owner is not None:
            for q in np.flatnonzero(extra_owner == r):
                need.append(edofs[eoff[q]:eoff[q + 1]])
        needs.append(np.concatenate(need) if need else np.empty(0, np.int64))
        mine_all.append(mine)
    lay = layout_from_owner(owner, needs, mine_all)
    lay.extra_owner = extra_owner
    return lay


# ------------------------------------------------------------------------------------------ rank-local inputs
@dataclass
class LocalLevel:
    """One rank's share of a level in LOCAL numbering (owned nodes first, then ghosts) — what a distributed-vector
    library instance is handed instead of the global `alfi_b200.multigrid.LevelInput`.

    Block rows exist for every local node — complete for the owned ones (SpMV runs over the first `n_owned_nodes`
    rows), restricted to local columns for the ghosts (read only by the patch gathers); columns, patch dofs,
    Dirichlet lists and the transfer's index sets are local ids.  `P` has one row per owned fine dof; its columns are local ids of the coarser level's transfer
    halo (`coarse_local`, global coarse dofs in that order) or global coarse dofs when the coarser level is
    replicated.  `send` / `recv` are the two exchange lists in local DOF ids per peer rank."""
    rank: int
    bs: int
    n_owned_nodes: int
    n_local_nodes: int
    local_dofs: np.ndarray            # global dof of every local dof
    rowptr: np.ndarray
    colidx: np.ndarray
    vals: np.ndarray
    bc_dofs: np.ndarray
    patch_offsets: np.ndarray
    patch_dofs: np.ndarray
    patch_order: np.ndarray
    patch_colours: np.ndarray | None
    patch_blocks: np.ndarray | None
    patch_ids: np.ndarray             # global ids of the owned patches
    send: dict
    recv: dict
    P: object | None = None
    coarse_local: np.ndarray | None = None
    cell_offsets: np.ndarray | None = None
    cell_dofs: np.ndarray | None = None
    cell_blocks: np.ndarray | None = None
    cell_ids: np.ndarray | None = None
    cb_dofs: np.ndarray | None = None
    a0_vals: np.ndarray | None = None
    d_vals: np.ndarray | None = None
    vals_sel: np.ndarray | None = None      # positions of the local blocks in the global value arrays (per-Newton refresh)

    @property
    def n_owned(self):
        return self.n_owned_nodes * self.bs

    @property
    def n_local(self):
        return self.n_local_nodes * self.bs


def _subset_ragged(offsets, data, ids):
    n = np.diff(offsets)[ids]
    off = np.concatenate(([0], np.cumsum(n))).astype(np.int64)
    idx = np.concatenate([np.arange(offsets[p], offsets[p + 1]) for p in ids]) if len(ids) else np.empty(0, np.int64)
    return off, idx


def local_level(li, layout: Layout, rank: int, halo: Layout | None = None) -> LocalLevel:
    """Rank `rank`'s LocalLevel of the global LevelInput `li` (alfi_b200.multigrid.LevelInput) for `layout`
    (built with `extra_sets` = the level's cell patches when it has a transfer).  `halo` = `transfer_halo` on the
    coarser level, or None when that level is replicated."""
    import scipy.sparse as sp
    r = layout.ranks[rank]
    bs = li.bs
    loc = r.local
    assert (loc.reshape(-1, bs)[:, 0] % bs == 0).all() and (np.diff(loc.reshape(-1, bs), axis=1) == 1).all(), \
        "ownership and ghosts must be node-wise"
    g2l = np.full(layout.ndofs, -1, dtype=np.int64)
    g2l[loc] = np.arange(loc.size)
    nodes = loc[::bs] // bs                                        # global node of every local node
    n_owned_nodes = r.n_owned // bs
    node_g2l = np.full(li.n_nodes, -1, dtype=np.int64)
    node_g2l[nodes] = np.arange(nodes.size)
    # operator rows of ALL local nodes (the patch gather needs the rows of ghost patch dofs too), columns restricted
    # to local nodes: complete for the owned rows (SpMV), truncated for ghost rows (only read inside patches)
    sel_all = np.concatenate([np.arange(li.rowptr[n], li.rowptr[n + 1]) for n in nodes]) if nodes.size else np.empty(0, np.int64)
    row_of = np.repeat(np.arange(nodes.size), (li.rowptr[nodes + 1] - li.rowptr[nodes]).astype(np.int64))
    col_all = node_g2l[li.colidx[sel_all]]
    assert (col_all[row_of < n_owned_nodes] >= 0).all(), "an operator column of an owned row is not local"
    keep = col_all >= 0
    sel, colidx = sel_all[keep], col_all[keep]
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(row_of[keep], minlength=nodes.size)))).astype(np.int32)
    bc = g2l[np.asarray(li.bc_dofs, dtype=np.int64)]
    # owned patches, local dofs, iteration order re-indexed
    mine = r.patches
    poff, pidx = _subset_ragged(np.asarray(li.patch_offsets, dtype=np.int64), li.patch_dofs, mine)
    pdofs = g2l[np.asarray(li.patch_dofs, dtype=np.int64)[pidx]]
    assert (pdofs >= 0).all()
    out = LocalLevel(
        rank=rank, bs=bs, n_owned_nodes=n_owned_nodes, n_local_nodes=nodes.size, local_dofs=loc,
        rowptr=rowptr, colidx=colidx.astype(np.int32), vals=np.asarray(li.vals)[sel], bc_dofs=np.sort(bc[bc >= 0]).astype(np.int32),
        patch_offsets=poff, patch_dofs=pdofs.astype(np.int32), patch_order=np.arange(mine.size, dtype=np.int32),
        patch_colours=None if li.patch_colours is None else np.asarray(li.patch_colours)[mine].astype(np.int32),
        patch_blocks=None if li.patch_blocks is None else np.asarray(li.patch_blocks)[pidx],
        patch_ids=mine,
        send={p: v.copy() for p, v in r.send.items()},
        recv={p: r.n_owned + v for p, v in r.recv.items()}, vals_sel=sel)
    if li.P is not None:
        P = li.P.tocsr() if li.P_dof_level else sp.kron(li.P, sp.identity(bs), format="csr")
        Pr = P[r.owned]
        if halo is None:
            out.P = Pr.tocsr()
        else:
            hl = halo.ranks[rank].local
            c2l = np.full(halo.ndofs, -1, dtype=np.int64)
            c2l[hl] = np.arange(hl.size)
            Pc = Pr.tocoo()
            assert (c2l[Pc.col] >= 0).all()
            out.P = sp.csr_matrix((Pc.data, (Pc.row, c2l[Pc.col])), shape=(r.n_owned, hl.size))
            out.coarse_local = hl
        if li.cell_offsets is not None:
            cells = np.flatnonzero(layout.extra_owner == rank)
            coff, cidx = _subset_ragged(np.asarray(li.cell_offsets, dtype=np.int64), li.cell_dofs, cells)
            cd = g2l[np.asarray(li.cell_dofs, dtype=np.int64)[cidx]]
            assert (cd >= 0).all(), "a cell patch of this rank has a dof that is not local"
            out.cell_offsets, out.cell_dofs, out.cell_ids = coff, cd.astype(np.int32), cells
            out.cell_blocks = None if li.cell_blocks is None else np.asarray(li.cell_blocks)[cidx]
            cb = g2l[np.asarray(li.cb_dofs, dtype=np.int64)]
            out.cb_dofs = np.sort(cb[cb >= 0]).astype(np.int32)
            out.a0_vals, out.d_vals = np.asarray(li.a0_vals)[sel], np.asarray(li.d_vals)[sel]
    return out
