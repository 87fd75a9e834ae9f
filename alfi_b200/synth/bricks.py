"""Rank-local generation of a box-shaped Kuhn problem (`Config.shape` = the rank grid): every rank builds ONLY its
brick plus a halo, never the global fine mesh — what a DMPlex partition hands each MPI rank in the reference
(alfi/solver.py:604-605, 661-662: overlap of one (macro) vertex star), and what makes the weak-scaling family
`ldc3d-sv-k3-w{2,4,8}` generable (the global generator needs ~17 GB and ~70 s per 1.46 M dofs, on every rank).

Rank r = bx + sx (by + sy bz) owns the brick [b_a, b_a + 1] * length of the box.  For every level l >= 1 it builds a
two-level Kuhn hierarchy over the brick grown by ONE cell of level l-1 on every side that has a neighbour (= two
cells of level l), assembles the operator, the patches, the robust transfer and the standard prolongation there with
the ordinary generator, and cuts the result down to its local set:

* ownership is geometric — a node (vertex, cell) belongs to the lowest brick whose closure contains it;
* the local set K of a level = the nodes in the closed box "brick + one cell of that level" (owned nodes first);
  every owned patch, every operator row of an owned node and every owned cell patch lies inside it;
* exchange lists need no communication: the peer's box is known, both sides sort the shared nodes by their integer
  lattice key (coordinates in units of h / 12);
* `P_H` of level l reads level l-1 in that level's own local set (its box is exactly the coarse box of level l), so the
  transfer halo of level l is the halo of level l-1; level 0 (the whole coarse box mesh) is replicated.

`build_rank_local` returns what `alfi_b200.multigrid.DistributedMultigrid.from_local` hands to the library;
tests/test_bricks.py checks, on small boxes, that the distributed cycle on these data equals the serial oracle of the
globally generated problem.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from ..halo import Layout, LocalLevel, RankLayout, local_level
from ..patches import macro_interior_blocks
from ..transfer import cell_patch_set
from .fem import BlockPattern, VectorSpace
from .hierarchy import build_hierarchy, prolongation_matrix
from .problem import Config, LevelData, assemble_level, assemble_transfer, lid_wind, smoother_patches

__all__ = ["RankLocalProblem", "build_rank_local", "node_keys", "brick_of_rank"]

KEY_UNITS = 12            # lattice units per cell edge: P3 / Alfeld node coordinates are multiples of h / 12


def brick_of_rank(rank: int, shape: tuple) -> tuple:
    out = []
    for s in shape:
        out.append(rank % s)
        rank //= s
    return tuple(out)


def rank_of_brick(b, shape) -> np.ndarray:
    b = np.asarray(b)
    r = np.zeros(b.shape[:-1], dtype=np.int64)
    mul = 1
    for a, s in enumerate(shape):
        r = r + mul * b[..., a]
        mul *= s
    return r


def node_keys(coords: np.ndarray, cells_per_length: int, length: float, shape: tuple):
    """(integer lattice coordinates (n, d), one int64 key per node): positions in units of h / 12 of this level."""
    q = np.rint(coords * (KEY_UNITS * cells_per_length / length)).astype(np.int64)
    key = np.zeros(q.shape[0], dtype=np.int64)
    mul = 1
    for a, s in enumerate(shape):
        key = key + mul * q[:, a]
        mul *= KEY_UNITS * cells_per_length * s + 1
    return q, key


def owner_brick(q: np.ndarray, brick_units: int, shape: tuple) -> np.ndarray:
    """Lowest brick whose closure contains the lattice point(s) q (n, d) -> brick indices (n, d)."""
    b = -(-q // brick_units) - 1                       # ceil(q / L) - 1: an interface point goes to the lower brick
    return np.clip(b, 0, np.asarray(shape)[None, :] - 1)


@dataclass
class RankLocalProblem:
    config: Config
    rank: int
    nranks: int
    level0: object                          # LevelInput of the replicated coarsest level (global numbering)
    local: list                             # [None] + LocalLevel per level >= 1
    keys: list                              # per level: int64 key of every local node (level 0: of every global node)
    n_global: list = field(default_factory=list)   # global dof count per level (bookkeeping only)


def _expand(nodes, bs):
    return (np.asarray(nodes, dtype=np.int64)[:, None] * bs + np.arange(bs)[None, :]).ravel()


def build_rank_local(cfg: Config, rank: int, nu: float | None = None, gamma: float | None = None,
                     verbose: bool = False) -> RankLocalProblem:
    import dataclasses
    import time

    from ..multigrid import level_input_from_synth
    if gamma is not None:
        cfg = dataclasses.replace(cfg, gamma=gamma)
    nu = cfg.nu if nu is None else nu
    d = cfg.dim
    shape = tuple(cfg.shape) if cfg.shape else (1,) * d
    nranks = int(np.prod(shape))
    assert cfg.domain == "ldc" and 0 <= rank < nranks
    b = np.asarray(brick_of_rank(rank, shape))
    gext = cfg.length * np.asarray(shape, dtype=np.float64)
    t0 = time.time()

    # ---- level 0: the whole (small) coarse box, replicated
    h0 = build_hierarchy(d, cfg.N, 0, cfg.bary, cfg.length, shape)[0]
    V0 = VectorSpace(h0.mesh, cfg.k, cfg.element)
    ld0 = LevelData(0, h0, V0, BlockPattern(V0), V0.boundary_nodes().astype(np.int32))
    assemble_level(cfg, ld0, nu, cfg.gamma)
    level0 = level_input_from_synth(ld0)
    _, key0 = node_keys(V0.node_coords, cfg.N, cfg.length, shape)
    keys = [key0]
    n_global = [V0.ndofs]
    local = [None]
    prev = None                                   # (sorted keys, local node ids) of level l-1's local set
    bs = V0.bs

    for l in range(1, cfg.nref + 1):
        nc, nf = cfg.N * 2 ** (l - 1), cfg.N * 2 ** l          # cells per brick edge on levels l-1, l
        lo = [int(b[a] * nc - (1 if b[a] > 0 else 0)) for a in range(d)]
        hi = [int((b[a] + 1) * nc + (1 if b[a] < shape[a] - 1 else 0)) for a in range(d)]
        hier = build_hierarchy(d, nc, 1, cfg.bary, cfg.length, (), tuple(h - o for h, o in zip(hi, lo)), tuple(lo))
        Vc = VectorSpace(hier[0].mesh, cfg.k, cfg.element)
        V = VectorSpace(hier[1].mesh, cfg.k, cfg.element)
        x = V.node_coords
        on_bdry = np.any((np.abs(x) < 1e-12) | (np.abs(x - gext[None, :]) < 1e-12), axis=1)
        ld = LevelData(1, hier[1], V, BlockPattern(V), np.flatnonzero(on_bdry).astype(np.int32))
        assemble_level(cfg, ld, nu, cfg.gamma, wind=V.interpolate(lambda xx: lid_wind(xx, gext)))
        ld.patches = smoother_patches(cfg, ld)
        if cfg.element == "p1fb":
            from ..bubble import bubble_transfer_matrix
            ld.P = bubble_transfer_matrix(Vc, V, hier[0].c2f)
            ld.P_dof_level = True
        else:
            ld.P = prolongation_matrix(Vc, V, hier[0].c2f)
        ld.cell_patches, ld.cb_nodes = cell_patch_set(hier, 1, V, cfg.bary)
        if cfg.bary:
            ld.cell_patches.blocks = macro_interior_blocks(hier[1].plex, V, ld.cell_patches)
        assemble_transfer(cfg, ld, nu, cfg.gamma)
        li = level_input_from_synth(ld)

        # ---- geometry: owners, local set, exchange lists (all in the numbering of this sub-box)
        q, key = node_keys(x, nf, cfg.length, shape)
        Lb = KEY_UNITS * nf                                         # brick edge in lattice units
        own_b = owner_brick(q, Lb, shape)
        owner_node = rank_of_brick(own_b, shape)
        box_lo = b * Lb - KEY_UNITS
        box_hi = (b + 1) * Lb + KEY_UNITS
        in_K = np.all((q >= box_lo[None, :]) & (q <= box_hi[None, :]), axis=1)
        owned_nodes = np.flatnonzero(owner_node == rank)
        assert in_K[owned_nodes].all()
        ghost_nodes = np.flatnonzero(in_K & (owner_node != rank))
        send, recv = {}, {}
        for peer in np.unique(owner_node[ghost_nodes]):
            sel = np.flatnonzero(owner_node[ghost_nodes] == peer)
            sel = sel[np.argsort(key[ghost_nodes[sel]], kind="stable")]
            recv[int(peer)] = (sel[:, None] * bs + np.arange(bs)[None, :]).ravel()      # positions in the ghost dof list
        for peer in range(nranks):
            if peer == rank:
                continue
            pb = np.asarray(brick_of_rank(peer, shape))
            if np.any(np.abs(pb - b) > 1):
                continue
            plo, phi = pb * Lb - KEY_UNITS, (pb + 1) * Lb + KEY_UNITS
            qo = q[owned_nodes]
            sel = np.flatnonzero(np.all((qo >= plo[None, :]) & (qo <= phi[None, :]), axis=1))
            if sel.size:
                sel = sel[np.argsort(key[owned_nodes[sel]], kind="stable")]
                send[int(peer)] = (sel[:, None] * bs + np.arange(bs)[None, :]).ravel()  # positions in the owned dof list
        # patches of owned vertices (iteration order kept), cell patches of owned coarse cells
        ps = ld.patches
        pq = np.rint(ps.centres * (KEY_UNITS * nf / cfg.length)).astype(np.int64)
        powner = rank_of_brick(owner_brick(pq, Lb, shape), shape)
        mine = np.asarray([p for p in ps.order if powner[p] == rank], dtype=np.int64)
        cp = ld.cell_patches
        cowner = np.zeros(cp.npatch, dtype=np.int64)
        node_of = np.asarray(cp.dofs, dtype=np.int64) // bs
        for c in range(cp.npatch):
            nodes = node_of[cp.offsets[c]:cp.offsets[c + 1]]
            if nodes.size:
                centre = np.floor(q[nodes].mean(axis=0) / Lb).astype(np.int64)      # strictly inside one brick
                cowner[c] = rank_of_brick(np.clip(centre, 0, np.asarray(shape) - 1), shape)
        rl = RankLayout(rank, _expand(owned_nodes, bs), _expand(ghost_nodes, bs), mine, send, recv)
        ranks = [None] * nranks
        ranks[rank] = rl
        owner_dof = np.repeat(owner_node, bs)
        lay = Layout(nranks, V.ndofs, owner_dof, ranks, extra_owner=cowner)
        ll = local_level(li, lay, rank, None)
        # ---- columns of P_H: level l-1 in its own local numbering (l >= 2) or the global level 0 (l == 1)
        _, keyc = node_keys(Vc.node_coords, nc, cfg.length, shape)
        if l == 1:
            order0 = np.argsort(key0)
            pos = np.searchsorted(key0[order0], keyc)
            assert (key0[order0][np.minimum(pos, key0.size - 1)] == keyc).all()
            cmap_nodes, ncols = order0[pos], V0.ndofs
        else:
            pk, pid = prev
            pos = np.searchsorted(pk, keyc)
            hit = (pos < pk.size) & (pk[np.minimum(pos, pk.size - 1)] == keyc)
            cmap_nodes = np.where(hit, pid[np.minimum(pos, pk.size - 1)], -1)
            ncols = local[l - 1].n_local
        P = ll.P.tocoo()
        cnode, ccomp = P.col // bs, P.col % bs
        assert (cmap_nodes[cnode] >= 0).all(), "a coarse node P_H reads is not in the coarser level's local set"
        ll.P = sp.csr_matrix((P.data, (P.row, cmap_nodes[cnode] * bs + ccomp)), shape=(ll.n_owned, ncols))
        ll.coarse_local = None
        local.append(ll)
        loc_nodes = ll.local_dofs[::bs] // bs                       # sub-box node of every local node
        lk = key[loc_nodes]
        o = np.argsort(lk)
        prev = (lk[o], o.astype(np.int64))
        keys.append(lk)
        n_global.append(None)
        if verbose:
            print("[bricks] rank %d level %d: sub-box %s cells, %d local dofs (%d owned), %d patches, %.1fs" % (
                rank, l, tuple(2 * (h - o_) for h, o_ in zip(hi, lo)), ll.n_local, ll.n_owned, mine.size, time.time() - t0), flush=True)
    return RankLocalProblem(cfg, rank, nranks, level0, local, keys, n_global)
