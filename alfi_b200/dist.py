"""Multi-GPU host logic: which rank owns which patch, and communicator bootstrap.

The reference partitions the mesh with DMPlex and lets every MPI rank build the patches of its
*owned* vertices (ghost exclusion, alfi/relaxation.py:120-121; overlap alfi/solver.py:604-605,
661-662).  Here one rank drives one GPU of a single NVSwitch box; patches are the sharded units
(they carry 8 n_i^2 bytes each and dominate both memory and time), level vectors stay replicated
and the library sums the per-rank contributions with one ncclAllReduce per application
(csrc/comm.cu).  `partition_patches` assigns contiguous runs of patches — ordered by where they
sit in the dof numbering, i.e. spatial slabs — so that every rank streams the same number of
factor bytes.
"""
from __future__ import annotations

import numpy as np

__all__ = ["partition_patches", "condensed_cost", "shard_patch_arrays", "shard_dof_array", "bootstrap_unique_id"]


def condensed_cost(offsets, blocks) -> np.ndarray:
    """Per-patch estimate of the bytes a condensed patch streams per application (csrc/condense.cu): the
    separator inverse |S|^2 plus, per block, D and the two coupling tiles (~3 b_k^2).  Only used to balance
    the partition, so a proxy is enough."""
    offsets = np.asarray(offsets, dtype=np.int64)
    blocks = np.asarray(blocks)
    npatch = offsets.size - 1
    patch_of = np.repeat(np.arange(npatch), np.diff(offsets))
    sep = np.bincount(patch_of[blocks < 0], minlength=npatch).astype(np.float64)
    cost = sep ** 2
    inb = blocks >= 0
    if inb.any():
        key = patch_of[inb].astype(np.int64) * (int(blocks.max()) + 1) + blocks[inb]
        uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        per_block_patch = (uniq // (int(blocks.max()) + 1)).astype(np.int64)
        cost += np.bincount(per_block_patch, weights=3.0 * cnt.astype(np.float64) ** 2, minlength=npatch)
    return cost


def partition_patches(offsets, dofs, nranks: int, cost=None) -> np.ndarray:
    """owner[p] in [0, nranks): contiguous in the patches' mean dof index, balanced by `cost` per patch
    (default n_p^2, the bytes of a dense inverse; pass `condensed_cost` for condensed sets)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    npatch = offsets.size - 1
    n = np.diff(offsets)
    if nranks == 1 or npatch == 0:
        return np.zeros(npatch, dtype=np.int32)
    csum = np.concatenate(([0.0], np.cumsum(np.asarray(dofs, dtype=np.float64))))
    mean = (csum[offsets[1:]] - csum[offsets[:-1]]) / np.maximum(n, 1)
    order = np.argsort(mean, kind="stable")
    cost = (n[order].astype(np.float64)) ** 2 if cost is None else np.asarray(cost, dtype=np.float64)[order]
    cum = np.cumsum(cost)
    total = cum[-1] if cum.size else 0.0
    # patch k goes to the rank whose share contains the midpoint of its cost interval
    mid = cum - 0.5 * cost
    owner_sorted = np.minimum((mid / max(total, 1e-300) * nranks).astype(np.int64), nranks - 1)
    owner = np.empty(npatch, dtype=np.int32)
    owner[order] = owner_sorted
    return owner


def shard_patch_arrays(offsets, dofs, order, colours, owner, rank):
    """Subset of (offsets, dofs, order, colours) owned by `rank`, order re-indexed locally."""
    offsets = np.asarray(offsets, dtype=np.int64)
    mine = np.flatnonzero(owner == rank)
    n = np.diff(offsets)[mine]
    new_off = np.concatenate(([0], np.cumsum(n))).astype(np.int64)
    idx = np.concatenate([np.arange(offsets[p], offsets[p + 1]) for p in mine]) if mine.size else np.empty(0, np.int64)
    new_dofs = np.asarray(dofs)[idx].astype(np.int32)
    local = np.full(offsets.size - 1, -1, dtype=np.int64)
    local[mine] = np.arange(mine.size)
    if order is None:
        order = np.arange(offsets.size - 1)
    order = np.asarray(order)
    new_order = local[order[owner[order] == rank]].astype(np.int32)
    new_col = None if colours is None else np.asarray(colours)[mine].astype(np.int32)
    return new_off, new_dofs, new_order, new_col, mine


def shard_dof_array(offsets, arr, mine):
    """Entries of a per-patch-dof array (e.g. the condensation block labels) of the patches `mine`."""
    if arr is None:
        return None
    offsets = np.asarray(offsets, dtype=np.int64)
    idx = np.concatenate([np.arange(offsets[p], offsets[p + 1]) for p in mine]) if len(mine) else np.empty(0, np.int64)
    return np.asarray(arr)[idx]


def bootstrap_unique_id(rank: int):
    """ncclUniqueId made on rank 0 and broadcast with torch.distributed (mpi4py in a deployment)."""
    import torch.distributed as dist
    from .lib import Context
    box = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]
