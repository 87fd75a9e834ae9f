"""Synthetic clustered operators for the edge cases of the condensed patch sets (shared by the CPU
checker tests/test_condense_host.py and the GPU test tests/test_gpu_edges.py).

Node graph: clusters of fully coupled nodes (the future blocks: 1 .. 64 dofs), separator nodes that
couple to some clusters and to each other, and no cluster-cluster coupling.  Patches are unions of
clusters and separator nodes chosen to hit: empty patch, separator-only patch, block-only patch
(no separator), a block without separator neighbours inside the patch, odd and even sizes, blocks at
the 64-dof limit, overlapping patches (several colours) and a repeated entry in the iteration set.
"""
import numpy as np
import scipy.sparse as sp


def clustered_problem(bs, seed=0):
    rng = np.random.default_rng(seed)
    max_nodes = 64 // bs
    cluster_sizes = [1, 2, 5, max_nodes, max_nodes - 1, 7, 3, 11]           # nodes per cluster
    nsepn = 40
    starts = np.concatenate(([0], np.cumsum(cluster_sizes)))
    ncl = len(cluster_sizes)
    sep0 = starts[-1]
    n_nodes = int(sep0 + nsepn)
    rows, cols = [], []
    for k in range(ncl):
        nodes = np.arange(starts[k], starts[k + 1])
        rr, cc = np.meshgrid(nodes, nodes, indexing="ij")
        rows.append(rr.ravel())
        cols.append(cc.ravel())
    # separator-separator: random sparse + diagonal
    S = sp.random(nsepn, nsepn, density=0.3, random_state=seed, format="coo")
    rows += [sep0 + S.row, sep0 + S.col, sep0 + np.arange(nsepn)]
    cols += [sep0 + S.col, sep0 + S.row, sep0 + np.arange(nsepn)]
    # cluster-separator couplings (symmetric pattern); cluster 6 couples to no separator at all
    nb = {}
    for k in range(ncl):
        if k == 6:
            nb[k] = np.empty(0, np.int64)
            continue
        cnt = [3, 21 if bs == 3 else 32, 6, 10, 4, 8, 0, 12][k]
        nb[k] = np.sort(rng.choice(nsepn, size=min(cnt, nsepn), replace=False))
        nodes = np.arange(starts[k], starts[k + 1])
        rr, cc = np.meshgrid(nodes, sep0 + nb[k], indexing="ij")
        rows += [rr.ravel(), cc.ravel()]
        cols += [cc.ravel(), rr.ravel()]
    pat = sp.csr_matrix((np.ones(sum(r.size for r in rows)), (np.concatenate(rows), np.concatenate(cols))),
                        shape=(n_nodes, n_nodes))
    pat.sum_duplicates()
    pat.sort_indices()
    rowptr, colidx = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
    vals = rng.standard_normal((colidx.size, bs, bs))
    r_of = np.repeat(np.arange(n_nodes), np.diff(rowptr))
    vals[r_of == colidx] += 3.0 * np.sqrt(n_nodes * bs) * np.eye(bs)         # comfortably non-singular
    A = sp.bsr_matrix((vals, colidx, rowptr), shape=(n_nodes * bs,) * 2).tocsr()

    def dofs_of(nodes):
        return (np.asarray(nodes, dtype=np.int64)[:, None] * bs + np.arange(bs)[None, :]).ravel()

    def cl(k):
        return np.arange(starts[k], starts[k + 1])

    sepn = sep0 + np.arange(nsepn)
    # (cluster ids, separator nodes) per patch
    spec = [
        ([], []),                                   # empty patch
        ([], sepn[:9]),                             # separators only (odd count of nodes)
        ([0, 2], []),                               # blocks only, no separator in the patch
        ([6, 5], sepn[nb[5][:3]]),                  # block 6 has no separator neighbour (m = 0)
        ([3, 4], sepn[np.union1d(nb[3], nb[4])]),   # blocks at the 64-dof limit
        ([1], sepn[nb[1]]),                         # neighbourhood at the 64-dof limit
        ([0, 1, 2, 5, 7], sepn),                    # everything coupled to everything, > 64 separator dofs
        ([7, 2], sepn[::2]),                        # overlaps the others: needs colours
    ]
    patches, blocks = [], []
    for cls, sn in spec:
        d, b = [], []
        # interleave: separator dofs first for odd patches, last for even ones (local order is arbitrary)
        parts = [(dofs_of(cl(k)), k) for k in cls]
        parts.insert(len(parts) // 2, (dofs_of(sn) if len(sn) else np.empty(0, np.int64), -1))
        for dd, lab in parts:
            d.append(dd)
            b.append(np.full(dd.size, lab if lab < 0 else 100 + lab, dtype=np.int32))
        d, b = np.concatenate(d), np.concatenate(b)
        perm = rng.permutation(d.size)              # scrambled patch-local order
        patches.append(d[perm].astype(np.int32))
        blocks.append(b[perm])
    sizes = np.array([p.size for p in patches])
    offsets = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    return dict(n_nodes=n_nodes, bs=bs, rowptr=rowptr, colidx=colidx, vals=vals, A=A, patches=patches,
                offsets=offsets, dofs=np.concatenate(patches), blocks=np.concatenate(blocks), sizes=sizes)


def dense_reference(case, order, x, bc=None):
    A = case["A"]
    y = np.zeros_like(x)
    for p in order:
        I = case["patches"][p]
        if I.size:
            y[I] += np.linalg.solve(A[I][:, I].toarray(), x[I])
    if bc is not None:
        y[bc] = x[bc]
    return y


def greedy_colours(case, order):
    used = {}
    colours = np.zeros(len(case["patches"]), np.int32)
    seen = set()
    for p in order:
        if p in seen:
            continue
        seen.add(p)
        taken_all = set()
        for d in case["patches"][p]:
            taken_all |= used.get(d, set())
        c = 0
        while c in taken_all:
            c += 1
        colours[p] = c
        for d in case["patches"][p]:
            used.setdefault(d, set()).add(c)
    return colours
