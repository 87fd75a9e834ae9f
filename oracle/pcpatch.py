"""CPU ORACLE — test infrastructure only.  Literal, loop-based restatement of PCPATCH's
topology -> dof-set construction (selected by alfi/solver.py:318-344 and
alfi/transfer.py:100-113).  PETSc's pcpatch.c is not under /root/reference and cannot be built
here, so this follows the published semantics ("parity unpinned"); it is the checker for the
vectorised builder in alfi_b200/patches.py.
"""
from __future__ import annotations

import numpy as np


def patch_dofs(plex, V, point_sets, bc_nodes):
    """For every user point set ``ht``:
       cht   = closure of every cell in the star of a point of ht
       cells = cells of cht (ascending)
       dofs  = dofs attached to points of ht, minus global Dirichlet dofs,
               numbered by first encounter over (cells x cell_node_list x components)."""
    bs = V.bs
    node_point = plex.node_points(V)
    isbc = np.zeros(V.nnodes, dtype=bool)
    if bc_nodes is not None and len(bc_nodes):
        isbc[np.asarray(bc_nodes)] = True
    cS, cE = plex.getHeightStratum(0)
    offsets, out = [0], []
    for pts in point_sets:
        ht = set(int(p) for p in pts)
        cells = set()
        for p in ht:
            star, _ = plex.getTransitiveClosure(p, useCone=False)
            for q in star:
                if cS <= q < cE:
                    cells.add(int(q))
        seen = set()
        local = []
        for c in sorted(cells):
            for node in V.cell_nodes[c]:
                node = int(node)
                if node in seen:
                    continue
                if int(node_point[node]) in ht and not isbc[node]:
                    seen.add(node)
                    for comp in range(bs):
                        local.append(node * bs + comp)
        out.extend(local)
        offsets.append(len(out))
    return np.asarray(offsets, dtype=np.int64), np.asarray(out, dtype=np.int32)


def greedy_colouring(offsets, dofs, order, ndofs):
    """Patches in iteration order, lowest colour not used by a patch sharing a dof."""
    owner_colours = [set() for _ in range(ndofs)]
    colours = np.full(len(offsets) - 1, -1, dtype=np.int32)
    for p in order:
        if colours[p] >= 0:
            continue
        I = dofs[offsets[p]:offsets[p + 1]]
        used = set()
        for d in I:
            used |= owner_colours[d]
        c = 0
        while c in used:
            c += 1
        if len(I) == 0:
            c = 0
        colours[p] = c
        for d in I:
            owner_colours[d].add(c)
    colours[colours < 0] = 0
    return colours
