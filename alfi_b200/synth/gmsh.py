"""Gmsh MSH 2.2 (ASCII) reader / writer and a backward-facing-step mesh generator (host side).

The reference's bfs2d example (BASELINE.json configs[2]) loads ``coarse*.msh`` files written by
Gmsh from ``backwards-facing-step.geo`` with ``firedrake.Mesh`` (examples/bfs2d/bfs2d.py:14-17):
format 2.2, 2-node lines (element type 1) carrying the physical tags 1 = Inflow, 2 = NoSlip,
3 = Outflow and 3-node triangles (type 2).  Boundary conditions are set on those tags
(bfs2d.py:25-27).  `read_msh` turns such a file into a :class:`SimplexMesh` with tagged boundary
facets; `step_mesh` builds a mesh of the same domain, ``[0,10]x[0,2]`` minus ``[0,1]x[0,1]``, with
the same tags without needing Gmsh (randomised diagonals and jittered interior vertices, so the
vertex valences and hence the macro-star patch sizes vary as on an unstructured mesh).
"""
from __future__ import annotations

import numpy as np

from .mesh import SimplexMesh

__all__ = ["read_msh", "write_msh", "step_mesh", "INFLOW", "NOSLIP", "OUTFLOW"]

INFLOW, NOSLIP, OUTFLOW = 1, 2, 3          # physical tags of backwards-facing-step.geo
_NODES_PER_TYPE = {1: 2, 2: 3, 3: 4, 4: 4, 15: 1}


def read_msh(path: str) -> SimplexMesh:
    """MSH 2.2 ASCII → 2-D SimplexMesh (triangles; tagged lines become `boundary_facets`)."""
    with open(path) as fh:
        tok = fh.read().split("\n")
    i = 0
    nodes = None
    node_ids = None
    lines, line_tags, tris = [], [], []
    while i < len(tok):
        sec = tok[i].strip()
        if sec == "$MeshFormat":
            ver = tok[i + 1].split()
            if not ver[0].startswith("2") or ver[1] != "0":
                raise ValueError("only MSH 2.x ASCII is supported, got %r" % tok[i + 1])
            i += 3
        elif sec == "$Nodes":
            n = int(tok[i + 1])
            arr = np.array([ln.split() for ln in tok[i + 2:i + 2 + n]], dtype=np.float64)
            node_ids = arr[:, 0].astype(np.int64)
            nodes = arr[:, 1:4]
            i += n + 3
        elif sec == "$Elements":
            n = int(tok[i + 1])
            for ln in tok[i + 2:i + 2 + n]:
                f = ln.split()
                etype, ntags = int(f[1]), int(f[2])
                if etype not in _NODES_PER_TYPE:
                    raise ValueError("unsupported Gmsh element type %d" % etype)
                conn = [int(v) for v in f[3 + ntags:3 + ntags + _NODES_PER_TYPE[etype]]]
                if etype == 1:
                    lines.append(conn)
                    line_tags.append(int(f[3]) if ntags else 0)      # first tag = physical group
                elif etype == 2:
                    tris.append(conn)
                elif etype in (3, 4):
                    raise ValueError("only triangle meshes are supported")
            i += n + 3
        else:
            i += 1
    if nodes is None or not tris:
        raise ValueError("no $Nodes / triangles in %s" % path)
    if np.abs(nodes[:, 2]).max() > 0:
        raise ValueError("expected a planar mesh (z = 0)")
    remap = np.full(int(node_ids.max()) + 1, -1, dtype=np.int64)
    remap[node_ids] = np.arange(node_ids.size)
    cells = remap[np.asarray(tris, dtype=np.int64)]
    used = np.zeros(node_ids.size, dtype=bool)
    used[cells.ravel()] = True
    if not used.all():                       # drop nodes no triangle uses (geometry points)
        compact = np.cumsum(used) - 1
        cells = compact[cells]
        remap_used = np.where(used, compact, -1)
    else:
        remap_used = np.arange(node_ids.size)
    bf = remap_used[remap[np.asarray(lines, dtype=np.int64).reshape(-1, 2)]]
    bt = np.asarray(line_tags, dtype=np.int64)
    keep = (bf >= 0).all(axis=1) & (bt > 0)
    mesh = SimplexMesh(dim=2, coords=np.ascontiguousarray(nodes[used, :2]), cells=np.sort(cells, axis=1),
                       length=0.0, M=0, boundary_facets=np.sort(bf[keep], axis=1), boundary_tags=bt[keep])
    return mesh.build_topology()


def write_msh(path: str, mesh: SimplexMesh, names=((1, INFLOW, "Inflow"), (1, NOSLIP, "NoSlip"), (1, OUTFLOW, "Outflow"),
                                                   (2, 4, "Channel"))):
    """Write a 2-D SimplexMesh with tagged boundary facets as MSH 2.2 ASCII (what Gmsh would emit)."""
    if mesh.dim != 2:
        raise ValueError("2-D meshes only")
    out = ["$MeshFormat", "2.2 0 8", "$EndMeshFormat", "$PhysicalNames", str(len(names))]
    out += ['%d %d "%s"' % t for t in names]
    out += ["$EndPhysicalNames", "$Nodes", str(mesh.nv)]
    out += ["%d %.16g %.16g 0" % (i + 1, x, y) for i, (x, y) in enumerate(mesh.coords)]
    out += ["$EndNodes", "$Elements"]
    bf = mesh.boundary_facets if mesh.boundary_facets is not None else np.empty((0, 2), np.int64)
    out.append(str(bf.shape[0] + mesh.nc))
    k = 1
    for (a, b), t in zip(bf, mesh.boundary_tags if bf.size else []):
        out.append("%d 1 2 %d %d %d %d" % (k, t, t, a + 1, b + 1))
        k += 1
    for a, b, c in mesh.cells:
        out.append("%d 2 2 4 1 %d %d %d" % (k, a + 1, b + 1, c + 1))
        k += 1
    out += ["$EndElements", ""]
    with open(path, "w") as fh:
        fh.write("\n".join(out))


def step_mesh(n: int, seed: int | None = 0, jitter: float = 0.2) -> SimplexMesh:
    """Triangulation of the backward-facing step ``[0,10]x[0,2]`` minus ``[0,1]x[0,1]`` with ``n`` cells per
    unit length and the tags of backwards-facing-step.geo: Inflow = {x = 0, 1 <= y <= 2}, Outflow = {x = 10},
    NoSlip = everything else.  ``seed=None`` gives the plain structured mesh (all diagonals "left")."""
    h = 1.0 / n
    nx, ny = 10 * n, 2 * n
    ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    inside = ~((ii < n) & (jj < n))                    # grid vertices of the closed domain
    vid = np.full(inside.shape, -1, dtype=np.int64)
    vid[inside] = np.arange(inside.sum())
    coords = np.stack([ii[inside] * h, jj[inside] * h], axis=1).astype(np.float64)
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ok = ~((ci < n) & (cj < n))
    ci, cj = ci[ok], cj[ok]
    ll, lr, ul, ur = vid[cj, ci], vid[cj, ci + 1], vid[cj + 1, ci], vid[cj + 1, ci + 1]
    rng = np.random.default_rng(seed) if seed is not None else None
    flip = rng.random(ci.size) < 0.5 if rng is not None else np.zeros(ci.size, dtype=bool)
    t1 = np.where(flip[:, None], np.stack([ll, lr, ur], 1), np.stack([ll, lr, ul], 1))
    t2 = np.where(flip[:, None], np.stack([ll, ur, ul], 1), np.stack([lr, ur, ul], 1))
    cells = np.sort(np.stack([t1, t2], axis=1).reshape(-1, 3), axis=1)
    # boundary edges of the grid, tagged
    bf, bt = [], []

    def seg(va, vb, tag):
        bf.append(np.stack([va, vb], 1))
        bt.append(np.full(va.size, tag, dtype=np.int64))
    jr = np.arange(n, ny)
    seg(vid[jr, 0], vid[jr + 1, 0], INFLOW)                                  # x = 0, y in [1, 2]
    jr = np.arange(ny)
    seg(vid[jr, nx], vid[jr + 1, nx], OUTFLOW)                               # x = 10
    ir = np.arange(nx)
    seg(vid[ny, ir], vid[ny, ir + 1], NOSLIP)                                # top wall
    ir = np.arange(n, nx)
    seg(vid[0, ir], vid[0, ir + 1], NOSLIP)                                  # bottom wall behind the step
    ir = np.arange(n)
    seg(vid[n, ir], vid[n, ir + 1], NOSLIP)                                  # top of the step
    jr = np.arange(n)
    seg(vid[jr, n], vid[jr + 1, n], NOSLIP)                                  # face of the step
    if rng is not None and jitter > 0:
        on_bdry = np.zeros(coords.shape[0], dtype=bool)
        on_bdry[np.concatenate(bf).ravel()] = True
        coords[~on_bdry] += (rng.random((int((~on_bdry).sum()), 2)) - 0.5) * (jitter * h)
    mesh = SimplexMesh(dim=2, coords=coords, cells=cells, length=0.0, M=0,
                       boundary_facets=np.sort(np.concatenate(bf), axis=1), boundary_tags=np.concatenate(bt))
    return mesh.build_topology()
