"""A minimal DMPlex look-alike over a SimplexMesh (host side).

The reference's patch constructors talk to a petsc4py ``DMPlex`` (alfi/relaxation.py:31-67,
110-150; alfi/transfer.py:18-45).  petsc4py is not available here, so this class offers the
handful of methods those callbacks use, with DMPlex's conventions:

* points are numbered cells, then vertices, then (3-D) faces, then edges — the interpolated
  DMPlex layout — so ``getDepthStratum(0)`` are vertices and ``getHeightStratum(0)`` cells;
* ``getTransitiveClosure(p, useCone)`` returns ``(points, orientations)``; the point order is
  ascending within each stratum walked from ``p`` outwards (DMPlex's own order is a BFS whose
  details nobody relies on: PCPATCH puts the points into a hash set);
* labels are plain integer arrays with -1 for "not labelled" (``getLabelValue``).

It also exposes the closure/star relations as sparse boolean matrices for the vectorised
patch builders in :mod:`alfi_b200.patches`.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .mesh import LOCAL_EDGES, SimplexMesh

__all__ = ["SynthPlex"]


class SynthPlex:
    def __init__(self, mesh: SimplexMesh):
        self.mesh = mesh
        d = self.dim = mesh.dim
        nc, nv, ne, nf = mesh.nc, mesh.nv, mesh.ne, mesh.nf
        self.cStart, self.cEnd = 0, nc
        self.vStart, self.vEnd = nc, nc + nv
        if d == 3:
            self.fStart, self.fEnd = nc + nv, nc + nv + nf
            self.eStart, self.eEnd = self.fEnd, self.fEnd + ne
        else:
            self.fStart = self.fEnd = nc + nv
            self.eStart, self.eEnd = nc + nv, nc + nv + ne
        self.npoints = self.eEnd
        self.labels: dict[str, np.ndarray] = {}
        if mesh.macro_vertex is not None:
            lab = np.full(self.npoints, -1, dtype=np.int64)
            lab[self.vStart:self.vEnd][mesh.macro_vertex] = 1
            self.labels["MacroVertices"] = lab
        # cone relation as CSR (points x points)
        rows, cols = [], []
        if d == 3:
            rows.append(np.repeat(np.arange(nc), 4))
            cols.append(self.fStart + mesh.cell_faces.ravel())
            # face -> edges: edges of face (a,b,c) are (a,b),(a,c),(b,c)
            f = mesh.faces
            fe = self._edge_ids(np.stack([f[:, [0, 1]], f[:, [0, 2]], f[:, [1, 2]]], axis=1).reshape(-1, 2))
            rows.append(np.repeat(self.fStart + np.arange(nf), 3))
            cols.append(self.eStart + fe)
        else:
            rows.append(np.repeat(np.arange(nc), 3))
            cols.append(self.eStart + mesh.cell_edges.ravel())
        rows.append(np.repeat(self.eStart + np.arange(ne), 2))
        cols.append(self.vStart + mesh.edges.ravel())
        rows = np.concatenate(rows)
        cols = np.concatenate(cols)
        n = self.npoints
        self.cone = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n, n))
        self.cone.sort_indices()
        self.support = self.cone.T.tocsr()
        self.support.sort_indices()
        eye = sp.identity(n, dtype=np.int32, format="csr")
        D = self.cone.astype(np.int32)
        C = eye + D
        for _ in range(d - 1):
            C = eye + D @ C
        C.data[:] = 1
        C.sort_indices()
        self.closure = C.tocsr()                 # closure[p] = points in the closure of p
        self.star = C.T.tocsr()                  # star[p]    = points whose closure contains p
        self.star.sort_indices()

    def _edge_ids(self, pairs):
        nv = np.int64(self.mesh.nv)
        keys = self.mesh.edges[:, 0] * nv + self.mesh.edges[:, 1]      # sorted by construction
        return np.searchsorted(keys, pairs[:, 0] * nv + pairs[:, 1])

    # ---- petsc4py.DMPlex protocol (the subset the reference uses) -------------------------
    def getDimension(self):
        return self.dim

    def getChart(self):
        return (0, self.npoints)

    def getDepthStratum(self, depth):
        d = self.dim
        if depth == 0:
            return (self.vStart, self.vEnd)
        if depth == 1:
            return (self.eStart, self.eEnd)
        if depth == d:
            return (self.cStart, self.cEnd)
        if depth == 2 and d == 3:
            return (self.fStart, self.fEnd)
        raise ValueError("no such stratum")

    def getHeightStratum(self, height):
        return self.getDepthStratum(self.dim - height)

    def getCone(self, p):
        return self.cone.indices[self.cone.indptr[p]:self.cone.indptr[p + 1]].copy()

    def getSupport(self, p):
        return self.support.indices[self.support.indptr[p]:self.support.indptr[p + 1]].copy()

    def getTransitiveClosure(self, p, useCone=True):
        rel = self.closure if useCone else self.star
        pts = rel.indices[rel.indptr[p]:rel.indptr[p + 1]]
        # p first, then outward by stratum distance (cells<verts<faces<edges numbering is not
        # monotone in dimension, so order explicitly)
        dist = np.abs(self.point_dim(pts) - self.point_dim(np.array([p]))[0])
        order = np.lexsort((pts, dist))
        pts = pts[order].astype(np.int32)
        return pts, np.zeros_like(pts)

    def point_dim(self, pts):
        pts = np.asarray(pts)
        out = np.empty(pts.shape, dtype=np.int64)
        out[(pts >= self.cStart) & (pts < self.cEnd)] = self.dim
        out[(pts >= self.vStart) & (pts < self.vEnd)] = 0
        out[(pts >= self.eStart) & (pts < self.eEnd)] = 1
        if self.dim == 3:
            out[(pts >= self.fStart) & (pts < self.fEnd)] = 2
        return out

    def getLabelValue(self, name, p):
        lab = self.labels.get(name)
        return -1 if lab is None else int(lab[p])

    def setLabelValue(self, name, p, value):
        lab = self.labels.setdefault(name, np.full(self.npoints, -1, dtype=np.int64))
        lab[p] = value

    def point_coords(self, p):
        """Mean of the vertex coordinates in the closure of p (alfi/relaxation.py:61-67)."""
        pts = self.closure.indices[self.closure.indptr[p]:self.closure.indptr[p + 1]]
        v = pts[(pts >= self.vStart) & (pts < self.vEnd)] - self.vStart
        return self.mesh.coords[v].mean(axis=0)

    # ---- node attachment (the PetscSection of a function space) ---------------------------
    def node_points(self, V):
        """point id each node of VectorSpace V is attached to → (nnodes,) int64."""
        m = self.mesh
        out = np.full(V.nnodes, -1, dtype=np.int64)
        out[V.vertex_nodes[:, 0]] = self.vStart + np.arange(m.nv)
        if V.edge_nodes.shape[1]:
            out[V.edge_nodes.ravel()] = np.repeat(self.eStart + np.arange(m.ne), V.edge_nodes.shape[1])
        if V.face_nodes.size:
            out[V.face_nodes.ravel()] = self.fStart + np.arange(m.nf)
        if V.cell_int_nodes.size:
            out[V.cell_int_nodes.ravel()] = np.arange(m.nc)
        assert (out >= 0).all()
        return out
