"""Synthetic simplicial meshes standing in for Firedrake/DMPlex (host side, numpy).

The reference builds its meshes with Firedrake (`RectangleMesh(..., diagonal="left")`,
examples/ldc2d/ldc2d.py:17-21; `BoxMesh`, examples/ldc3d/ldc3d.py:13-16), refines them
uniformly and Alfeld-splits every level (alfi/bary.py:16-27, 29-194).  None of that stack is
available here, so this module generates the same family of meshes directly:

* Kuhn (Freudenthal) triangulations of ``[0, L]^d`` with ``M`` cells per side.  The Kuhn mesh
  with ``2M`` cells per side is the red refinement of the one with ``M`` (Bey), so the uniform
  hierarchy is ``M_l = N * 2**l`` and is nested.
* Alfeld (barycentric) split: macro cell ``c`` becomes cells ``c*(d+1) .. c*(d+1)+d`` — the
  numbering alfi/bary.py:148-157 relies on — macro vertices keep their ids (label
  ``MacroVertices`` = 1, alfi/bary.py:18-19) and the barycentre of macro cell ``c`` is vertex
  ``nv_macro + c``.

Cells always list their vertices in ascending global id, so local edge/face orientations agree
between neighbouring cells and no orientation fix-up is needed for P3 edge nodes.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np

__all__ = ["SimplexMesh", "kuhn_mesh", "alfeld_split", "refine_uniform", "LOCAL_EDGES", "LOCAL_FACES"]

# local sub-entity -> local vertices, lexicographic
LOCAL_EDGES = {2: [(0, 1), (0, 2), (1, 2)],
               3: [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]}
LOCAL_FACES = {2: [], 3: [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]}


def _unique_rows(keys: np.ndarray):
    """Unique of 1-D int64 keys → (unique keys, inverse)."""
    uniq, inv = np.unique(keys, return_inverse=True)
    return uniq, inv.astype(np.int64)


@dataclass
class SimplexMesh:
    dim: int
    coords: np.ndarray            # (nv, dim) float64
    cells: np.ndarray             # (nc, dim+1) int64, each row ascending
    macro_vertex: np.ndarray | None = None   # (nv,) bool, Alfeld meshes only
    macro: "SimplexMesh | None" = None       # the mesh this one is the Alfeld split of
    length: float = 2.0
    M: int = 0                               # cells per side of the underlying Kuhn grid
    # box-shaped Kuhn grids (weak-scaling family): axis a has M * shape[a] cells and extent length * shape[a];
    # () = the cube
    shape: tuple = ()
    # sub-boxes of a Kuhn grid (rank-local generation, synth/bricks.py): explicit cells per axis and the lower corner
    # in cells of this grid; () = derived from M and shape, origin 0
    counts: tuple = ()
    origin: tuple = ()
    # boundary markers of general (Gmsh) meshes: (nb, dim) vertex tuples of tagged boundary facets, rows
    # ascending, and their physical tags; `facet_tag` (per facet id, 0 = untagged) is derived from them
    boundary_facets: np.ndarray | None = None
    boundary_tags: np.ndarray | None = None
    facet_tag: np.ndarray = field(default=None, repr=False)
    # derived topology
    edges: np.ndarray = field(default=None, repr=False)       # (ne, 2)
    faces: np.ndarray = field(default=None, repr=False)       # (nf, 3) (3-D only)
    cell_edges: np.ndarray = field(default=None, repr=False)  # (nc, n_local_edges)
    cell_faces: np.ndarray = field(default=None, repr=False)  # (nc, 4) (3-D only)

    @property
    def nv(self):
        return self.coords.shape[0]

    @property
    def nc(self):
        return self.cells.shape[0]

    @property
    def ne(self):
        return self.edges.shape[0]

    @property
    def nf(self):
        return 0 if self.faces is None else self.faces.shape[0]

    def build_topology(self):
        d, nv = self.dim, np.int64(self.nv)
        c = self.cells
        le = LOCAL_EDGES[d]
        ek = np.stack([c[:, a] * nv + c[:, b] for a, b in le], axis=1)
        uniq, inv = _unique_rows(ek.ravel())
        self.edges = np.stack([uniq // nv, uniq % nv], axis=1)
        self.cell_edges = inv.reshape(c.shape[0], len(le))
        if d == 3:
            lf = LOCAL_FACES[3]
            fk = np.stack([(c[:, a] * nv + c[:, b]) * nv + c[:, e] for a, b, e in lf], axis=1)
            uniq, inv = _unique_rows(fk.ravel())
            self.faces = np.stack([uniq // (nv * nv), (uniq // nv) % nv, uniq % nv], axis=1)
            self.cell_faces = inv.reshape(c.shape[0], 4)
        if self.boundary_facets is not None:
            self.facet_tag = self._tags_of_facets()
        return self

    def _facet_keys(self, f):
        nv = np.int64(self.nv)
        key = f[:, 0].astype(np.int64)
        for j in range(1, f.shape[1]):
            key = key * nv + f[:, j]
        return key

    def _tags_of_facets(self):
        """Physical tag of every facet (0 where none): match the tagged vertex tuples to facet ids."""
        fac = self.facets
        tag = np.zeros(fac.shape[0], dtype=np.int64)
        if self.boundary_facets.size:
            keys = self._facet_keys(fac)                     # ascending: facets come out of np.unique
            want = self._facet_keys(np.sort(self.boundary_facets, axis=1))
            pos = np.searchsorted(keys, want)
            ok = (pos < keys.size) & (keys[np.minimum(pos, keys.size - 1)] == want)
            if not ok.all():
                raise ValueError("a tagged boundary facet is not a facet of the mesh")
            tag[pos] = self.boundary_tags
        return tag

    # ---- facets (codim 1): edges in 2-D, faces in 3-D ----
    @property
    def facets(self):
        return self.edges if self.dim == 2 else self.faces

    @property
    def cell_facets(self):
        return self.cell_edges if self.dim == 2 else self.cell_faces

    @property
    def axis_shape(self):
        return tuple(self.shape) if self.shape else (1,) * self.dim

    @property
    def axis_counts(self):
        """Cells per axis of the underlying Kuhn grid."""
        return tuple(self.counts) if self.counts else tuple(self.M * v for v in self.axis_shape)

    @property
    def lower(self):
        """Lower corner of the Kuhn box (general meshes: of the bounding box)."""
        if not self.M:
            return self.coords.min(axis=0)
        h = self.length / self.M
        return h * np.asarray(self.origin if self.origin else (0,) * self.dim, dtype=np.float64)

    @property
    def extent(self):
        """Upper corner of the Kuhn box (lower corner: `lower`, the origin unless this is a sub-box)."""
        if not self.M:
            return self.coords.max(axis=0)
        return self.lower + (self.length / self.M) * np.asarray(self.axis_counts, dtype=np.float64)

    def boundary_vertex_mask(self, tol=1e-12):
        x = self.coords
        return np.any((np.abs(x) < tol) | (np.abs(x - self.extent[None, :]) < tol), axis=1)


def kuhn_mesh(dim: int, M: int, length: float = 2.0, shape: tuple = (), counts: tuple = (), origin: tuple = ()) -> SimplexMesh:
    """Kuhn triangulation of [0, length]^dim with M cells per side — or, with `shape`, of the box
    [0, length * shape[a]] with M * shape[a] cells along axis a (same cell size; the weak-scaling family).

    2-D: each square (ll, lr, ur, ul) is cut along lr–ul (Firedrake ``diagonal="left"``,
    examples/ldc2d/ldc2d.py:11-12).  3-D: six tetrahedra per cube sharing the main diagonal
    (Firedrake ``BoxMesh``, examples/ldc3d/ldc3d.py:13-15).
    """
    shp = tuple(int(v) for v in shape) if shape else (1,) * dim
    if len(shp) != dim or min(shp) < 1:
        raise ValueError("shape must have one positive integer per axis")
    # `counts` / `origin` (in cells of size length / M): an arbitrary sub-box of the grid, same cut directions
    Ma = [int(v) for v in counts] if counts else [M * v for v in shp]
    org = np.asarray([int(v) for v in origin] if origin else [0] * dim, dtype=np.float64)
    if len(Ma) != dim or min(Ma) < 1:
        raise ValueError("counts must have one positive integer per axis")
    n1 = [m + 1 for m in Ma]
    h = length / M
    if dim == 2:
        # vertex id = i + n1x*j  (x fastest)
        J, I = np.meshgrid(np.arange(n1[1]), np.arange(n1[0]), indexing="ij")
        coords = (np.stack([I.ravel(), J.ravel()], axis=1) + org[None, :]) * h
        j, i = np.meshgrid(np.arange(Ma[1]), np.arange(Ma[0]), indexing="ij")
        i, j = i.ravel(), j.ravel()
        ll = i + n1[0] * j
        lr = ll + 1
        ul = ll + n1[0]
        ur = ul + 1
        cells = np.stack([np.stack([ll, lr, ul], 1), np.stack([lr, ur, ul], 1)], axis=1).reshape(-1, 3)
    elif dim == 3:
        K, J, I = np.meshgrid(np.arange(n1[2]), np.arange(n1[1]), np.arange(n1[0]), indexing="ij")
        coords = (np.stack([I.ravel(), J.ravel(), K.ravel()], axis=1) + org[None, :]) * h
        k, j, i = np.meshgrid(np.arange(Ma[2]), np.arange(Ma[1]), np.arange(Ma[0]), indexing="ij")
        base = (i + n1[0] * (j + n1[1] * k)).ravel()
        step = np.array([1, n1[0], n1[0] * n1[1]])
        tets = []
        for perm in itertools.permutations(range(3)):
            v0 = base
            v1 = v0 + step[perm[0]]
            v2 = v1 + step[perm[1]]
            v3 = v2 + step[perm[2]]
            tets.append(np.stack([v0, v1, v2, v3], 1))
        cells = np.stack(tets, axis=1).reshape(-1, 4)
    else:
        raise ValueError("dim must be 2 or 3")
    cells = np.sort(cells.astype(np.int64), axis=1)
    m = SimplexMesh(dim=dim, coords=coords.astype(np.float64), cells=cells, length=length, M=M,
                    shape=shp if any(v != 1 for v in shp) else (), counts=tuple(Ma) if counts else (),
                    origin=tuple(int(v) for v in origin) if origin else ())
    return m.build_topology()


def alfeld_split(macro: SimplexMesh) -> SimplexMesh:
    """Barycentric refinement with the numbering conventions of alfi/bary.py:16-27,148-157."""
    d = macro.dim
    nvm, ncm = macro.nv, macro.nc
    bary = macro.coords[macro.cells].mean(axis=1)
    coords = np.concatenate([macro.coords, bary], axis=0)
    bid = nvm + np.arange(ncm, dtype=np.int64)
    sub = []
    for r in range(d + 1):
        cc = macro.cells.copy()
        cc[:, r] = bid             # replace local vertex r by the barycentre
        sub.append(cc)
    cells = np.sort(np.stack(sub, axis=1).reshape(-1, d + 1), axis=1)
    mv = np.zeros(nvm + ncm, dtype=bool)
    mv[:nvm] = True
    # macro vertices keep their ids, so tagged boundary facets of the macro mesh are facets of the split
    m = SimplexMesh(dim=d, coords=coords, cells=cells, macro_vertex=mv, macro=macro,
                    length=macro.length, M=macro.M, shape=macro.shape, counts=macro.counts, origin=macro.origin,
                    boundary_facets=macro.boundary_facets,
                    boundary_tags=macro.boundary_tags)
    return m.build_topology()


def refine_uniform(mesh: SimplexMesh):
    """Red refinement of a general triangle mesh (Firedrake ``MeshHierarchy`` / DMPlex uniform refinement,
    alfi/problem.py:10-24): a new vertex on every edge (id ``nv + edge id``), four children per cell.

    Returns ``(fine, c2f)`` with ``c2f[c] = [4c, 4c+1, 4c+2, 4c+3]`` (three corner children in the
    parent's vertex order, then the middle one).  Tagged boundary facets are split with their tag."""
    if mesh.dim != 2:
        raise NotImplementedError("general uniform refinement is implemented for triangles (Kuhn hierarchies "
                                  "cover the 3-D configurations)")
    nv = mesh.nv
    c, ce = mesh.cells, mesh.cell_edges                      # local edges (0,1), (0,2), (1,2)
    m01, m02, m12 = nv + ce[:, 0], nv + ce[:, 1], nv + ce[:, 2]
    kids = np.stack([np.stack([c[:, 0], m01, m02], 1), np.stack([c[:, 1], m01, m12], 1),
                     np.stack([c[:, 2], m02, m12], 1), np.stack([m01, m02, m12], 1)], axis=1).reshape(-1, 3)
    coords = np.concatenate([mesh.coords, mesh.coords[mesh.edges].mean(axis=1)], axis=0)
    bf = bt = None
    if mesh.boundary_facets is not None:
        tagged = np.flatnonzero(mesh.facet_tag)
        e = mesh.edges[tagged]
        mid = nv + tagged
        bf = np.sort(np.concatenate([np.stack([e[:, 0], mid], 1), np.stack([e[:, 1], mid], 1)], axis=0), axis=1)
        bt = np.concatenate([mesh.facet_tag[tagged], mesh.facet_tag[tagged]])
    fine = SimplexMesh(dim=2, coords=coords, cells=np.sort(kids.astype(np.int64), axis=1), length=mesh.length,
                       M=0, boundary_facets=bf, boundary_tags=bt)
    return fine.build_topology(), np.arange(4 * mesh.nc, dtype=np.int64).reshape(mesh.nc, 4)


def locate_in_kuhn(mesh: SimplexMesh, pts: np.ndarray) -> np.ndarray:
    """Index of the Kuhn cell of ``mesh`` (a plain Kuhn mesh) containing each point.

    Used to build ``coarse_to_fine_cells`` of the uniform hierarchy by locating fine-cell
    centroids (strictly interior, so there are no ties).
    """
    d, h = mesh.dim, mesh.length / mesh.M
    Ma = np.asarray(mesh.axis_counts, dtype=np.int64)
    g = (pts - mesh.lower[None, :]) / h
    ijk = np.clip(np.floor(g).astype(np.int64), 0, Ma[None, :] - 1)
    frac = g - ijk
    if d == 2:
        cube = ijk[:, 0] + Ma[0] * ijk[:, 1]
        # cell 0 = (ll, lr, ul): x + y <= 1 ; cell 1 = (lr, ur, ul)
        which = (frac.sum(axis=1) > 1.0).astype(np.int64)
        return cube * 2 + which
    cube = ijk[:, 0] + Ma[0] * (ijk[:, 1] + Ma[1] * ijk[:, 2])
    # tet for permutation perm: frac[perm0] >= frac[perm1] >= frac[perm2]
    order = np.argsort(-frac, axis=1, kind="stable")
    perms = list(itertools.permutations(range(3)))
    lut = {p: n for n, p in enumerate(perms)}
    which = np.array([lut[tuple(o)] for o in order.tolist()], dtype=np.int64)
    return cube * 6 + which
