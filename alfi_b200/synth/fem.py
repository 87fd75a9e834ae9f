"""Vector Lagrange finite elements on the synthetic meshes + assembly of the velocity block.

Stands in for UFL/TSFC/PyOP2 assembly (host side; in a deployment Firedrake does this and hands
the BAIJ values over once per Newton step).  The operator is the (1,1) block of the Newton
linearisation of the reference's residuals:

* Scott–Vogelius (alfi/solver.py:613-623):
  ``nu*(2 sym grad u, grad v) + gamma*(div u, div v) + advect*((w.grad)u + (u.grad)w, v)``
* [Pk]^d–P0 (alfi/solver.py:562-572): same with ``gamma*(cell_avg(div u), div v)``.

and the transfer forms of alfi/transfer.py:295-309 (SV) / 319-332 (PkP0).

Everything is reduced to reference tensors contracted with the affine cell geometry, so the
assembly is a handful of dense matmuls per chunk of cells.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from functools import lru_cache

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi

from .mesh import LOCAL_EDGES, LOCAL_FACES, SimplexMesh

__all__ = ["LagrangeElement", "P1FBElement", "make_element", "VectorSpace", "assemble_velocity_block", "BSR",
           "reference_tensors", "FacetBlockPattern", "facet_adjacency", "burman_facet_tensors"]


# --------------------------------------------------------------------------- reference element
def _lattice(dim: int, k: int):
    """Barycentric coordinates of the Pk nodes, entity by entity.

    Returns (bary (n, dim+1), entity list [(edim, local entity index, #nodes)]) in the order
    vertices, edges (low→high vertex), faces, cell interior.  Supports k <= 3.
    """
    if not 1 <= k <= 3:
        raise NotImplementedError("Lagrange degree 1..3 only")
    nvl = dim + 1
    pts, ents = [], []
    for v in range(nvl):
        lam = np.zeros(nvl)
        lam[v] = 1.0
        pts.append(lam)
        ents.append((0, v, 1))
    for e, (a, b) in enumerate(LOCAL_EDGES[dim]):
        for j in range(1, k):
            lam = np.zeros(nvl)
            lam[a] = (k - j) / k
            lam[b] = j / k
            pts.append(lam)
        ents.append((1, e, k - 1))
    if k == 3:
        if dim == 2:
            pts.append(np.full(3, 1.0 / 3.0))
            ents.append((2, 0, 1))
        else:
            for f, (a, b, c) in enumerate(LOCAL_FACES[3]):
                lam = np.zeros(4)
                lam[[a, b, c]] = 1.0 / 3.0
                pts.append(lam)
                ents.append((2, f, 1))
    return np.array(pts), ents


def _monomials(dim: int, k: int):
    return [e for e in itertools.product(range(k + 1), repeat=dim) if sum(e) <= k]


def _eval_monomials(expo, x, deriv=None):
    """x: (..., dim).  deriv=None → values (..., nmono); deriv=a → d/dx_a."""
    out = np.ones(x.shape[:-1] + (len(expo),))
    for m, e in enumerate(expo):
        v = np.ones(x.shape[:-1])
        for a, p in enumerate(e):
            if deriv == a:
                v = v * (p * x[..., a] ** (p - 1) if p > 0 else 0.0)
            else:
                v = v * x[..., a] ** p
        out[..., m] = v
    return out


def simplex_quadrature(dim: int, degree: int):
    """Collapsed Gauss–Jacobi rule on the reference simplex, exact to ``degree``."""
    n = degree // 2 + 1
    x0, w0 = roots_jacobi(n, 0, 0)
    x0, w0 = (x0 + 1) / 2, w0 / 2
    x1, w1 = roots_jacobi(n, 1, 0)
    x1, w1 = (x1 + 1) / 2, w1 / 4
    if dim == 2:
        X1, X0 = np.meshgrid(x1, x0, indexing="ij")
        W = np.outer(w1, w0)
        pts = np.stack([X1.ravel(), (X0 * (1 - X1)).ravel()], axis=1)
        return pts, W.ravel()
    x2, w2 = roots_jacobi(n, 2, 0)
    x2, w2 = (x2 + 1) / 2, w2 / 8
    X2, X1, X0 = np.meshgrid(x2, x1, x0, indexing="ij")
    W = w2[:, None, None] * w1[None, :, None] * w0[None, None, :]
    a = X2
    b = X1 * (1 - X2)
    c = X0 * (1 - X1) * (1 - X2)
    return np.stack([a.ravel(), b.ravel(), c.ravel()], axis=1), W.ravel()


@dataclass(frozen=True)
class LagrangeElement:
    dim: int
    degree: int

    @property
    def nodes_bary(self):
        return _lattice(self.dim, self.degree)[0]

    @property
    def entities(self):
        return _lattice(self.dim, self.degree)[1]

    @property
    def nnodes(self):
        return self.nodes_bary.shape[0]

    @property
    def nodes_ref(self):
        return self.nodes_bary[:, 1:]      # reference coords: vertex 0 at origin, vertex i at e_i

    def _coeffs(self):
        return _coeffs(self.dim, self.degree)

    def tabulate(self, x):
        """Basis values at reference points x (npts, dim) → (npts, nnodes)."""
        expo = _monomials(self.dim, self.degree)
        return _eval_monomials(expo, x) @ self._coeffs()

    def tabulate_grad(self, x):
        """Reference gradients → (npts, nnodes, dim)."""
        expo = _monomials(self.dim, self.degree)
        C = self._coeffs()
        return np.stack([_eval_monomials(expo, x, deriv=a) @ C for a in range(self.dim)], axis=-1)


@lru_cache(maxsize=None)
def _coeffs(dim, k):
    nodes = _lattice(dim, k)[0][:, 1:]
    V = _eval_monomials(_monomials(dim, k), nodes)
    return np.linalg.inv(V)             # column i = monomial coefficients of basis function i


@dataclass(frozen=True)
class P1FBElement:
    """P1 enriched with facet bubbles on a tetrahedron, nodal basis at the 4 vertices and the 4
    face centroids: Firedrake's ``NodalEnrichedElement(P1, FacetBubble)`` of the [P1+FB]^3-P0
    scheme (alfi/solver.py:574-584).  With b_f = 27 lambda_a lambda_b lambda_c the bubble of face f
    (opposite vertex f):  phi_vertex_i = lambda_i - (1/3) sum_{f != i} b_f,  phi_face_f = b_f —
    the change of basis hard-wired in alfi/bubble.py:58-147.  Every basis function is a cubic, so
    it is tabulated through the P3 Lagrange basis."""
    dim: int = 3
    degree: int = 3                      # polynomial degree (for quadrature)

    @property
    def nnodes(self):
        return 8

    @property
    def nodes_bary(self):
        pts = [np.eye(4)[i] for i in range(4)]
        for f in range(4):                                   # face f is opposite vertex f
            lam = np.full(4, 1.0 / 3.0)
            lam[f] = 0.0
            pts.append(lam)
        return np.array(pts)

    @property
    def nodes_ref(self):
        return self.nodes_bary[:, 1:]

    @property
    def entities(self):
        return [(0, v, 1) for v in range(4)] + [(2, f, 1) for f in range(4)]

    def _to_p3(self):
        """C[i, j] = phi_i(x_j) at the 20 P3 lattice points x_j."""
        lam = _lattice(3, 3)[0]                              # (20, 4) barycentric
        bub = np.stack([27.0 * np.prod(np.delete(lam, f, axis=1), axis=1) for f in range(4)], axis=0)
        C = np.zeros((8, lam.shape[0]))
        for i in range(4):
            C[i] = lam[:, i] - sum(bub[f] for f in range(4) if f != i) / 3.0
            C[4 + i] = bub[i]
        return C

    def tabulate(self, x):
        return LagrangeElement(3, 3).tabulate(x) @ self._to_p3().T

    def tabulate_grad(self, x):
        g = LagrangeElement(3, 3).tabulate_grad(x)           # (q, 20, 3)
        return np.einsum("qja,ij->qia", g, self._to_p3())


def make_element(dim: int, k: int, kind: str = "lagrange"):
    if kind == "lagrange":
        return LagrangeElement(dim, k)
    if kind == "p1fb":
        if dim != 3:
            raise NotImplementedError("P1+FacetBubble is the 3-D element of alfi (solver.py:576-579)")
        return P1FBElement()
    raise ValueError(kind)


@lru_cache(maxsize=None)
def reference_tensors(el):
    """Exact reference integrals used by the assembly (see module docstring).

    K[a,b,i,j] = ∫ d_a phi_i d_b phi_j          T1[k,a,i,j] = ∫ phi_k phi_i d_a phi_j
    T2[k,a,i,j] = ∫ d_a phi_k phi_i phi_j       Dv[a,i] = ∫ d_a phi_i       M[i,j] = ∫ phi_i phi_j
    """
    dim = el.dim
    x, w = simplex_quadrature(dim, 3 * el.degree)
    phi = el.tabulate(x)                # (q, n)
    dphi = el.tabulate_grad(x)          # (q, n, dim)
    K = np.einsum("q,qia,qjb->abij", w, dphi, dphi)
    T1 = np.einsum("q,qk,qi,qja->kaij", w, phi, phi, dphi)
    T2 = np.einsum("q,qka,qi,qj->kaij", w, dphi, phi, phi)
    Dv = np.einsum("q,qia->ai", w, dphi)
    Mm = np.einsum("q,qi,qj->ij", w, phi, phi)
    return dict(K=K, T1=T1, T2=T2, Dv=Dv, M=Mm)


# --------------------------------------------------------------------------- function space
@dataclass
class VectorSpace:
    """[Pk]^d on a SimplexMesh; node numbering = first encounter walking cells in order."""
    mesh: SimplexMesh
    degree: int
    kind: str = "lagrange"               # "lagrange" | "p1fb"
    element: object = field(init=False)
    cell_nodes: np.ndarray = field(init=False, repr=False)      # (nc, nnodes_local)
    nnodes: int = field(init=False)
    node_coords: np.ndarray = field(init=False, repr=False)
    # node ids attached to each mesh entity, -1 padded: vertex (nv,1), edge (ne,k-1), face/cell
    vertex_nodes: np.ndarray = field(init=False, repr=False)
    edge_nodes: np.ndarray = field(init=False, repr=False)
    face_nodes: np.ndarray = field(init=False, repr=False)      # 3-D faces (k==3) else empty
    cell_int_nodes: np.ndarray = field(init=False, repr=False)  # 2-D k==3 else empty

    def __post_init__(self):
        m, k = self.mesh, self.degree
        d = m.dim
        self.element = el = make_element(d, k, self.kind)
        nv, ne, nf, nc = m.nv, m.ne, m.nf, m.nc
        cols = [m.cells]                                    # provisional ids, entity blocks
        off = nv
        p1fb = self.kind == "p1fb"
        per_edge = 0 if p1fb else k - 1
        if p1fb:
            # local face f of the element is opposite vertex f; mesh.cell_faces lists faces in
            # lexicographic vertex order (0,1,2),(0,1,3),(0,2,3),(1,2,3) = opposite 3,2,1,0
            cols.append(off + m.cell_faces[:, ::-1])
            off += nf
            k = 1                                           # no further Lagrange entities
        if per_edge:
            ce = m.cell_edges
            cols.append((off + ce[:, :, None] * per_edge + np.arange(per_edge)[None, None, :]).reshape(nc, -1))
            off += ne * per_edge
        if k == 3:
            if d == 3:
                cols.append(off + m.cell_faces)
                off += nf
            else:
                cols.append((off + np.arange(nc))[:, None])
                off += nc
        prov = np.concatenate(cols, axis=1).astype(np.int64)
        uniq, first = np.unique(prov.ravel(), return_index=True)
        order = np.argsort(first, kind="stable")
        new = np.empty(off, dtype=np.int64)
        new[uniq[order]] = np.arange(uniq.size)
        assert uniq.size == off
        self.cell_nodes = new[prov]
        self.nnodes = int(off)
        self.vertex_nodes = new[:nv].reshape(nv, 1)
        o = nv
        self.edge_nodes = new[o:o + ne * per_edge].reshape(ne, per_edge)
        o += ne * per_edge
        if p1fb:
            self.face_nodes = new[o:o + nf].reshape(nf, 1)
            self.cell_int_nodes = np.empty((nc, 0), dtype=np.int64)
        elif k == 3 and d == 3:
            self.face_nodes = new[o:o + nf].reshape(nf, 1)
            self.cell_int_nodes = np.empty((nc, 0), dtype=np.int64)
        elif k == 3 and d == 2:
            self.face_nodes = np.empty((0, 0), dtype=np.int64)
            self.cell_int_nodes = new[o:o + nc].reshape(nc, 1)
        else:
            self.face_nodes = np.empty((nf, 0), dtype=np.int64)
            self.cell_int_nodes = np.empty((nc, 0), dtype=np.int64)
        # node coordinates
        xc = np.einsum("nl,cld->cnd", el.nodes_bary, m.coords[m.cells])
        self.node_coords = np.empty((self.nnodes, d))
        self.node_coords[self.cell_nodes.ravel()] = xc.reshape(-1, d)

    @property
    def bs(self):
        return self.mesh.dim

    @property
    def ndofs(self):
        return self.nnodes * self.bs

    def boundary_nodes(self, tol=1e-12):
        x, L = self.node_coords, self.mesh.extent
        return np.flatnonzero(np.any((np.abs(x) < tol) | (np.abs(x - L[None, :]) < tol), axis=1))

    def tagged_boundary_nodes(self, tags):
        """Nodes in the closure of the boundary facets carrying one of the physical `tags` — what
        `DirichletBC(V, g, tag).nodes` is for a Gmsh mesh (examples/bfs2d/bfs2d.py:25-27).  2-D."""
        m = self.mesh
        if m.dim != 2 or m.facet_tag is None:
            raise NotImplementedError("tagged boundaries are implemented for 2-D meshes with boundary markers")
        e = np.flatnonzero(np.isin(m.facet_tag, np.asarray(tags)))
        nodes = [self.vertex_nodes[m.edges[e].ravel()].ravel(), self.edge_nodes[e].ravel()]
        return np.unique(np.concatenate(nodes))

    def interpolate(self, fn):
        """Nodal interpolant of fn(x)->(n, d); returns (nnodes, d)."""
        return np.asarray(fn(self.node_coords), dtype=np.float64)


# --------------------------------------------------------------------------- BSR container
@dataclass
class BSR:
    """Block CSR, square blocks stored row-major: vals[(k, r, c)]."""
    nbrows: int
    bs: int
    rowptr: np.ndarray      # int32 (nbrows+1)
    colidx: np.ndarray      # int32 (nnzb), ascending within a row
    vals: np.ndarray        # float64 (nnzb, bs, bs)

    @property
    def nnzb(self):
        return self.colidx.size

    def to_scipy(self):
        return sp.bsr_matrix((self.vals, self.colidx, self.rowptr),
                             shape=(self.nbrows * self.bs, self.nbrows * self.bs))

    def to_csr(self):
        A = self.to_scipy().tocsr()
        A.sort_indices()
        return A

    def copy(self):
        return BSR(self.nbrows, self.bs, self.rowptr, self.colidx, self.vals.copy())


class BlockPattern:
    """Node–node sparsity of a space + the scatter map from (cell, i, j) to a block slot."""

    def __init__(self, V: VectorSpace, chunk: int = 1 << 15):
        cn = V.cell_nodes
        nn = np.int64(V.nnodes)
        nl = cn.shape[1]
        keys = (cn[:, :, None] * nn + cn[:, None, :]).reshape(-1)
        self.perm = np.argsort(keys, kind="stable")
        ks = keys[self.perm]
        start = np.flatnonzero(np.concatenate(([True], ks[1:] != ks[:-1])))
        self.start = start
        uk = ks[start]
        rows = (uk // nn).astype(np.int64)
        self.colidx = (uk % nn).astype(np.int32)
        self.rowptr = np.zeros(V.nnodes + 1, dtype=np.int32)
        np.add.at(self.rowptr, rows + 1, 1)
        self.rowptr = np.cumsum(self.rowptr).astype(np.int32)
        self.rows = rows
        self.nl = nl
        self.nnzb = uk.size

    def scatter(self, elem: np.ndarray) -> np.ndarray:
        """elem: (nc, nl, nl) scalar contributions → (nnzb,) summed per block slot."""
        v = elem.reshape(-1)[self.perm]
        return np.add.reduceat(v, self.start)


class FacetBlockPattern(BlockPattern):
    """Sparsity of a form with interior-facet (dS) integrals — Burman's jump stabilisation, alfi/stabilisation.py:
    156-162 — on top of the cell integrals: a node couples to the nodes of its cells AND of their facet neighbours.
    `scatter` takes cell tensors, `scatter_facets` the (2 nl) x (2 nl) macro-element tensors of the interior facets."""

    def __init__(self, V: VectorSpace):
        cn = V.cell_nodes
        nn = np.int64(V.nnodes)
        nl = cn.shape[1]
        self.facets, self.facet_cells, self.facet_local = facet_adjacency(V.mesh)
        fn = np.concatenate([cn[self.facet_cells[:, 0]], cn[self.facet_cells[:, 1]]], axis=1)     # (nF, 2 nl)
        self.facet_nodes = fn
        ckeys = (cn[:, :, None] * nn + cn[:, None, :]).reshape(-1)
        fkeys = (fn[:, :, None] * nn + fn[:, None, :]).reshape(-1)
        uk = np.unique(np.concatenate([ckeys, fkeys]))
        self.cell_slot = np.searchsorted(uk, ckeys)
        self.facet_slot = np.searchsorted(uk, fkeys)
        rows = (uk // nn).astype(np.int64)
        self.colidx = (uk % nn).astype(np.int32)
        self.rowptr = np.zeros(V.nnodes + 1, dtype=np.int32)
        np.add.at(self.rowptr, rows + 1, 1)
        self.rowptr = np.cumsum(self.rowptr).astype(np.int32)
        self.rows = rows
        self.nl = nl
        self.nnzb = uk.size

    def scatter(self, elem: np.ndarray) -> np.ndarray:
        return np.bincount(self.cell_slot, weights=elem.reshape(-1), minlength=self.nnzb)

    def scatter_facets(self, elem: np.ndarray) -> np.ndarray:
        return np.bincount(self.facet_slot, weights=elem.reshape(-1), minlength=self.nnzb)


def facet_adjacency(mesh: SimplexMesh):
    """Interior facets: (facet ids (nF,), their two cells (nF, 2), the local facet index in each (nF, 2))."""
    cf = mesh.cell_facets
    nlf = cf.shape[1]
    f = cf.ravel()
    order = np.argsort(f, kind="stable")
    fs = f[order]
    first = np.flatnonzero(fs[1:] == fs[:-1])
    c, j = order // nlf, order % nlf
    return fs[first], np.stack([c[first], c[first + 1]], axis=1), np.stack([j[first], j[first + 1]], axis=1)


def _facet_quadrature(dim: int, degree: int):
    """Barycentric points (q, dim) on a facet of a dim-simplex and weights that sum to one."""
    if dim == 2:
        n = degree // 2 + 1
        x, w = roots_jacobi(n, 0, 0)
        x, w = (x + 1) / 2, w / 2
        return np.stack([1 - x, x], axis=1), w
    pts, w = simplex_quadrature(2, degree)
    return np.concatenate([1 - pts.sum(axis=1, keepdims=True), pts], axis=1), w / w.sum()


def burman_facet_tensors(V: VectorSpace, wind, weight: float):
    """Macro-element tensors of Burman's stabilisation (alfi/stabilisation.py:139-162),

        0.5 * weight * avg(h)^2 * beta * dot(jump(grad(u), n), jump(grad(v), n)) * dS,
        h = FacetArea (2-D) / FacetArea^0.5 (3-D),   beta = avg(facet_avg(sqrt(inner(wind, wind) + 1e-10))),

    for every interior facet: S[f, a*nl + i, b*nl + j] couples node i of side a with node j of side b; the form is the
    same scalar matrix for every velocity component (jump(grad(u), n) = the jump of the normal derivative, component
    by component).  Returns (facet ids, cells (nF, 2), S (nF, 2 nl, 2 nl))."""
    mesh, d, el = V.mesh, V.mesh.dim, V.element
    nl = el.nnodes
    fid, fc, fj = facet_adjacency(mesh)
    lam, wq = _facet_quadrature(d, 2 * el.degree)
    LF = LOCAL_EDGES[2] if d == 2 else LOCAL_FACES[3]
    opp = np.array([next(v for v in range(d + 1) if v not in lf) for lf in LF])
    PHI = np.empty((d + 1, lam.shape[0], nl))
    DPHI = np.empty((d + 1, lam.shape[0], nl, d))
    for j, lf in enumerate(LF):                       # facet vertices and cell vertices are both in ascending global order
        lamK = np.zeros((lam.shape[0], d + 1))
        lamK[:, list(lf)] = lam
        PHI[j], DPHI[j] = el.tabulate(lamK[:, 1:]), el.tabulate_grad(lamK[:, 1:])
    G, _ = cell_geometry(mesh)
    gradlam = np.concatenate([-G.sum(axis=1, keepdims=True), G], axis=1)         # grad of the barycentric coordinates
    dn = []
    for a in range(2):
        c, j = fc[:, a], fj[:, a]
        g = gradlam[c, opp[j]]
        n = -g / np.linalg.norm(g, axis=1, keepdims=True)                        # outward normal of side a
        Gn = np.einsum("fab,fb->fa", G[c], n)
        dn.append(np.einsum("fqia,fa->fqi", DPHI[j], Gn))                        # grad(phi_i) . n_a at the facet points
    X = mesh.coords[mesh.facets[fid]]
    if d == 2:
        area = np.linalg.norm(X[:, 1] - X[:, 0], axis=1)
        h = area
    else:
        area = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
        h = np.sqrt(area)
    wf = np.einsum("fqi,fib->fqb", PHI[fj[:, 0]], wind[V.cell_nodes[fc[:, 0]]])
    beta = np.sqrt(np.einsum("fqb,fqb->fq", wf, wf) + 1e-10) @ wq
    DN = np.concatenate(dn, axis=2)
    S = np.einsum("f,q,fqi,fqj->fij", 0.5 * weight * h ** 2 * beta * area, wq, DN, DN, optimize=True)
    return fid, fc, S


def cell_geometry(mesh: SimplexMesh):
    """G[c,a,b] = (J^-1)[a,b] (so d/dx_b = sum_a G[a,b] d/dxi_a) and |det J|."""
    X = mesh.coords[mesh.cells]                 # (nc, d+1, d)
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))   # J[:, :, a] = v_a - v_0
    det = np.linalg.det(J)
    G = np.linalg.inv(J)
    return G, np.abs(det)


def element_matrices(V: VectorSpace, nu: float, gamma: float, wind=None, advect: float = 1.0,
                     divform: str = "sv", cells=slice(None), parts=("visc", "div", "adv")):
    """Dense element tensors E[c, i, r, j, s] (test node i comp r, trial node j comp s)."""
    mesh, d = V.mesh, V.mesh.dim
    rt = reference_tensors(V.element)
    G, det = cell_geometry(mesh)
    G, det = G[cells], det[cells]
    nc = G.shape[0]
    nl = V.element.nnodes
    E = np.zeros((nc, nl, d, nl, d))
    need_K = ("visc" in parts and nu != 0.0) or ("div" in parts and divform == "sv" and gamma != 0.0)
    if need_K:
        # Kp[c,x,y,i,j] = det * sum_ab G[a,x] G[b,y] K[a,b,i,j]
        Kp = np.einsum("c,cax,cby,abij->cxyij", det, G, G, rt["K"], optimize=True)
        if "visc" in parts and nu != 0.0:
            lap = np.einsum("cxxij->cij", Kp)
            for r in range(d):
                E[:, :, r, :, r] += nu * lap
            # nu * d_r phi_j d_s phi_i = nu * Kp[s, r, i, j]
            E += nu * np.einsum("csrij->cirjs", Kp)
        if "div" in parts and divform == "sv" and gamma != 0.0:
            E += gamma * np.einsum("crsij->cirjs", Kp)
    if "div" in parts and divform == "pkp0" and gamma != 0.0:
        dv = np.einsum("c,car,ai->cir", det, G, rt["Dv"])          # ∫ d_r phi_i
        vol = det / (2.0 if d == 2 else 6.0)
        E += gamma * np.einsum("c,cir,cjs->cirjs", 1.0 / vol, dv, dv)
    adv1 = "adv" in parts or "adv1" in parts          # (w.grad u, v)   — the convective part
    adv2 = "adv" in parts or "adv2" in parts          # (u.grad w, v)   — the Newton part
    if (adv1 or adv2) and wind is not None and advect != 0.0:
        W = wind[V.cell_nodes[cells]]                               # (nc, nl, d)  W[c,k,b]
        if adv1:
            # (w.grad u_j, v_i): delta_rs det sum_{k,a} (sum_b W[k,b] G[a,b]) T1[k,a,i,j]
            cw = np.einsum("ckb,cab->cka", W, G)
            A1 = np.einsum("c,cka,kaij->cij", det, cw, rt["T1"], optimize=True)
            for r in range(d):
                E[:, :, r, :, r] += advect * A1
        if adv2:
            # (u_j.grad w, v_i)_{r,s} = det sum_{k,a} W[k,r] G[a,s] T2[k,a,i,j]
            A2 = np.einsum("c,ckr,cas,kaij->cirjs", det, W, G, rt["T2"], optimize=True)
            E += advect * A2
    return E


def element_parts(V: VectorSpace, wind=None, divform: str = "sv", cells=slice(None), want=("visc", "div", "adv1", "adv2")):
    """Element tensors per form, component-major: part -> E[r, s, c, i, j] (test comp r node i, trial
    comp s node j), each for unit coefficient:
      visc  (2 sym grad u, grad v)      div   (div u, div v) or (cell_avg(div u), div v)
      adv1  (w.grad u, v)               adv2  (u.grad w, v)
    The operator is linear in (nu, gamma) and only the adv parts depend on the wind, so a
    continuation run re-assembles only those per Newton step."""
    mesh, d = V.mesh, V.mesh.dim
    rt = reference_tensors(V.element)
    G, det = cell_geometry(mesh)
    G, det = G[cells], det[cells]
    nc, nl = G.shape[0], V.element.nnodes
    out = {}
    need_K = "visc" in want or ("div" in want and divform == "sv")
    if need_K:
        Kp = np.einsum("c,cax,cby,abij->xycij", det, G, G, rt["K"], optimize=True)       # [x, y, c, i, j]
        if "visc" in want:
            E = np.ascontiguousarray(np.swapaxes(Kp, 0, 1))           # d_r phi_j d_s phi_i = Kp[s, r, i, j]
            lap = np.einsum("xxcij->cij", Kp)
            for r in range(d):
                E[r, r] += lap
            out["visc"] = E
        if "div" in want and divform == "sv":
            out["div"] = Kp
    if "div" in want and divform == "pkp0":
        dv = np.einsum("c,car,ai->cri", det, G, rt["Dv"])              # ∫ d_r phi_i
        vol = det / (2.0 if d == 2 else 6.0)
        out["div"] = np.einsum("c,cri,csj->rscij", 1.0 / vol, dv, dv)
    if wind is not None and ("adv1" in want or "adv2" in want):
        W = wind[V.cell_nodes[cells]]                                   # W[c, k, b]
        if "adv1" in want:
            cw = np.einsum("ckb,cab->cka", W, G)
            A1 = np.einsum("c,cka,kaij->cij", det, cw, rt["T1"], optimize=True)
            E = np.zeros((d, d, nc, nl, nl))
            for r in range(d):
                E[r, r] = A1
            out["adv1"] = E
        if "adv2" in want:
            out["adv2"] = np.einsum("c,ckr,cas,kaij->rscij", det, W, G, rt["T2"], optimize=True)
    return out


def assemble_parts(V: VectorSpace, pattern: "BlockPattern", wind=None, divform: str = "sv",
                   want=("visc", "div", "adv1", "adv2"), chunk: int = 16384):
    """part -> block values (nnzb, d, d) on `pattern`, no boundary conditions applied."""
    d, nc = V.bs, V.mesh.nc
    want = tuple(w for w in want if wind is not None or not w.startswith("adv"))
    vals = {w: np.zeros((pattern.nnzb, d, d)) for w in want}
    nl = V.element.nnodes
    buf = {w: np.empty((d, d, nc, nl, nl)) for w in want}
    for c0 in range(0, nc, chunk):
        sl = slice(c0, min(nc, c0 + chunk))
        E = element_parts(V, wind, divform, sl, want)
        for w in want:
            buf[w][:, :, sl] = E[w]
    for w in want:
        for r in range(d):
            for s_ in range(d):
                vals[w][:, r, s_] = pattern.scatter(buf[w][r, s_])
    return vals


def assemble_velocity_block(V: VectorSpace, nu: float, gamma: float, wind=None, advect: float = 1.0,
                            divform: str = "sv", bc_nodes=None, pattern: BlockPattern | None = None,
                            parts=("visc", "div", "adv"), chunk: int = 8192) -> BSR:
    """Assemble the velocity block as BSR(bs=d); Dirichlet rows/cols zeroed, unit diagonal.

    Mirrors what Firedrake hands PETSc for `fieldsplit_0` with
    ``default_sub_matrix_type = "baij"`` (alfi/solver.py:512).
    """
    d = V.bs
    pat = pattern or BlockPattern(V)
    nc, nl = V.mesh.nc, V.element.nnodes
    vals = np.zeros((pat.nnzb, d, d))
    # element tensors in chunks, then one scatter per (r, s) component
    Efull = np.empty((nc, nl, d, nl, d)) if nc * (nl * d) ** 2 * 8 < 6e9 else None
    if Efull is not None:
        for c0 in range(0, nc, chunk):
            sl = slice(c0, min(nc, c0 + chunk))
            Efull[sl] = element_matrices(V, nu, gamma, wind, advect, divform, sl, parts)
        for r in range(d):
            for s in range(d):
                vals[:, r, s] = pat.scatter(np.ascontiguousarray(Efull[:, :, r, :, s]))
    else:                                   # pragma: no cover - very large meshes
        raise MemoryError("mesh too large for in-core assembly")
    A = BSR(V.nnodes, d, pat.rowptr, pat.colidx, vals)
    if bc_nodes is not None:
        apply_dirichlet(A, bc_nodes, pat.rows)
    return A


# --------------------------------------------------------------------------- pressure space
def assemble_divergence(V: VectorSpace, kq: int):
    """B (pressure dofs x velocity dofs) of  -(div u, q)  with q in discontinuous P_kq
    (alfi/solver.py:564-571, 615-622: ``- p*div(v) - div(u)*q``), and the block-diagonal inverse
    of the pressure mass matrix used by DGMassInv (alfi/solver.py:15-38).  Pressure dofs are
    numbered cell by cell."""
    mesh, d = V.mesh, V.mesh.dim
    G, det = cell_geometry(mesh)
    el = V.element
    nc, nl = mesh.nc, el.nnodes
    x, w = simplex_quadrature(d, el.degree + kq)
    dphi = el.tabulate_grad(x)                                   # (q, nl, d)
    if kq == 0:
        psi = np.ones((x.shape[0], 1))
    else:
        psi = LagrangeElement(d, kq).tabulate(x)                 # (q, nq)
    nq = psi.shape[1]
    Bref = np.einsum("q,qi,qja->aij", w, psi, dphi)              # ∫ psi_i d_a phi_j
    Mref = np.einsum("q,qi,qj->ij", w, psi, psi)
    # Bc[c, i, j, s] = -det sum_a G[a,s] Bref[a,i,j]
    Bc = -np.einsum("c,cas,aij->cijs", det, G, Bref)
    rows = np.repeat(np.arange(nc * nq).reshape(nc, nq, 1, 1), nl, axis=2)
    rows = np.repeat(rows, d, axis=3)
    cols = (V.cell_nodes[:, None, :, None] * d + np.arange(d)[None, None, None, :])
    cols = np.broadcast_to(cols, (nc, nq, nl, d))
    B = sp.csr_matrix((Bc.ravel(), (rows.ravel(), cols.ravel())), shape=(nc * nq, V.ndofs))
    B.sum_duplicates()
    Minv_blocks = np.linalg.inv(Mref)[None, :, :] / det[:, None, None]
    Minv = sp.block_diag(list(Minv_blocks), format="csr") if nc * nq < 200000 else \
        sp.bsr_matrix((Minv_blocks, np.arange(nc), np.arange(nc + 1)), shape=(nc * nq, nc * nq)).tocsr()
    return B, Minv


def apply_dirichlet(A: BSR, bc_nodes, rows=None):
    """Zero block rows and columns of the Dirichlet nodes, identity on their diagonal blocks."""
    if rows is None:
        rows = np.repeat(np.arange(A.nbrows), np.diff(A.rowptr))
    isbc = np.zeros(A.nbrows, dtype=bool)
    isbc[np.asarray(bc_nodes)] = True
    kill = isbc[rows] | isbc[A.colidx]
    A.vals[kill] = 0.0
    diag = kill & (rows == A.colidx)
    A.vals[diag] = np.eye(A.bs)
    return A
