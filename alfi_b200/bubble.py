"""Flux-preserving transfer for the 3-D [P1+FacetBubble]^3 space — host side of ``alfi.bubble``.

The reference's `BubbleTransfer` (alfi/bubble.py:10-265; selected for 3-D CG1-based spaces at
alfi/transfer.py:334-356) fixes the standard prolongation, which under-estimates the flux of a
coarse facet bubble across the coarse facets by the factor 0.625 (bubble.py:247-250):

    prolong(c):  split c into its P1 and facet-bubble parts           (bubble.py:58-91, 237-244)
                 scale the *normal* component of every bubble by 1/0.625   (bubble.py:25-39, 251-253)
                 prolong the P1 and the bubble part separately        (bubble.py:256-257)
                 combine on the fine mesh                             (bubble.py:126-147, 259-265)
    restrict  =  the adjoint steps in reverse order                   (bubble.py:204-231)

Every step is linear, so the whole transfer is one sparse matrix on scalar dofs.  It is built
here from five factors and handed to the device as a dof-level CSR
(`alfib_transfer_set(..., dof_level=1)`): unlike the Lagrange case it is not "scalar (x) I_3",
because the normal scaling ``I + 0.6 n n^T`` couples the components on every facet.  The
reference's restrict is the exact transpose, which the library applies from the same matrix.
The literal per-cell restatement of the reference's C kernels is `oracle/bubble.py`.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .synth.fem import VectorSpace
from .synth.hierarchy import prolongation_matrix

__all__ = ["bubble_transfer_matrix", "BubbleTransfer", "FLUX_FACTOR"]

FLUX_FACTOR = 0.625          # alfi/bubble.py:36


def _split_matrix(V: VectorSpace):
    """nodal P1FB -> (P1 vertex values, hierarchical bubble coefficients): scalar, nodes x nodes.
    fb_f = u(face centroid) - (1/3) sum_{v in f} u_v   (matrix `b` of bubble.py:71-78)."""
    m = V.mesh
    nn = V.nnodes
    vn, fn = V.vertex_nodes[:, 0], V.face_nodes[:, 0]
    rows = np.concatenate([vn, fn, np.repeat(fn, 3)])
    cols = np.concatenate([vn, fn, vn[m.faces].ravel()])
    vals = np.concatenate([np.ones(vn.size), np.ones(fn.size), np.full(3 * fn.size, -1.0 / 3.0)])
    return sp.csr_matrix((vals, (rows, cols)), shape=(nn, nn))


def _combine_matrix(V: VectorSpace):
    """inverse change of basis (matrices `a`, `b` of bubble.py:129-137): u_f = fb_f + mean of p1 on f."""
    m = V.mesh
    nn = V.nnodes
    vn, fn = V.vertex_nodes[:, 0], V.face_nodes[:, 0]
    rows = np.concatenate([vn, fn, np.repeat(fn, 3)])
    cols = np.concatenate([vn, fn, vn[m.faces].ravel()])
    vals = np.concatenate([np.ones(vn.size), np.ones(fn.size), np.full(3 * fn.size, 1.0 / 3.0)])
    return sp.csr_matrix((vals, (rows, cols)), shape=(nn, nn))


def _normal_scaling(V: VectorSpace):
    """dof-level block diagonal: identity on vertex nodes, I + (1/0.625 - 1) n n^T on face nodes.

    `ainv * assemble(L)` of bubble.py:25-39 with the facet mass of the bubble cancelling: only the
    bubble of a facet is non-zero on that facet, so the facet system is diagonal and the solve
    leaves c + (1/0.625 - 1)(c.n) n."""
    m, d = V.mesh, 3
    X = m.coords[m.faces]
    nrm = np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    blocks = np.tile(np.eye(d), (V.nnodes, 1, 1))
    blocks[V.face_nodes[:, 0]] += (1.0 / FLUX_FACTOR - 1.0) * nrm[:, :, None] * nrm[:, None, :]
    idx = np.arange(V.nnodes)
    return sp.bsr_matrix((blocks, idx, np.arange(V.nnodes + 1)), shape=(V.ndofs, V.ndofs)).tocsr()


def _hier_prolongation(Vc: VectorSpace, Vf: VectorSpace, c2f):
    """(P1 part, bubble part) prolonged separately: fine hierarchical coefficients from coarse ones.
    P1: nested linear interpolation.  Bubbles: the coarse facet bubbles evaluated at the fine face
    centroids (FIAT's FacetBubble has point evaluations at the facet centroids as dofs)."""
    P1c, P1f = VectorSpace(Vc.mesh, 1), VectorSpace(Vf.mesh, 1)
    Pp1 = prolongation_matrix(P1c, P1f, c2f)                       # fine vertices x coarse vertices (P1 numbering)
    # P1 spaces number nodes by first encounter too: map to the P1FB node numbers
    pc = sp.csr_matrix((np.ones(P1c.nnodes), (Vc.vertex_nodes[:, 0], P1c.vertex_nodes[:, 0])), shape=(Vc.nnodes, P1c.nnodes))
    pf = sp.csr_matrix((np.ones(P1f.nnodes), (Vf.vertex_nodes[:, 0], P1f.vertex_nodes[:, 0])), shape=(Vf.nnodes, P1f.nnodes))
    Pv = pf @ Pp1 @ pc.T
    # bubble part: evaluate the full coarse nodal basis at the fine nodes, keep face -> face entries
    # of the *hierarchical* bubble functions: b_f is the nodal face function itself
    Pfull = prolongation_matrix(Vc, Vf, c2f)                       # values of coarse nodal functions at fine nodes
    isface_c = np.zeros(Vc.nnodes, dtype=bool)
    isface_c[Vc.face_nodes[:, 0]] = True
    isface_f = np.zeros(Vf.nnodes, dtype=bool)
    isface_f[Vf.face_nodes[:, 0]] = True
    Pb = sp.diags(isface_f.astype(float)) @ Pfull @ sp.diags(isface_c.astype(float))
    return (Pv + Pb).tocsr()


def bubble_transfer_matrix(Vc: VectorSpace, Vf: VectorSpace, c2f) -> sp.csr_matrix:
    """dof-level CSR (fine dofs x coarse dofs) of `BubbleTransfer.prolong` (bubble.py:233-265)."""
    d = 3
    I3 = sp.identity(d, format="csr")
    split_c = sp.kron(_split_matrix(Vc), I3, format="csr")
    comb_f = sp.kron(_combine_matrix(Vf), I3, format="csr")
    Ph = sp.kron(_hier_prolongation(Vc, Vf, c2f), I3, format="csr")
    P = (comb_f @ Ph @ _normal_scaling(Vc) @ split_c).tocsr()
    P.data[np.abs(P.data) < 1e-14] = 0.0
    P.eliminate_zeros()
    P.sort_indices()
    return P


class BubbleTransfer:
    """Same call surface as the reference class: ``prolong(coarse, fine)`` / ``restrict(fine, coarse)``
    on flat dof arrays; the matrix is built once per level pair (bubble.py:10-201)."""

    def __init__(self, Vc: VectorSpace, Vf: VectorSpace, c2f):
        self.P = bubble_transfer_matrix(Vc, Vf, c2f)
        self.PT = self.P.T.tocsr()

    def prolong(self, coarse, fine):
        fine[...] = (self.P @ np.asarray(coarse).ravel()).reshape(np.shape(fine))

    def restrict(self, fine, coarse):
        coarse[...] = (self.PT @ np.asarray(fine).ravel()).reshape(np.shape(coarse))
