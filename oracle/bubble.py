"""CPU ORACLE — test infrastructure only.  Literal restatement of alfi/bubble.py: the five C
kernels (`split`, `splitadj`, `combine`, `combineadj`, `count`; bubble.py:57-185) as per-cell
Python loops with the same 8x4 / 4x8 tables and multiplicity divides, the facet "solve" of
bubble.py:25-39 and the prolong / restrict sequences of bubble.py:204-265.  The standard
prolongations of the P1 and FacetBubble parts (Firedrake `prolong`, not in the reference tree) are
point evaluations: linear interpolation at fine vertices, coarse bubbles at fine face centroids."""
from __future__ import annotations

import numpy as np

A_SPLIT = np.vstack([np.eye(4), np.zeros((4, 4))])                                    # a[8][4], bubble.py:64-71
B_SPLIT = np.vstack([-(np.ones((4, 4)) - np.eye(4)) / 3.0, np.eye(4)])                # b[8][4], bubble.py:73-80
A_COMB = np.hstack([np.eye(4), (np.ones((4, 4)) - np.eye(4)) / 3.0])                  # a[4][8], bubble.py:130-133
B_COMB = np.hstack([np.zeros((4, 4)), np.eye(4)])                                     # b[4][8], bubble.py:134-137


class LiteralBubbleTransfer:
    def __init__(self, Vc, Vf, c2f):
        self.Vc, self.Vf, self.c2f = Vc, Vf, c2f
        self.cnt = {}
        for name, V in (("c", Vc), ("f", Vf)):
            cv = np.zeros(V.nnodes)
            for c in range(V.mesh.nc):                       # count kernel, bubble.py:176-185
                cv[V.cell_nodes[c]] += 1
            self.cnt[name] = cv

    # cell-local views: local nodes 0-3 vertices, 4-7 faces (face f opposite vertex f)
    def split(self, V, both, cnt):
        p1 = np.zeros_like(both)
        fb = np.zeros_like(both)
        for c in range(V.mesh.nc):
            nodes = V.cell_nodes[c]
            loc = both[nodes]                                # (8, 3)
            p1[nodes[:4]] += A_SPLIT.T @ loc                 # p1[i] += a[k][i] both[k]
            fb[nodes[4:]] += B_SPLIT.T @ loc
        p1[V.vertex_nodes[:, 0]] /= cnt[V.vertex_nodes[:, 0], None]
        fb[V.face_nodes[:, 0]] /= cnt[V.face_nodes[:, 0], None]
        return p1, fb

    def combine(self, V, p1, fb, cnt):
        both = np.zeros_like(p1)
        for c in range(V.mesh.nc):
            nodes = V.cell_nodes[c]
            both[nodes] += A_COMB.T @ p1[nodes[:4]] + B_COMB.T @ fb[nodes[4:]]
        return both / cnt[:, None]

    def scale_normal(self, V, fb):
        out = fb.copy()
        X = V.mesh.coords[V.mesh.faces]
        n = np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0])
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        fn = V.face_nodes[:, 0]
        c = fb[fn]
        cn = (c * n).sum(axis=1, keepdims=True)
        out[fn] = cn * n / 0.625 + (c - cn * n)              # ainv * assemble(L), bubble.py:36-39
        return out

    def point_prolong(self, which, coarse):
        """standard prolong of the P1 (which='p1') or bubble (which='fb') part by point evaluation."""
        Vc, Vf = self.Vc, self.Vf
        el = Vc.element
        out = np.zeros((Vf.nnodes, coarse.shape[1]))
        fine_nodes = Vf.vertex_nodes[:, 0] if which == "p1" else Vf.face_nodes[:, 0]
        done = np.zeros(Vf.nnodes, dtype=bool)
        want = np.zeros(Vf.nnodes, dtype=bool)
        want[fine_nodes] = True
        for c in range(Vc.mesh.nc):
            X = Vc.mesh.coords[Vc.mesh.cells[c]]
            J = (X[1:] - X[0]).T
            cn = Vc.cell_nodes[c]
            for fcell in self.c2f[c]:
                for node in Vf.cell_nodes[fcell]:
                    if done[node] or not want[node]:
                        continue
                    xi = np.linalg.solve(J, Vf.node_coords[node] - X[0])
                    lam = np.concatenate(([1 - xi.sum()], xi))
                    if which == "p1":
                        out[node] = lam @ coarse[cn[:4]]
                    else:
                        bub = np.array([27.0 * np.prod(np.delete(lam, f)) for f in range(4)])
                        out[node] = bub @ coarse[cn[4:]]
                    done[node] = True
        return out

    def prolong(self, coarse):
        """bubble.py:233-265"""
        coarse = coarse.reshape(self.Vc.nnodes, 3)
        p1c, fbc = self.split(self.Vc, coarse, self.cnt["c"])
        fbc = self.scale_normal(self.Vc, fbc)
        p1f = self.point_prolong("p1", p1c)
        fbf = self.point_prolong("fb", fbc)
        return self.combine(self.Vf, p1f, fbf, self.cnt["f"]).ravel()
