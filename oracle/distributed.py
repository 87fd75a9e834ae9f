"""Simulated multi-rank level smoother on owned/ghost vectors (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every "rank" holds only its local arrays (alfi_b200.halo.RankLayout) and its own patches' factors; the two
exchange steps are `Layout.update_ghosts` / `Layout.reduce_ghosts`; dots are sums of owned parts (the
all-reduce).  FGMRES(m) is the serial algorithm of oracle/hotpath.py (Appendix A.4) written on those pieces.
The result must equal the serial oracle: that is the statement the device implementation will be held to.
Also counts the exchanges and bytes per smoother call (the communication model of DESIGN §6).
"""
from __future__ import annotations

import numpy as np

from . import hotpath as hp


class DistLevel:
    def __init__(self, lv: hp.OracleLevel, layout):
        self.lv, self.layout = lv, layout
        self.local_A, self.local_patches, self.local_bc = [], [], []
        for r in layout.ranks:
            loc = r.local
            g2l = np.full(layout.ndofs, -1, dtype=np.int64)
            g2l[loc] = np.arange(loc.size)
            self.local_A.append(lv.A[r.owned][:, loc].tocsr())             # owned rows, local columns
            pts = []
            for p in r.patches:
                I = lv.dofs[lv.offsets[p]:lv.offsets[p + 1]]
                assert (g2l[I] >= 0).all()
                pts.append((g2l[I], lv.factors[p]))
            self.local_patches.append(pts)
            bc = g2l[lv.bc_dofs]
            self.local_bc.append(bc[(bc >= 0) & (bc < r.n_owned)])
        self.stats = {"update": 0, "reduce": 0, "allreduce": 0}

    # ---- operators on lists of local arrays with consistent ghosts on input
    def spmv(self, xs):
        out = []
        for r, A, x in zip(self.layout.ranks, self.local_A, xs):
            y = np.zeros(r.n_local)
            y[:r.n_owned] = A @ x
            out.append(y)
        self.layout.update_ghosts(out)
        self.stats["update"] += 1
        return out

    def patch_apply(self, xs):
        out = []
        for r, pts, x, bc in zip(self.layout.ranks, self.local_patches, xs, self.local_bc):
            y = np.zeros(r.n_local)
            for I, F in pts:
                if I.size:
                    y[I] += hp._solve(F, x[I])
            out.append(y)
        self.layout.reduce_ghosts(out)                                     # ghost -> owner sum
        self.stats["reduce"] += 1
        for r, y, x, bc in zip(self.layout.ranks, out, xs, self.local_bc):
            y[bc] = x[bc]                                                  # y[bc] = x[bc] on owned Dirichlet dofs
        self.layout.update_ghosts(out)                                     # owner -> ghost for the next SpMV
        self.stats["update"] += 1
        return out

    def dot(self, xs, ys):
        self.stats["allreduce"] += 1
        return float(sum(x[:r.n_owned] @ y[:r.n_owned] for r, x, y in zip(self.layout.ranks, xs, ys)))

    def fgmres(self, bs_, xs, m):
        """hp.fgmres on distributed vectors: right preconditioned, CGS, m iterations, x = x0 + Z y."""
        ranks = self.layout.ranks
        axpy = lambda a, xs_, ys_: [y + a * x for x, y in zip(xs_, ys_)]      # noqa: E731
        Ax = self.spmv(xs)
        r0 = [b - a for b, a in zip(bs_, Ax)]
        beta = np.sqrt(self.dot(r0, r0))
        if beta == 0.0:
            return xs
        V = [[v / beta for v in r0]]
        Z = []
        H = np.zeros((m + 1, m))
        k_used = m
        for k in range(m):
            z = self.patch_apply(V[k])
            Z.append(z)
            w = self.spmv(z)
            h = np.array([self.dot(w, V[j]) for j in range(k + 1)])        # classical Gram-Schmidt: all dots first
            for j in range(k + 1):
                w = axpy(-h[j], V[j], w)
            H[:k + 1, k] = h
            hk = np.sqrt(self.dot(w, w))
            H[k + 1, k] = hk
            if hk == 0.0:
                k_used = k + 1
                break
            V.append([x / hk for x in w])
        e1 = np.zeros(k_used + 1)
        e1[0] = beta
        y, *_ = np.linalg.lstsq(H[:k_used + 1, :k_used], e1, rcond=None)
        out = [x.copy() for x in xs]
        for j in range(k_used):
            out = axpy(y[j], Z[j], out)
        return out


def smooth(lv: hp.OracleLevel, layout, b, x, m):
    """Distributed hp.smooth: returns (global result, statistics)."""
    d = DistLevel(lv, layout)
    out = d.fgmres(layout.scatter(b), layout.scatter(x), m)
    return layout.gather(out), d.stats
