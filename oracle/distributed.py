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


# ------------------------------------------------------------------------------------------ whole F-cycle
class DistHierarchy:
    """The F-cycle of oracle/hotpath.py (`fcycle`) on distributed vectors.

    Levels >= 1 carry a patch-based Layout (with the transfer's cell patches local to one rank each) and, for the
    transfer from the level below, a transfer halo on that coarser level.  Level 0 is replicated: its right-hand
    side is assembled with one all-reduce and solved redundantly (all-gather + redundant coarse
    solve).  Every operation touches only a rank's local arrays; exchanges go through Layout.update_ghosts /
    reduce_ghosts.
    """

    def __init__(self, levels, layouts, halos):
        self.levels, self.layouts, self.halos = levels, layouts, halos      # layouts[0] is None (replicated)
        self.dl = [None] + [DistLevel(levels[l], layouts[l]) for l in range(1, len(levels))]
        self.tr = [None] + [self._transfer_pieces(l) for l in range(1, len(levels))]
        self.exchanges = 0

    def _transfer_pieces(self, l):
        lv, lay, halo = self.levels[l], self.layouts[l], self.halos[l]
        out = []
        cown = lay.extra_owner
        for r in lay.ranks:
            loc = r.local
            g2l = np.full(lay.ndofs, -1, dtype=np.int64)
            g2l[loc] = np.arange(loc.size)
            P = lv.P[r.owned]                                              # owned fine rows
            if halo is None:
                Pl = P.tocsr()                                             # coarse level replicated: global columns
            else:
                hl = halo.ranks[r.rank].local
                c2l = np.full(halo.ndofs, -1, dtype=np.int64)
                c2l[hl] = np.arange(hl.size)
                Pc = P.tocoo()
                import scipy.sparse as sp
                Pl = sp.csr_matrix((Pc.data, (Pc.row, c2l[Pc.col])), shape=(r.n_owned, hl.size))
            D = lv.D[r.owned][:, loc].tocsr()
            cb = g2l[lv.cb_dofs]
            bc = g2l[lv.bc_dofs]
            cells = []
            for q in np.flatnonzero(cown == r.rank):
                I = lv.c_dofs[lv.c_offsets[q]:lv.c_offsets[q + 1]]
                assert (g2l[I] >= 0).all()
                cells.append((g2l[I], lv.c_factors[q]))
            out.append(dict(P=Pl, D=D, cb=cb[cb >= 0], cb_owned=cb[(cb >= 0) & (cb < r.n_owned)],
                            bc_owned=bc[(bc >= 0) & (bc < r.n_owned)], cells=cells))
        return out

    # ---- exchanges (counted)
    def _update(self, lay, locs):
        lay.update_ghosts(locs)
        self.exchanges += 1

    def _reduce(self, lay, locs):
        lay.reduce_ghosts(locs)
        self.exchanges += 1

    def _block_solve(self, l, bs_):
        """hp._block_solve: y = blockdiag(A0)^-1 b on the cell patches, y[cb] = b[cb]; input ghosts consistent."""
        lay = self.layouts[l]
        out = []
        for r, pieces, b in zip(lay.ranks, self.tr[l], bs_):
            y = np.zeros(r.n_local)
            for I, F in pieces["cells"]:
                if I.size:
                    y[I] += hp._solve(F, b[I])
            out.append(y)
        self._reduce(lay, out)
        for pieces, y, b in zip(self.tr[l], out, bs_):
            y[pieces["cb_owned"]] = b[pieces["cb_owned"]]
        self._update(lay, out)
        return out

    def prolong(self, l, coarse):
        """hp.prolong; `coarse` = replicated global vector (l == 1) or list of local arrays on layout l-1."""
        lay, halo = self.layouts[l], self.halos[l]
        if halo is None:
            cl = [coarse] * lay.nranks
        else:
            cl = [np.concatenate([c[:r.n_owned], np.zeros(h.ghost.size)]) for c, r, h in
                  zip(coarse, self.layouts[l - 1].ranks, halo.ranks)]
            self._update(halo, cl)
        rhs = []
        for r, pieces, c in zip(lay.ranks, self.tr[l], cl):
            v = np.zeros(r.n_local)
            v[:r.n_owned] = pieces["P"] @ c
            rhs.append(v)
        self._update(lay, rhs)
        b = []
        for r, pieces, v in zip(lay.ranks, self.tr[l], rhs):
            w = np.zeros(r.n_local)
            w[:r.n_owned] = pieces["D"] @ v
            w[pieces["cb_owned"]] = 0.0
            b.append(w)
        self._update(lay, b)
        t = self._block_solve(l, b)
        fine = [v - w for v, w in zip(rhs, t)]
        for pieces, f in zip(self.tr[l], fine):
            f[pieces["bc_owned"]] = 0.0
        self._update(lay, fine)
        return fine

    def restrict(self, l, fine):
        """hp.restrict; returns the coarse vector replicated (l == 1) or as local arrays on layout l-1."""
        lay, halo = self.layouts[l], self.halos[l]
        t = [f.copy() for f in fine]
        for pieces, v in zip(self.tr[l], t):
            v[pieces["cb"]] = 0.0                                          # bcs.apply(tildeu), ghosts included
        r_ = self._block_solve(l, t)
        r2 = []
        for r, pieces, f, v in zip(lay.ranks, self.tr[l], fine, r_):
            w = f.copy()
            w[:r.n_owned] -= pieces["D"] @ v
            r2.append(w)
        coarse_bc = self.levels[l - 1].bc_dofs
        if halo is None:
            c = np.zeros(self.levels[l - 1].n)
            for r, pieces, w in zip(lay.ranks, self.tr[l], r2):
                c += pieces["P"].T @ w[:r.n_owned]                         # the all-reduce of the small coarse vector
            self.exchanges += 1
            c[coarse_bc] = 0.0
            return c
        parts = [pieces["P"].T @ w[:r.n_owned] for r, pieces, w in zip(lay.ranks, self.tr[l], r2)]
        self._reduce(halo, parts)
        cl = self.layouts[l - 1]
        out = []
        for rc, hr, p in zip(cl.ranks, halo.ranks, parts):
            v = np.zeros(rc.n_local)
            v[:rc.n_owned] = p[:hr.n_owned]
            g2l = np.full(cl.ndofs, -1, dtype=np.int64)
            g2l[rc.owned] = np.arange(rc.n_owned)
            bc = g2l[coarse_bc]
            v[bc[bc >= 0]] = 0.0
            out.append(v)
        self._update(cl, out)
        return out

    def smooth(self, l, b, x, m):
        d = self.dl[l]
        out = d.fgmres(b, x, m)
        self.exchanges += d.stats["update"] + d.stats["reduce"]
        d.stats = {"update": 0, "reduce": 0, "allreduce": 0}
        return out

    def vcycle(self, l, b, x, m):
        if l == 0:
            return hp.coarse_solve(self.levels[0], b)                       # replicated: redundant solve
        x = self.smooth(l, b, x, m)
        Ax = self.dl[l].spmv(x)
        self.exchanges += 1
        r = [bb - a for bb, a in zip(b, Ax)]
        bc = self.restrict(l, r)
        zero = np.zeros_like(bc) if l == 1 else [np.zeros_like(v) for v in bc]
        xc = self.vcycle(l - 1, bc, zero, m)
        p = self.prolong(l, xc)
        x = [a + c for a, c in zip(x, p)]
        return self.smooth(l, b, x, m)

    def fcycle(self, b, m):
        L = len(self.levels)
        bs_ = [None] * L
        bs_[L - 1] = self.layouts[L - 1].scatter(b)
        for l in range(L - 1, 0, -1):
            bs_[l - 1] = self.restrict(l, bs_[l])
        x = np.zeros_like(bs_[0])
        for l in range(L - 1):
            x = self.vcycle(l, bs_[l], x, m)
            x = self.prolong(l + 1, x)
        out = self.vcycle(L - 1, bs_[L - 1], x, m)
        return self.layouts[L - 1].gather(out)


def build_hierarchy_layouts(prob, nranks):
    """Layouts + transfer halos for every level of a synthetic Problem (level 0 replicated)."""
    import scipy.sparse as sp

    from alfi_b200.dist import partition_patches
    from alfi_b200.halo import build_layout, transfer_halo
    layouts, halos = [None], [None]
    for l in range(1, len(prob.levels)):
        ld = prob.levels[l]
        ps, cp = ld.patches, ld.cell_patches
        owner = partition_patches(ps.offsets, ps.dofs, nranks)
        layouts.append(build_layout(ps.offsets, ps.dofs, ps.order, owner, ld.A.rowptr, ld.A.colidx, ld.V.bs,
                                    ld.V.ndofs, extra_sets=(cp.offsets, cp.dofs), nranks=nranks))
    for l in range(1, len(prob.levels)):
        ld = prob.levels[l]
        if l == 1:
            halos.append(None)
        else:
            P = ld.P.tocsr() if ld.P_dof_level else sp.kron(ld.P, sp.identity(ld.V.bs), format="csr")
            halos.append(transfer_halo(P, layouts[l], layouts[l - 1].owner))
    return layouts, halos


def fcycle(prob, levels, b, m, nranks):
    layouts, halos = build_hierarchy_layouts(prob, nranks)
    h = DistHierarchy(levels, layouts, halos)
    return h.fcycle(b, m), h


# ------------------------------------------------------------------------------------------ from LocalLevel data only
class LocalRankData:
    """What one rank computes with, rebuilt from an alfi_b200.halo.LocalLevel alone (no global arrays): the check
    that the rank-local inputs are self-sufficient."""

    def __init__(self, ll):
        import scipy.sparse as sp
        self.ll = ll
        bs = ll.bs
        n = ll.n_local_nodes

        def mat(vals):
            return sp.bsr_matrix((vals, ll.colidx, ll.rowptr), shape=(n * bs, n * bs)).tocsr()
        self.A = mat(ll.vals)
        self.patches = []
        for p in range(ll.patch_offsets.size - 1):
            I = ll.patch_dofs[ll.patch_offsets[p]:ll.patch_offsets[p + 1]].astype(np.int64)
            self.patches.append((I, np.linalg.inv(self.A[I][:, I].toarray()) if I.size else np.empty((0, 0))))
        self.cells = []
        if ll.cell_offsets is not None:
            A0 = mat(ll.a0_vals)
            self.D = mat(ll.d_vals)
            for q in range(ll.cell_offsets.size - 1):
                I = ll.cell_dofs[ll.cell_offsets[q]:ll.cell_offsets[q + 1]].astype(np.int64)
                self.cells.append((I, np.linalg.inv(A0[I][:, I].toarray()) if I.size else np.empty((0, 0))))


def local_smoother_apply(layout, locals_, data, xs):
    """PCApply_PATCH on LocalLevel data: per rank gather/solve/scatter in local numbering, then the two exchanges."""
    out = []
    for ll, d, x in zip(locals_, data, xs):
        y = np.zeros(ll.n_local)
        for p in ll.patch_order:
            I, X = d.patches[p]
            if I.size:
                y[I] += X @ x[I]
        out.append(y)
    layout.reduce_ghosts(out)
    for ll, y, x in zip(locals_, out, xs):
        bc = ll.bc_dofs[ll.bc_dofs < ll.n_owned]
        y[bc] = x[bc]
    layout.update_ghosts(out)
    return out


def local_spmv(layout, locals_, data, xs):
    out = []
    for ll, d, x in zip(locals_, data, xs):
        y = np.zeros(ll.n_local)
        y[:ll.n_owned] = (d.A @ x)[:ll.n_owned]
        out.append(y)
    layout.update_ghosts(out)
    return out
