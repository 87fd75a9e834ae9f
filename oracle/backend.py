"""CPU ORACLE — test infrastructure only: `fieldsplit_0` backend for the continuation stand-in
(alfi_b200.synth.outer.ContinuationSolver) built on oracle.hotpath.  Takes the same LevelInput
hand-over data the CUDA library gets."""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from . import hotpath as hp


class OracleBackend:
    def __init__(self, smoothing, mode="inverse"):
        self.m, self.mode = smoothing, mode
        self.levels = None

    def _level(self, l, li, old=None):
        A = sp.bsr_matrix((li.vals, li.colidx, li.rowptr), shape=(li.n_nodes * li.bs,) * 2).tocsr()
        lv = hp.OracleLevel(A=A, bc_dofs=np.asarray(li.bc_dofs), bs=li.bs)
        if li.patch_offsets is not None:
            lv.offsets, lv.dofs, lv.order = li.patch_offsets, li.patch_dofs, li.patch_order
            corr = None
            if getattr(li, "patch_corr_off", None) is not None:
                corr = (li.patch_corr_off, li.patch_corr_rows, li.patch_corr_cols, li.patch_corr_vals)
            lv.factors = hp.factor_patches(hp.patch_matrices(A, lv.offsets, lv.dofs, corr), self.mode)
        if li.P is not None:
            if old is not None:
                lv.P = old.P
            else:
                lv.P = li.P.tocsr() if li.P_dof_level else sp.kron(li.P, sp.identity(li.bs), format="csr")
            if old is not None:
                lv.D, lv.cb_dofs, lv.c_offsets, lv.c_dofs, lv.c_factors = old.D, old.cb_dofs, old.c_offsets, old.c_dofs, old.c_factors
        if l == 0:
            lv.coarse_lu = sla.lu_factor(A.toarray())
        return lv

    def _transfer(self, lv, li):
        if li.cell_offsets is None:
            return
        mk = lambda v: sp.bsr_matrix((v, li.colidx, li.rowptr), shape=(li.n_nodes * li.bs,) * 2).tocsr()   # noqa: E731
        lv.D, lv.cb_dofs = mk(li.d_vals), np.asarray(li.cb_dofs)
        lv.c_offsets, lv.c_dofs = li.cell_offsets, li.cell_dofs
        lv.c_factors = hp.factor_patches(hp.patch_matrices(mk(li.a0_vals), lv.c_offsets, lv.c_dofs), "lu")

    def setup(self, levels):
        self.levels = [self._level(l, li) for l, li in enumerate(levels)]
        self.update_transfers(levels)

    def update_operators(self, levels):
        self.levels = [self._level(l, li, old) for (l, li), old in zip(enumerate(levels), self.levels)]

    def update_transfers(self, levels):
        for lv, li in zip(self.levels, levels):
            self._transfer(lv, li)

    def apply(self, b):
        return hp.fcycle(self.levels, np.asarray(b, dtype=np.float64), self.m)
