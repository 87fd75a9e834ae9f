"""TEST INFRASTRUCTURE: a stand-in for the PETSc objects that interpret alfi's `fieldsplit_0` dictionary.

petsc4py / PETSc are absent here, so the fine-grained drop-in of INTEGRATION.md §1 — PETSc keeps KSPRichardson(1),
PCMG "full", the level KSPFGMRES and the coarse LU on the host and only *instantiates the python objects the
dictionary names* — is exercised with this restatement of those PETSc parts:

* the level PC is whatever class ``mg_levels.pc_python_type`` names, created through importlib like
  ``PCPythonSetType`` does, given a PC whose options are the ``mg_levels`` sub-dictionary under the level's prefix,
  and driven through ``setUp`` / ``apply`` with Vec-like objects;
* prolongation / restriction call the transfer object's ``prolong(coarse, fine)`` / ``restrict(fine, coarse)`` — the
  callables alfi registers with Firedrake's TransferManager (solver.py:595-596) — and zero the Dirichlet rows of the
  level they write (Firedrake's transfer Mats, Appendix A.6);
* FGMRES(m) is oracle.hotpath.fgmres, the coarse solve a dense LU.
"""
import importlib

import numpy as np
import scipy.linalg as sla

from alfi_b200.synth.fakepetsc import FakePC, FakeVec
from oracle import hotpath as hp


def python_pc(dotted):
    mod, _, cls = dotted.rpartition(".")
    return getattr(importlib.import_module(mod), cls)()


class MiniFieldsplit0:
    def __init__(self, fs0: dict, operators, bc_dofs, dms, adapters, transfer):
        """fs0: the reference's fieldsplit_0 dictionary; operators / bc_dofs / dms / adapters: per level, coarsest
        first (scipy CSR, Dirichlet dofs, DM stand-in, alfi_b200 HostAdapter); transfer: the registered transfer object."""
        assert fs0["ksp_type"] == "richardson" and fs0["ksp_max_it"] == 1 and fs0["pc_type"] == "mg" and fs0["pc_mg_type"] == "full"
        lv = fs0["mg_levels"]
        assert lv["ksp_type"] == "fgmres" and lv["pc_type"] == "python" and lv["ksp_convergence_test"] == "skip"
        self.m = int(lv["ksp_max_it"])
        self.A, self.bc, self.transfer = operators, bc_dofs, transfer
        self.pcs, self.objs = [None], [None]
        for l in range(1, len(operators)):
            pc = FakePC(dms[l], options=dict(lv), prefix="", attrs={"alfi_b200_adapter": adapters[l]})
            obj = python_pc(lv["pc_python_type"])
            obj.setUp(pc)
            self.pcs.append(pc)
            self.objs.append(obj)
        self.coarse_lu = sla.lu_factor(operators[0].toarray())

    def setUp(self):
        """PCSetUp on every level = once per Newton step."""
        for pc, obj in zip(self.pcs[1:], self.objs[1:]):
            obj.setUp(pc)
        self.coarse_lu = sla.lu_factor(self.A[0].toarray())

    def _smooth(self, l, b, x):
        def Mop(v):
            y = FakeVec(v.size)
            self.objs[l].apply(self.pcs[l], FakeVec(v), y)
            return y.array
        return hp.fgmres(lambda v: self.A[l] @ v, Mop, b, x, self.m)

    def _prolong(self, l, xc):
        f = np.empty(self.A[l].shape[0])
        self.transfer.prolong(np.ascontiguousarray(xc), f, level=l)
        f[self.bc[l]] = 0.0
        return f

    def _restrict(self, l, r):
        c = np.empty(self.A[l - 1].shape[0])
        self.transfer.restrict(np.ascontiguousarray(r), c, level=l)
        c[self.bc[l - 1]] = 0.0
        return c

    def _v(self, l, b, x):
        if l == 0:
            return sla.lu_solve(self.coarse_lu, b)
        x = self._smooth(l, b, x)
        bc = self._restrict(l, b - self.A[l] @ x)
        x = x + self._prolong(l, self._v(l - 1, bc, np.zeros_like(bc)))
        return self._smooth(l, b, x)

    def apply(self, b):
        L = len(self.A)
        bs = [None] * L
        bs[L - 1] = b
        for l in range(L - 1, 0, -1):
            bs[l - 1] = self._restrict(l, bs[l])
        x = np.zeros_like(bs[0])
        for l in range(L - 1):
            x = self._prolong(l + 1, self._v(l, bs[l], x))
        return self._v(L - 1, bs[L - 1], x)
