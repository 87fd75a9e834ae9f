"""Minimal stand-ins for the petsc4py objects the plugins touch (petsc4py is absent here).

Only what `alfi_b200.pc`, `alfi_b200.relaxation` and `alfi_b200.transfer` call: ``PC.getDM``,
``getOptionsPrefix``, ``getOperators``, ``getAttr``; ``Vec.array_r`` / ``array_w``.  The
`SynthAdapter` is the :class:`alfi_b200.pc.HostAdapter` for :mod:`alfi_b200.synth` problems.
"""
from __future__ import annotations

import numpy as np

from ..multigrid import level_input_from_synth
from ..pc import HostAdapter


class FakeVec:
    def __init__(self, n_or_array):
        self.array = np.zeros(n_or_array) if np.isscalar(n_or_array) else np.array(n_or_array, dtype=np.float64)

    @property
    def array_r(self):
        return self.array

    @property
    def array_w(self):
        return self.array


class FakePC:
    def __init__(self, dm, options=None, prefix="", attrs=None, operators=(None, None)):
        self.dm, self.options, self.prefix = dm, options or {}, prefix
        self.attrs = attrs or {}
        self.operators = operators

    def getDM(self):
        return self.dm

    def getOptionsPrefix(self):
        return self.prefix

    def getOperators(self):
        return self.operators

    def getAttr(self, name):
        return self.attrs.get(name)


class SynthAdapter(HostAdapter):
    """Reads operator / space / bcs of one level (PatchPC) or all levels (VelocityMGPC) of a
    synthetic Problem."""

    def __init__(self, problem, level=None, device=0, deterministic=False, restriction=True):
        self.problem, self.level = problem, level
        self.device, self.deterministic, self.restriction = device, deterministic, restriction
        self.smoothing = problem.config.m

    def _ld(self):
        return self.problem.levels[self.level]

    def operator(self, pc):
        A = self._ld().A
        return A.rowptr, A.colidx, A.vals, False

    def function_space(self, pc):
        return self._ld().V

    def plex(self, pc):
        return self._ld().level.plex

    def bc_nodes(self, pc):
        return self._ld().bc_nodes

    def levels(self, pc):
        return [level_input_from_synth(l) for l in self.problem.levels]

    def parameters(self, pc):
        return (self.problem.nu, self.problem.gamma)

    def patch_corrections(self, pc, patches):
        ps = self._ld().patches
        if ps is None or ps.corrections is None:
            return None
        return ps.corrections.off, ps.corrections.rows, ps.corrections.cols, ps.corr_vals

    def pressure_operators(self, pc):
        """(B, M_p^-1, Dirichlet velocity dofs) of the finest level — what alfi_b200.ALFieldsplitPC adds to the levels."""
        from .fem import assemble_divergence
        cfg, fine = self.problem.config, self.problem.finest
        B, Minv = assemble_divergence(fine.V, cfg.k - 1 if cfg.discretisation == "sv" else 0)
        return B, Minv, fine.bc_dofs
