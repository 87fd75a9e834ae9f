"""Stand-in for the host stack *around* the hot path: Newton + outer FGMRES + Schur fieldsplit.

In a deployment this is PETSc SNES/KSP/PCFIELDSPLIT driven by alfi's dictionaries
(alfi/solver.py:386-421, 463-499) and stays on the host, unchanged.  It is restated here only so
that the whole continuation solve (alfi/driver.py:95-129) can be run and timed around the
velocity-block multigrid — with the CPU oracle or with the CUDA library as the `fieldsplit_0`
backend — and Krylov iteration counts compared (north-star condition 3):

* SNES ``newtonls``, ``basic`` line search (full steps), <= 20 iterations, tolerances of
  solver.py:475-499;
* outer KSP ``fgmres`` (restart 30, right preconditioning, unpreconditioned norm), <= 500 its;
* ``PCFIELDSPLIT`` Schur, ``full`` factorisation, user Schur preconditioner
  ``-(nu + gamma) M_p^-1`` (DGMassInv, solver.py:15-38): the velocity solve is applied twice;
* pressure nullspace (constants, problem.py:33-38) removed after every preconditioner application;
* every level operator is *rediscretised* with the wind injected to the coarse levels
  (SURVEY A.6), the transfer operators are rebuilt once per Reynolds number.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from ..multigrid import level_input_from_synth
from .fem import assemble_divergence
from .hierarchy import prolongation_matrix
from .problem import Problem, assemble_transfer, build_problem, lid_wind

__all__ = ["ContinuationSolver", "fgmres_outer"]


def fgmres_outer(Aop, Mop, b, rtol, atol, maxit=500, restart=30):
    """Right-preconditioned flexible GMRES with restarts (PETSc KSPFGMRES defaults: classical
    Gram-Schmidt, restart 30).  Returns (x, iterations, residual history)."""
    x = np.zeros_like(b)
    r = b.copy()
    beta = np.linalg.norm(r)
    r0 = beta
    hist = [beta]
    its = 0
    if beta <= max(atol, 0.0):
        return x, 0, hist
    while its < maxit:
        m = min(restart, maxit - its)
        V = np.zeros((m + 1, b.size))
        Z = np.zeros((m, b.size))
        H = np.zeros((m + 1, m))
        cs, sn = np.zeros(m), np.zeros(m)
        g = np.zeros(m + 1)
        g[0] = beta
        V[0] = r / beta
        k_done = 0
        converged = False
        for k in range(m):
            Z[k] = Mop(V[k])
            w = Aop(Z[k])
            h = V[:k + 1] @ w
            w = w - V[:k + 1].T @ h
            H[:k + 1, k] = h
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] > 0:
                V[k + 1] = w / H[k + 1, k]
            for j in range(k):                      # apply previous rotations
                t = cs[j] * H[j, k] + sn[j] * H[j + 1, k]
                H[j + 1, k] = -sn[j] * H[j, k] + cs[j] * H[j + 1, k]
                H[j, k] = t
            rr = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = (1.0, 0.0) if rr == 0 else (H[k, k] / rr, H[k + 1, k] / rr)
            H[k, k], H[k + 1, k] = rr, 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            its += 1
            k_done = k + 1
            res = abs(g[k + 1])
            hist.append(res)
            if res <= max(rtol * r0, atol):
                converged = True
                break
        y = np.linalg.solve(np.triu(H[:k_done, :k_done]), g[:k_done])
        x = x + Z[:k_done].T @ y
        if converged:
            break
        r = b - Aop(x)
        beta = np.linalg.norm(r)
        if beta <= max(rtol * r0, atol):
            break
    return x, its, hist


SNES_MAX_IT, KSP_MAX_IT = 20, 500                 # outer_base / outer_fieldsplit, alfi/solver.py:450-474


def tolerances(tdim):
    """Newton / Krylov tolerances of `get_parameters` without --high-accuracy (alfi/solver.py:484-499)."""
    if tdim == 2:
        return dict(ksp_rtol=1e-9, ksp_atol=1e-10, snes_rtol=1e-9, snes_atol=1e-8)
    return dict(ksp_rtol=1e-8, ksp_atol=1e-8, snes_rtol=1e-8, snes_atol=1e-8)


def dg_mass_inv_apply(Minv, nu, gamma, x):
    """`alfi.solver.DGMassInv.apply` (solver.py:32-35): y = -(nu + gamma) M_p^-1 x, the Schur complement
    approximation of the augmented-Lagrangian preconditioner (fieldsplit_1, solver.py:386-390)."""
    return -(float(nu) + float(gamma)) * (Minv @ x)


@dataclass
class ContinuationSolver:
    """Lid-driven cavity continuation in Reynolds number around a pluggable velocity-block PC.

    `backend` has three methods: ``setup(levels: list[LevelInput])`` (once), ``update_operators(levels)``
    (per Newton step), ``update_transfers(levels)`` (per Reynolds number) and ``apply(b) -> x``."""
    config: object
    backend: object
    verbose: bool = False
    # "host": numpy outer FGMRES + Schur pieces around backend.apply (what PETSc does in a deployment);
    # "schur": the Schur-complement preconditioner application on the device (backend.schur_apply), host FGMRES;
    # "device": the whole linear solve of a Newton step on the device (backend.outer_solve)
    outer: str = "host"
    prob: Problem = field(init=False)

    def __post_init__(self):
        cfg = self.config
        self.prob = build_problem(cfg, nu=1.0)
        fine = self.prob.finest
        self.d = fine.V.bs
        kq = cfg.k - 1 if cfg.discretisation == "sv" else 0
        self.B, self.Minv = assemble_divergence(fine.V, kq)
        self.nu_dofs, self.np_dofs = fine.ndofs, self.B.shape[0]
        # injection of the fine velocity to every coarser level (point evaluation at coarse nodes)
        self.inject = {}
        hier = [l.level for l in self.prob.levels]
        for l in range(len(hier) - 1, 0, -1):
            c2f = hier[l - 1].c2f
            f2c = np.full((hier[l].mesh.nc, (self.d + 1) if cfg.bary else 1), -1, dtype=np.int64)
            fill = np.zeros(hier[l].mesh.nc, dtype=np.int64)
            for c in range(c2f.shape[0]):
                for f in c2f[c]:
                    if fill[f] < f2c.shape[1]:
                        f2c[f, fill[f]] = c
                        fill[f] += 1
            self.inject[l - 1] = prolongation_matrix(self.prob.levels[l].V, self.prob.levels[l - 1].V, f2c)
        self.u = np.zeros((fine.V.nnodes, self.d))
        bcv = lid_wind(fine.V.node_coords, fine.V.mesh.extent)
        self.u[fine.bc_nodes] = bcv[fine.bc_nodes]
        self.p = np.zeros(self.np_dofs)
        self._setup_done = False
        self.history = []

    # -- assembly (the host's job in a deployment) -------------------------------------------------
    def _winds(self):
        w = {len(self.prob.levels) - 1: self.u}
        for l in range(len(self.prob.levels) - 2, -1, -1):
            w[l] = self.inject[l] @ w[l + 1]
        return w

    def _assemble(self, nu, gamma, advect):
        """Jacobian velocity block on every level (wind injected) and, on the finest level, the
        operator K + N1(u) of the residual.  Only the advection parts depend on u."""
        from .problem import _bsr, _linear_parts, assemble_level
        cfg = self.config
        winds = self._winds()
        nl = len(self.prob.levels)
        stab = getattr(self, "_stab_winds", None)
        for l, ld in enumerate(self.prob.levels):
            assemble_level(cfg, ld, nu, gamma, advect, wind=winds[l], stab_wind=None if stab is None else stab[l])
            if l == nl - 1:
                lin = _linear_parts(cfg, ld)
                vals = nu * lin["visc"] + gamma * lin["div"]
                if advect != 0.0:
                    vals = vals + advect * ld.adv1
                    if cfg.stabilisation == "burman":           # F += advect * stabilisation_form (solver.py:233-234)
                        vals = vals + advect * ld.stab
                self.M1 = _bsr(ld, vals).to_csr()
        self.Afine = self.prob.finest.A.to_csr()

    def _residual(self):
        fine = self.prob.finest
        Fu = self.M1 @ self.u.ravel() + self.B.T @ self.p
        Fu[fine.bc_dofs] = 0.0
        Fp = self.B @ self.u.ravel()
        return Fu, Fp

    # -- one Reynolds number ----------------------------------------------------------------------
    def solve(self, re, min_newton=0, ksp_tol=None, max_newton=None):
        """One Reynolds number.  min_newton: take at least that many Newton steps even if the residual already meets
        the tolerances; ksp_tol = (rtol, atol) overrides the linear tolerances (both used to polish a converged state:
        scripts/cont3d.py — the reference's ksp_atol would stop the linear solve of such a step at iteration 0);
        max_newton: take at most that many (with min_newton = max_newton a run follows another run's Newton counts, so
        that a stopping test decided in the third digit of a residual norm does not fork the two trajectories; the
        residual after every step is returned, so what the free-running test would have done is still known)."""
        cfg = self.config
        tdim = self.d
        tol = tolerances(tdim)
        if ksp_tol is not None:
            tol["ksp_rtol"], tol["ksp_atol"] = ksp_tol
        # the wind inside the stabilisation is the state BEFORE this Reynolds number's Newton solve (solver.py:270-271:
        # stabilisation.update(z) precedes solver.solve(); Stabilisation.update injects it to the coarse levels)
        self._stab_winds = {l: w.copy() for l, w in self._winds().items()} if cfg.stabilisation == "burman" else None
        nu = cfg.length * 1.0 / re if re > 0 else cfg.length
        advect = 1.0 if re > 0 else 0.0
        gamma = cfg.gamma
        fine = self.prob.finest
        t0 = time.time()
        for ld in self.prob.levels[1:]:
            assemble_transfer(cfg, ld, nu, gamma)
        lin_its, newton = 0, 0
        fnorm0 = None
        fhist = []
        nbc = fine.bc_dofs
        for newton in range(SNES_MAX_IT + 1):
            self._assemble(nu, gamma, advect)
            Fu, Fp = self._residual()
            fnorm = np.sqrt(Fu @ Fu + Fp @ Fp)
            fnorm0 = fnorm if fnorm0 is None else fnorm0
            fhist.append(float(fnorm))
            if self.verbose:
                print("  Re %g  SNES %d  |F| = %.6e" % (re, newton, fnorm), flush=True)
            if (newton >= min_newton and fnorm <= max(tol["snes_atol"], tol["snes_rtol"] * fnorm0)) or newton == SNES_MAX_IT \
                    or (max_newton is not None and newton >= max_newton):
                break
            levels = [level_input_from_synth(l) for l in self.prob.levels]
            if not self._setup_done:
                self.backend.setup(levels)
                if self.outer != "host":
                    self.backend.setup_outer(self.B, self.Minv, nbc)
                self._setup_done = True
                self._transfer_key = (nu, gamma)
            else:
                self.backend.update_operators(levels)
                if self._transfer_key != (nu, gamma):
                    self.backend.update_transfers(levels)
                    self._transfer_key = (nu, gamma)
            A, B, Minv = self.Afine, self.B, self.Minv
            nu_d = self.nu_dofs

            def Jop(z):
                zu, zp = z[:nu_d], z[nu_d:]
                ou = A @ zu
                btp = B.T @ zp
                btp[nbc] = 0.0
                ou += btp
                zu0 = zu.copy()
                zu0[nbc] = 0.0
                return np.concatenate([ou, B @ zu0])

            def Pop(r):
                ru, rp = r[:nu_d], r[nu_d:]
                y1 = self.backend.apply(ru)
                y10 = y1.copy()
                y10[nbc] = 0.0
                yp = dg_mass_inv_apply(Minv, nu, gamma, rp - B @ y10)
                yp -= yp.mean()                              # constant-pressure nullspace
                t = B.T @ yp
                t[nbc] = 0.0
                yu = self.backend.apply(ru - t)
                return np.concatenate([yu, yp])

            rhs = -np.concatenate([Fu, Fp])
            if self.outer == "device":
                dz, its, _ = self.backend.outer_solve(nu, gamma, rhs, tol["ksp_rtol"], tol["ksp_atol"], KSP_MAX_IT, 30)
            elif self.outer == "schur":
                dz, its, _ = fgmres_outer(Jop, lambda r: self.backend.schur_apply(nu, gamma, r), rhs, tol["ksp_rtol"], tol["ksp_atol"])
            else:
                dz, its, _ = fgmres_outer(Jop, Pop, rhs, tol["ksp_rtol"], tol["ksp_atol"])
            lin_its += its
            self.u += dz[:nu_d].reshape(self.u.shape)
            self.p += dz[nu_d:]
            if self.verbose:
                print("      KSP iterations %d" % its, flush=True)
        self.p -= self.p.mean()
        info = {"Re": re, "nu": nu, "linear_iter": lin_its, "nonlinear_iter": newton,
                "time": (time.time() - t0) / 60.0, "residual": fnorm, "residual0": fnorm0, "residual_history": fhist,
                "snes_tolerance": max(tol["snes_atol"], tol["snes_rtol"] * fnorm0)}
        self.history.append(info)
        return info
