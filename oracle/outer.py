"""CPU ORACLE — test infrastructure only: the pieces that bracket `fieldsplit_0` in alfi's outer solver, restated in
numpy so that the device versions (csrc/outer.cu: alfib_schur_apply, alfib_jacobian_apply, alfib_outer_solve) have
something to be compared with.  Parity unpinned by the reference: the arithmetic is PETSc's (PCFIELDSPLIT, KSPFGMRES;
SURVEY §8c), restated from its documented semantics.

* `schur_apply`     PCFIELDSPLIT schur / full (alfi/solver.py:405-421; PCApply_FieldSplit_Schur, FACT_FULL):
                    y1 = A^-1 r_u ; y_p = S^-1 (r_p - B y1) ; y_u = A^-1 (r_u - B^T y_p), with
                    S^-1 = alfi.solver.DGMassInv.apply = -(nu + gamma) M_p^-1 (solver.py:24, 32-35) and the constant
                    pressure nullspace (alfi/problem.py:33-38) projected out of y_p;
* `jacobian_apply`  MatMult of the nest matrix [A B^T; B 0] whose off-diagonal blocks carry the velocity bcs
                    (rows of B^T / columns of B at Dirichlet dofs are zero);
* `fgmres`          KSPFGMRES as configured by solver.py:463-474 (right preconditioning, classical Gram-Schmidt,
                    restart 30, zero initial guess, unpreconditioned recurrence residual against max(rtol |b|, atol)).
"""
from __future__ import annotations

import numpy as np

from .backend import OracleBackend


def schur_apply(apply_A, B, Minv, bc_dofs, nu, gamma, r, remove_constant=True):
    nu_d = B.shape[1]
    ru, rp = r[:nu_d], r[nu_d:]
    y1 = np.array(apply_A(ru), dtype=np.float64)
    y1[bc_dofs] = 0.0
    yp = -(float(nu) + float(gamma)) * (Minv @ (rp - B @ y1))
    if remove_constant:
        yp = yp - yp.mean()
    t = B.T @ yp
    t[bc_dofs] = 0.0
    yu = apply_A(ru - t)
    return np.concatenate([yu, yp])


def jacobian_apply(A, B, bc_dofs, z):
    nu_d = B.shape[1]
    zu, zp = z[:nu_d], z[nu_d:]
    ou = A @ zu
    t = B.T @ zp
    t[bc_dofs] = 0.0
    zu0 = zu.copy()
    zu0[bc_dofs] = 0.0
    return np.concatenate([ou + t, B @ zu0])


def fgmres(Aop, Mop, b, rtol, atol, maxit=500, restart=30):
    """Returns (x, iterations, residual history).  Saad's FGMRES with Givens rotations, as PETSc implements it."""
    n = b.size
    x = np.zeros(n)
    r = b.copy()
    beta = float(np.sqrt(r @ r))
    r0, hist, its = beta, [beta], 0
    if not beta > max(atol, 0.0):
        return x, 0, hist
    target = max(rtol * r0, atol)
    while its < maxit:
        m = min(restart, maxit - its)
        V, Z = np.zeros((m + 1, n)), np.zeros((m, n))
        H = np.zeros((m + 1, m))
        c, s, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
        g[0] = beta
        V[0] = r / beta
        kd, done = 0, False
        for k in range(m):
            Z[k] = Mop(V[k])
            w = Aop(Z[k])
            h = V[:k + 1] @ w                      # classical Gram-Schmidt: all dots against the unmodified w
            w = w - h @ V[:k + 1]
            H[:k + 1, k], H[k + 1, k] = h, np.sqrt(w @ w)
            if H[k + 1, k] > 0.0:
                V[k + 1] = w / H[k + 1, k]
            for j in range(k):
                H[j, k], H[j + 1, k] = c[j] * H[j, k] + s[j] * H[j + 1, k], -s[j] * H[j, k] + c[j] * H[j + 1, k]
            rr = np.hypot(H[k, k], H[k + 1, k])
            c[k], s[k] = (1.0, 0.0) if rr == 0.0 else (H[k, k] / rr, H[k + 1, k] / rr)
            H[k, k], H[k + 1, k] = rr, 0.0
            g[k], g[k + 1] = c[k] * g[k], -s[k] * g[k]
            its, kd = its + 1, k + 1
            hist.append(abs(g[k + 1]))
            if hist[-1] <= target:
                done = True
                break
        y = np.zeros(kd)
        for j in range(kd - 1, -1, -1):
            y[j] = (g[j] - H[j, j + 1:kd] @ y[j + 1:]) / H[j, j] if H[j, j] != 0.0 else 0.0
        x = x + y @ Z[:kd]
        if done:
            break
        r = b - Aop(x)
        beta = float(np.sqrt(r @ r))
        if beta <= target:
            break
    return x, its, hist


class OracleOuterBackend(OracleBackend):
    """OracleBackend + the outer pieces, with the interface of alfi_b200.multigrid.DeviceBackend."""

    def setup_outer(self, B, Minv, bc_dofs, remove_constant=True):
        self.B, self.Minv, self.bc = B.tocsr(), Minv.tocsr(), np.asarray(bc_dofs)
        self.remove_constant = remove_constant

    def schur_apply(self, nu, gamma, r):
        return schur_apply(self.apply, self.B, self.Minv, self.bc, nu, gamma, np.asarray(r, dtype=np.float64), self.remove_constant)

    def jacobian_apply(self, z):
        return jacobian_apply(self.levels[-1].A, self.B, self.bc, np.asarray(z, dtype=np.float64))

    def outer_solve(self, nu, gamma, rhs, rtol, atol, maxit=500, restart=30):
        return fgmres(self.jacobian_apply, lambda r: self.schur_apply(nu, gamma, r), np.asarray(rhs, dtype=np.float64),
                      rtol, atol, maxit, restart)
